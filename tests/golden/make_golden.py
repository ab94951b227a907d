"""Generate golden vectors by running the REFERENCE's own Python code in place.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):   python tests/golden/make_golden.py

``/root/reference/python/kdsource/{kde,geom,utils}.py`` are imported unmodified
from where they lie through a scratch package of symlinks.  The two
third-party modules they need and that are not installable here are shimmed:

* ``KDEpy.TreeKDE`` -> exact Gaussian sum taken from the oracle (so these
  vectors pin the reference's fold logic, score formula, grid search, kNN
  batching and the knn+mlcv chain, NOT KDEpy's kernel sum itself: that part
  stays "parity unpinned", see oracle/kds_oracle.c);
* ``joblib.Parallel`` is the real one (installed), sklearn is the real one.

Outputs: tests/golden/ref_kde.npz, ref_geom.npz (small, committed).
"""

import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/python/kdsource"

from oracle import kds_oracle as O  # noqa: E402


def import_reference():
    scratch = tempfile.mkdtemp(prefix="refpkg_")
    pkg = os.path.join(scratch, "refkds")
    os.makedirs(pkg)
    open(os.path.join(pkg, "__init__.py"), "w").close()
    for f in ("kde.py", "geom.py", "utils.py", "kdsource.py", "plist.py"):
        os.symlink(os.path.join(REF, f), os.path.join(pkg, f))

    class TreeKDE:  # shim with KDEpy's interface
        def __init__(self, kernel="gaussian", bw=1.0):
            self.kernel, self.bw = kernel, bw

        def fit(self, data, weights=None):
            self.data, self.weights = np.asarray(data), weights
            return self

        def evaluate(self, points):
            return O.kde_sum(self.data, self.weights, self.bw, points, kernel=self.kernel)

    shim = types.ModuleType("KDEpy")
    shim.TreeKDE = TreeKDE
    sys.modules["KDEpy"] = shim
    # `mcpl` (python module of the MCPL distribution, plist.py:10): MCPLFile over this
    # repository's own reader, which is pinned by the reference's 9-particle fixture
    from kdsource_b200 import mcpl as our_mcpl

    mshim = types.ModuleType("mcpl")
    mshim.MCPLFile = our_mcpl.MCPLFile
    sys.modules["mcpl"] = mshim
    # matplotlib is not installed: the plot_* methods import it only to draw
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("errorbar", "xscale", "yscale", "xlabel", "ylabel", "grid", "legend",
                 "tight_layout", "pcolormesh", "colorbar", "title", "plot", "show"):
        setattr(plt, name, lambda *a, **k: None)
    plt.gcf = lambda: None
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    col = types.ModuleType("matplotlib.colors")
    col.LogNorm = lambda *a, **k: None
    mpl.colors = col
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.colors": col})
    sys.path.insert(0, scratch)
    import refkds.geom as rgeom
    import refkds.kde as rkde
    return rkde, rgeom


def main():
    import joblib

    rkde, rgeom = import_reference()
    # loky workers cannot import the in-memory KDEpy shim: run the reference's
    # joblib.Parallel loop (kde.py:208-211) on threads instead (same results)
    with joblib.parallel_backend("threading"):
        kde_vectors(rkde, rgeom)
    geom_vectors(rgeom)
    fixture_files()
    kdsource_vectors(rgeom)
    print("wrote ref_kde.npz, ref_geom.npz, ref_kdsource.npz, ref_fixture_samples.mcpl.gz, "
          "ref_fixture_source.xml")


def kdsource_vectors(rgeom):
    """The reference's own KDSource object (python/kdsource/kdsource.py, plist.py, run
    unmodified) on a small MCPL list: fit (scaling, N_eff, Silverman bw), evaluate
    (densities + errors), and the marginal densities behind plot_integr / plot_E /
    plot2D_integr, for the Gaussian kernel and the finite-support ones."""
    import refkds.kdsource as rks
    import refkds.plist as rpl

    from kdsource_b200 import mcpl as our_mcpl

    out = {}
    # The one reference defect that has to be worked around to run evaluate() at all:
    # Geometry.jac reduces over the particle axis (geom.py:205, np.prod(jacs, axis=1)),
    # which returns one number per metric and makes kdsource.py:246 fail for any number
    # of query particles other than len(geom.ms).  The per-particle product over the
    # metrics (axis=0) is what the formula means; everything else runs unmodified.
    ref_jac = rgeom.Geometry.jac

    def jac_per_particle(self, parts):
        prod_axis1 = np.prod
        try:
            np.prod = lambda a, axis=None, **k: prod_axis1(a, axis=0 if axis == 1 else axis, **k)
            return ref_jac(self, parts)
        finally:
            np.prod = prod_axis1

    rgeom.Geometry.jac = jac_per_particle
    parts, w = O.synthetic_particles(3000, seed=21, wseed=22)
    tmp = tempfile.mkdtemp(prefix="refks_")
    fn = os.path.join(tmp, "list.mcpl.gz")
    our_mcpl.write_mcpl_file(fn, ekin=parts[:, 0], x=parts[:, 1], y=parts[:, 2], z=parts[:, 3],
                             ux=parts[:, 4], uy=parts[:, 5], uz=parts[:, 6], time=parts[:, 7],
                             weight=w, pdgcode=2112, srcname="golden")
    with open(fn, "rb") as fh:
        out["list_mcpl_gz"] = np.frombuffer(fh.read(), dtype=np.uint8)
    q = O.synthetic_particles(200, seed=23, wseed=24)[0]
    out["queries"] = q
    gE = np.logspace(-5, -1, 25)
    gx = np.linspace(-25, 25, 21)
    gy = np.linspace(-20, 20, 9)
    out["grid_E"], out["grid_x"], out["grid_y"] = gE, gx, gy
    vec0 = np.array([0, -100, -100, -1, -1, 0.2])
    vec1 = np.array([20, 100, 15, 1, 1, 1])
    out["vec0"], out["vec1"] = vec0, vec1
    for kern in ("gaussian", "epa", "box"):
        pl = rpl.PList(fn)
        geo = rgeom.Geometry([rgeom.Lethargy(10), rgeom.SurfXY(), rgeom.Isotrop()])
        s = rks.KDSource(pl, geo, bw="silv", J=3.5, kernel=kern)
        s.fit(importance=[2, 1, 1, 1, 1, 1])
        k = kern[0]
        out[k + "_scaling"] = s.scaling
        out[k + "_N_eff"] = s.N_eff
        out[k + "_bw"] = s.kde.bw
        ev, er = s.evaluate(q.copy())
        out[k + "_evaluate"], out[k + "_evaluate_err"] = ev, er
        _, (sc, e) = s.plot_integr("x", gx)
        out[k + "_integr_x"], out[k + "_integr_x_err"] = sc, e
        _, (sc, e) = s.plot_integr(1, gx, vec0=vec0, vec1=vec1, fact=2.0)
        out[k + "_integr_x_range"], out[k + "_integr_x_range_err"] = sc, e
        _, (sc, e) = s.plot_E(gE, vec0=vec0, vec1=vec1)
        out[k + "_E"], out[k + "_E_err"] = sc, e
        _, (sc, e) = s.plot2D_integr(["x", "y"], [gx, gy], vec0=vec0, vec1=vec1)
        out[k + "_2D_xy"], out[k + "_2D_xy_err"] = sc, e
    # time as the last metric (plot_t) with a Decade parametrisation
    parts_t = parts.copy()
    parts_t[:, 7] = np.random.default_rng(25).lognormal(0, 1, len(parts))
    fnt = os.path.join(tmp, "list_t.mcpl.gz")
    our_mcpl.write_mcpl_file(fnt, ekin=parts_t[:, 0], x=parts_t[:, 1], y=parts_t[:, 2],
                             z=parts_t[:, 3], ux=parts_t[:, 4], uy=parts_t[:, 5],
                             uz=parts_t[:, 6], time=parts_t[:, 7], weight=w, pdgcode=2112,
                             srcname="golden")
    with open(fnt, "rb") as fh:
        out["list_t_mcpl_gz"] = np.frombuffer(fh.read(), dtype=np.uint8)
    geo = rgeom.Geometry([rgeom.Lethargy(10), rgeom.SurfXY(), rgeom.Isotrop(), rgeom.Decade()])
    s = rks.KDSource(rpl.PList(fnt), geo, bw=0.4, J=1.0)
    s.fit()
    gt = np.logspace(-1.5, 1.5, 17)
    out["grid_t"] = gt
    _, (sc, e) = s.plot_t(gt)
    out["t_scaling"], out["t"], out["t_err"] = s.scaling, sc, e
    np.savez_compressed(os.path.join(HERE, "ref_kdsource.npz"), **out)


def fixture_files():
    """The only test fixture the reference ships: the 9-particle MCPL list and
    XML source embedded (base64) in python/kdsource/test.py:19-45, written out
    byte for byte by calling the reference's own accessor functions."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_test", os.path.join(REF, "test.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(HERE, "ref_fixture_samples.mcpl.gz"), "wb") as fh:
        fh.write(mod._test_samples_mcpl_gz_data())
    with open(os.path.join(HERE, "ref_fixture_source.xml"), "wb") as fh:
        fh.write(mod._test_source_xml_data())


def kde_vectors(rkde, rgeom):
    out = {}
    # ---- data: small synthetic list in scaled KDE variables -----------------
    parts, w = O.synthetic_particles(2400, seed=11, wseed=12)
    g = rgeom.Geometry([rgeom.Lethargy(10), rgeom.SurfXY(), rgeom.Isotrop()])
    vecs = g.transform(parts.copy())
    scaling = g.std(vecs=vecs, weights=w)
    data = vecs / scaling
    out["data"] = data
    out["weights"] = w
    out["scaling"] = scaling
    out["bw_silv_6_99011.96"] = rkde.bw_silv(6, 99011.96)
    # ---- kNN -----------------------------------------------------------------
    out["knn_k10_b600"] = rkde.bw_knn(data, weights=w, k=10, batch_size=600)
    out["knn_keff100_b800"] = rkde.bw_knn(data, weights=w, K_eff=100, batch_size=800)
    out["knn_k5_onebatch"] = rkde.bw_knn(data, weights=w, k=5, batch_size=len(data))
    # ---- CV score / MLCV / auto (N small: the shim is O(N^2) numpy) -----------
    n = 700
    d7, w7 = data[:n], w[:n]
    seed = rkde.bw_silv(6, np.sum(w7) ** 2 / np.sum(w7 ** 2))
    out["cv_n"] = n
    out["cv_scalar_bw"] = seed
    out["cv_score_scalar"] = rkde._kde_cv_score(seed, d7, weights=w7, n_splits=10)
    out["cv_score_scalar_k7"] = rkde._kde_cv_score(seed, d7, weights=w7, n_splits=7)
    out["cv_score_unweighted"] = rkde._kde_cv_score(seed, d7, weights=None, n_splits=10)
    bw_arr = out["knn_k10_b600"][:n]
    out["cv_score_array"] = rkde._kde_cv_score(bw_arr, d7, weights=w7, n_splits=10)
    out["cv_score_loo"] = rkde._kde_cv_score(seed, d7[:150], weights=w7[:150], n_splits=150)
    grid = np.logspace(-0.25, 0.35, 12)
    out["mlcv_grid"] = grid
    out["mlcv_scalar"] = rkde.bw_mlcv(d7, weights=w7, n_splits=10, seed=None, grid=grid,
                                      show=False)
    grid_a = np.logspace(-0.9, 0.1, 12)
    out["mlcv_grid_array"] = grid_a
    out["mlcv_scores_array"] = np.array(
        [rkde._kde_cv_score(gv * bw_arr, d7, weights=w7, n_splits=10) for gv in grid_a])
    print("array-seed scores", out["mlcv_scores_array"])
    out["mlcv_array"] = rkde.bw_mlcv(d7, weights=w7, n_splits=10, seed=bw_arr, grid=grid_a,
                                     show=False)
    out["mlcv_scores_scalar"] = np.array(
        [rkde._kde_cv_score(gv * seed, d7, weights=w7, n_splits=10) for gv in grid])
    out["auto"] = rkde.bw_auto(data, w, grid=grid_a, Nmlcv=700, Nbatch=600, Nknn=10, show=False,
                               n_splits=10)
    try:
        rkde.bw_mlcv(d7, weights=w7, grid=np.logspace(0.5, 0.8, 4), show=False)
        out["mlcv_edge_raises"] = False
    except Exception:
        out["mlcv_edge_raises"] = True
    np.savez_compressed(os.path.join(HERE, "ref_kde.npz"), **out)



def geom_vectors(rgeom):
    gout = {}
    rng = np.random.default_rng(5)
    P = np.empty((64, 8))
    P[:, 0] = rng.uniform(1e-6, 9.0, 64)
    P[:, 1:4] = rng.normal(0, 3, (64, 3))
    d = rng.normal(size=(64, 3))
    P[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    P[:, 7] = rng.uniform(0.1, 50, 64)
    wg = rng.uniform(0.5, 1.5, 64)
    gout["parts"] = P
    gout["w"] = wg
    cases = {
        "Energy": rgeom.Energy(), "Lethargy": rgeom.Lethargy(10), "Wavelength": rgeom.Wavelength(),
        "Time": rgeom.Time(), "Decade": rgeom.Decade(), "Vol": rgeom.Vol(-9, 9, -9, 9, -9, 9),
        "SurfXY": rgeom.SurfXY(z=0.5), "SurfR": rgeom.SurfR(z=0.5), "SurfR2": rgeom.SurfR2(z=0.5),
        "SurfCircle": rgeom.SurfCircle(z=0.5), "Isotrop": rgeom.Isotrop(keep_zdir=True),
        "Polar": rgeom.Polar(), "PolarMu": rgeom.PolarMu(),
    }
    for name, m in cases.items():
        sub = P[:, m.partvars]
        v = m.transform(sub.copy())
        v = v.reshape(len(P), -1)
        gout[name + "_transform"] = v
        gout[name + "_jac"] = np.asarray(m.jac(sub.copy()), dtype=np.float64)
        gout[name + "_std"] = m.std(vecs=v, weights=wg)
        try:  # several reference metrics cannot invert a 1-D mean vector
            gout[name + "_mean"] = np.asarray(m.mean(vecs=v, weights=wg), dtype=np.float64)
        except Exception:
            pass
        try:  # SurfR/SurfR2.inverse_transform raise in the reference (shape bug)
            gout[name + "_inverse"] = np.asarray(m.inverse_transform(v.copy()),
                                                 dtype=np.float64)
        except Exception:
            pass
    guide = rgeom.Guide(3.0, 5.0, 100.0, 2000.0)
    PG = P.copy()
    PG[:, 1] = np.where(rng.uniform(size=64) < 0.5, 1.5, -1.5)
    PG[:, 2] = rng.uniform(-2.4, 2.4, 64)
    PG[:, 3] = rng.uniform(1, 90, 64)
    gout["Guide_parts"] = PG
    gv = guide.transform(PG[:, guide.partvars].copy())
    gout["Guide_transform"] = gv
    gout["Guide_inverse"] = guide.inverse_transform(gv.copy())
    geo = rgeom.Geometry([rgeom.Lethargy(10), rgeom.SurfXY(), rgeom.Isotrop()],
                         trasl=[1.0, -2.0, 0.5], rot=[0.1, 0.2, -0.3])
    gout["Geometry_transform"] = geo.transform(P.copy())
    gout["Geometry_std"] = geo.std(vecs=gout["Geometry_transform"], weights=wg)
    gout["Geometry_inverse"] = geo.inverse_transform(gout["Geometry_transform"].copy())
    np.savez_compressed(os.path.join(HERE, "ref_geom.npz"), **gout)


if __name__ == "__main__":
    main()

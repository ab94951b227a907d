"""Host-side logic (no GPU): MCPL-3 I/O, metrics/geometry against the reference's
own outputs, XML source files, API surface, and the C-ABI library's exports."""

import ctypes
import gzip
import os
import re
import shutil

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURE = os.path.join(GOLDEN, "ref_fixture_samples.mcpl.gz")
FIXTURE_XML = os.path.join(GOLDEN, "ref_fixture_source.xml")


# ------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    from kdsource_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "kdsb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(kdsb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 19
    L = ctypes.CDLL(_lib.libpath())
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.EXPORTS) == declared
    lib = _lib.load()
    assert lib.kdsb200_version() >= 1000
    assert lib.kdsb200_packed_floats(1, 6) == 8 * 512
    assert lib.kdsb200_eval_mpad(1) == 1024 and lib.kdsb200_mlcv_mpad(257) == 512
    assert lib.kdsb200_mlcv_gtile(10) == 10 and lib.kdsb200_mlcv_gtile(11) == 16
    assert ctypes.sizeof(_lib.Geom) == 8 + 8 * ctypes.sizeof(_lib.Metric) + 16 + 48


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from kdsource_b200.treekde import TreeKDE

    k = TreeKDE(bw=0.3).fit(np.zeros((4, 2)))
    with pytest.raises(RuntimeError, match="CUDA device"):
        k.evaluate(np.zeros((1, 2)))
    from kdsource_b200 import kde

    with pytest.raises(RuntimeError, match="CUDA device"):
        kde.mlcv_scores(np.random.rand(50, 2), None, 0.3, [0.9, 1.0, 1.1])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "kdsource_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "kds_oracle" not in txt and "from oracle" not in txt, f
                assert "import oracle" not in txt, f


# -------------------------------------------------------------------- MCPL
def test_reference_fixture_decodes():
    from kdsource_b200 import mcpl

    f = mcpl.MCPLFile(FIXTURE)
    assert f.nparticles == 9 and f.hdr.opt_singleprec and f.hdr.particlesize == 36
    assert f.hdr.srcname == "ASCII SSV" and f.hdr.headersize == 61
    pb = f.read_block()
    assert len(pb) == 9 and f.read_block() is None
    assert np.all(pb.pdgcode == 2112) and np.all(pb.z == 0) and np.all(pb.time == 0)
    # first rows as decoded by hand in SURVEY section 4
    assert np.allclose(pb.ekin[:3], [0.0612344, 0.208316, 0.0267], rtol=2e-5)
    assert np.allclose(pb.x[:3], [-6.1461, -6.1838, 18.757], rtol=1e-4)
    assert np.allclose(pb.ux[:3], [0.096119, -0.12026, 0.072888], rtol=1e-4)
    assert np.allclose(pb.uz[:3], [0.96935, 0.98991, 0.99405], rtol=1e-4)
    assert np.allclose(pb.weight[:3], [0.948121, 0.970539, 1.25875], rtol=1e-5)
    assert np.allclose(pb.ux ** 2 + pb.uy ** 2 + pb.uz ** 2, 1, atol=1e-6)


@pytest.mark.parametrize("double_prec", [False, True])
@pytest.mark.parametrize("gz", [False, True])
def test_mcpl_write_read_roundtrip(tmp_path, double_prec, gz):
    from kdsource_b200 import mcpl

    rng = np.random.default_rng(0)
    n = 1000
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[0] = [1, 0, 0]
    d[1] = [0, -1, 0]
    d[2] = [0, 0, -1]
    d[3] = [0.6, 0.8, 0.0]  # z == 0 -> 1/z = inf branch
    ek = rng.uniform(0, 10, n)
    ek[4] = 0.0
    pos = rng.normal(size=(n, 3)) * 10
    tm = rng.uniform(0, 5, n)
    w = rng.uniform(0.1, 2, n)
    pdg = rng.choice([2112, 22, 11], n).astype(np.int32)
    fn = str(tmp_path / ("a.mcpl" + (".gz" if gz else "")))
    mcpl.write_mcpl_file(fn, ekin=ek, x=pos[:, 0], y=pos[:, 1], z=pos[:, 2], ux=d[:, 0],
                         uy=d[:, 1], uz=d[:, 2], time=tm, weight=w, pdgcode=pdg,
                         double_prec=double_prec)
    with mcpl.MCPLFile(fn, blocklength=300) as f:
        assert f.nparticles == n
        blocks = list(f.particle_blocks)
    assert [len(b) for b in blocks] == [300, 300, 300, 100]
    cat = lambda a: np.concatenate([getattr(b, a) for b in blocks])  # noqa: E731
    tol = 1e-14 if double_prec else 2e-7
    for name, ref in (("ekin", ek), ("x", pos[:, 0]), ("time", tm), ("weight", w)):
        assert np.allclose(cat(name), ref, rtol=tol, atol=tol), name
    got = np.stack((cat("ux"), cat("uy"), cat("uz")), axis=1)
    assert np.allclose(got, d, atol=5e-7 if not double_prec else 1e-14)
    assert np.array_equal(cat("pdgcode"), pdg)
    # skip_forward
    with mcpl.MCPLFile(fn, blocklength=10) as f:
        f.skip_forward(995)
        assert len(f.read_block()) == 5


def test_indexed_gzip_members_parallel_reader(tmp_path, monkeypatch):
    """The .gz files this package writes are sequences of gzip members that carry their own
    size ("KD" FEXTRA subfield): any gzip reader inflates them as one stream, MCPLFile hops
    from header to header and inflates the members in parallel -- same bytes; a file without
    the index (or with one foreign member) falls back to the streaming reader."""
    import gzip
    import zlib

    from kdsource_b200 import mcpl

    rng = np.random.default_rng(3)
    n = 200_000
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    fn = str(tmp_path / "idx.mcpl.gz")
    mcpl.write_mcpl_file(fn, ekin=rng.uniform(0, 10, n), x=d[:, 0], y=d[:, 1], z=d[:, 2],
                         ux=d[:, 0], uy=d[:, 1], uz=d[:, 2], time=rng.uniform(0, 5, n),
                         weight=rng.uniform(0.1, 2, n), pdgcode=2112)
    raw = open(fn, "rb").read()
    index = mcpl.gz_member_index(raw)
    assert index is not None and len(index[0]) >= 7  # header member + 1 MiB members
    offs, sizes, raws = index
    assert offs[0] == 0 and np.array_equal(offs[1:], np.cumsum(sizes)[:-1])
    assert offs[-1] + sizes[-1] == len(raw)
    whole = gzip.decompress(raw)
    assert int(raws.sum()) == len(whole)
    assert zlib.decompress(raw[offs[2]:offs[2] + sizes[2]], 31) == \
        whole[int(raws[:2].sum()):int(raws[:3].sum())]
    with mcpl.MCPLFile(fn) as f:
        assert isinstance(f._fh, mcpl._IndexedGzReader)
        buf, cnt = f.read_raw_bytes(n)
        assert cnt == n and bytes(buf) == whole[len(whole) - len(buf):]
        f.rewind()
        f.skip_forward(n - 3)
        assert len(f.read_block()) == 3
    # skipping hops over whole members without inflating them
    monkeypatch.setattr(mcpl._IndexedGzReader, "WINDOW_BYTES", 1 << 20)
    inflated = []
    real_inflate = mcpl._IndexedGzReader._inflate
    monkeypatch.setattr(mcpl._IndexedGzReader, "_inflate",
                        lambda self, k: inflated.append(k) or real_inflate(self, k))
    for skip in (0, 1, 29_127, 29_128, 100_000, n - 6):
        with mcpl.MCPLFile(fn) as f:
            f.read_raw_bytes(5)
            del inflated[:]
            f.skip_forward(skip)
            got, c = f.read_raw_bytes(7)
            lo = (5 + skip) * 32
            assert c == min(7, n - 5 - skip) and bytes(got) == bytes(buf[lo:lo + c * 32]), skip
            if skip == 100_000:  # 3.2 MB further on: two whole members were never inflated
                assert len(inflated) <= 2 and min(inflated) >= 4, inflated
    monkeypatch.undo()
    # small reads crossing member boundaries
    rd = mcpl._open_maybe_gz(fn)
    got = b"".join(iter(lambda: rd.read(777_777), b""))
    rd.close()
    assert got == whole
    # no index -> streaming gzip; index broken by a foreign member -> streaming gzip
    plain = str(tmp_path / "plain.mcpl.gz")
    open(plain, "wb").write(gzip.compress(whole, 1))
    mixed = str(tmp_path / "mixed.mcpl.gz")
    open(mixed, "wb").write(raw + gzip.compress(b"", 1))
    for other in (plain, mixed):
        with mcpl.MCPLFile(other) as f:
            assert not isinstance(f._fh, mcpl._IndexedGzReader)
            buf2, cnt2 = f.read_raw_bytes(n)
            assert cnt2 == n and bytes(buf2) == bytes(buf)


def test_mcpl_universal_fields_polarisation_userflags(tmp_path):
    from kdsource_b200 import mcpl

    n = 17
    one = np.ones(n)
    fn = str(tmp_path / "u.mcpl")
    mcpl.write_mcpl_file(fn, ekin=one, x=one, y=one, z=one, ux=0 * one, uy=0 * one, uz=one,
                         time=one, weight=2.5, pdgcode=22, polx=0.1 * one, poly=0.2 * one,
                         polz=0.3 * one, userflags=np.arange(n, dtype=np.uint32))
    f = mcpl.MCPLFile(fn)
    assert f.hdr.opt_universalpdgcode == 22 and f.hdr.opt_universalweight == 2.5
    assert f.hdr.particlesize == 12 + 12 + 8 + 4 + 4 + 4  # pol, pos, dirpack, ekin, time, flags
    pb = f.read_block()
    assert np.all(pb.weight == 2.5) and np.all(pb.pdgcode == 22)
    assert np.allclose(pb.poly, 0.2) and np.array_equal(pb.userflags, np.arange(n))


def test_write_mcpl_api(tmp_path):
    from kdsource_b200 import mcpl
    from kdsource_b200.api import write_mcpl

    n = 5
    a = np.linspace(0.1, 1, n)
    kw = dict(ekin=a, x=a, y=a, z=a, ux=0 * a, uy=0 * a, uz=1 + 0 * a, time=a)
    out = tmp_path / "w.mcpl"
    write_mcpl(str(out), weight=1.0, pdgcode=2112, **kw)
    assert (tmp_path / "w.mcpl.gz").is_file()
    f = mcpl.MCPLFile(str(tmp_path / "w.mcpl.gz"))
    assert f.nparticles == n and f.hdr.srcname == "KDSource MCPL writer"
    assert f.hdr.particlesize == 28  # universal weight and pdg code
    with pytest.raises(RuntimeError, match="Already exist"):
        write_mcpl(str(out), weight=1.0, pdgcode=2112, **kw)
    write_mcpl(str(out), weight=1.0, pdgcode=2112, force=True, **kw)
    with pytest.raises(RuntimeError, match="does not end with"):
        write_mcpl(str(tmp_path / "w.dat"), weight=1.0, pdgcode=2112, **kw)
    with pytest.raises(RuntimeError, match="same length"):
        write_mcpl(str(tmp_path / "x.mcpl"), weight=np.ones(3), pdgcode=2112, **kw)
    with pytest.raises(RuntimeError, match="all or none"):
        write_mcpl(str(tmp_path / "y.mcpl"), weight=1.0, pdgcode=2112, polx=a, **kw)


# ----------------------------------------------------------------- geometry
_NAMES = ["Energy", "Lethargy", "Wavelength", "Time", "Decade", "Vol", "SurfXY", "SurfR",
          "SurfR2", "SurfCircle", "Isotrop", "Polar", "PolarMu"]


def _make(name):
    from kdsource_b200 import geom as G

    return {
        "Energy": G.Energy(), "Lethargy": G.Lethargy(10), "Wavelength": G.Wavelength(),
        "Time": G.Time(), "Decade": G.Decade(), "Vol": G.Vol(-9, 9, -9, 9, -9, 9),
        "SurfXY": G.SurfXY(z=0.5), "SurfR": G.SurfR(z=0.5), "SurfR2": G.SurfR2(z=0.5),
        "SurfCircle": G.SurfCircle(z=0.5), "Isotrop": G.Isotrop(keep_zdir=True),
        "Polar": G.Polar(), "PolarMu": G.PolarMu(),
    }[name]


@pytest.mark.parametrize("name", _NAMES)
def test_metrics_vs_reference(golden_geom, name):
    g = golden_geom
    m = _make(name)
    P, w = g["parts"], g["w"]
    sub = P[:, m.partvars]
    v = np.asarray(m.transform(sub.copy())).reshape(len(P), -1)
    assert np.array_equal(v, g[name + "_transform"])
    assert np.array_equal(np.asarray(m.jac(sub.copy()), dtype=np.float64).reshape(-1),
                          g[name + "_jac"].reshape(-1))
    assert np.array_equal(m.std(vecs=v, weights=w), g[name + "_std"])
    if name + "_mean" in g.files:
        assert np.array_equal(np.asarray(m.mean(vecs=v, weights=w), np.float64), g[name + "_mean"])
    if name + "_inverse" in g.files:
        assert np.array_equal(np.asarray(m.inverse_transform(v.copy()), np.float64),
                              g[name + "_inverse"])
    else:  # SurfR / SurfR2: the reference raises; ours must invert the transform
        back = m.inverse_transform(v.copy())
        assert np.allclose(back[:, :2], sub[:, :2], atol=1e-12)


def test_guide_and_geometry_vs_reference(golden_geom):
    from kdsource_b200 import geom as G

    g = golden_geom
    guide = G.Guide(3.0, 5.0, 100.0, 2000.0)
    gv = guide.transform(g["Guide_parts"][:, guide.partvars].copy())
    assert np.allclose(gv, g["Guide_transform"], rtol=0, atol=0)
    assert np.array_equal(guide.inverse_transform(gv.copy()), g["Guide_inverse"])
    geo = G.Geometry([G.Lethargy(10), G.SurfXY(), G.Isotrop()], trasl=[1.0, -2.0, 0.5],
                     rot=[0.1, 0.2, -0.3])
    assert geo.dim == 6 and geo.varnames == ["u", "x", "y", "dx", "dy", "dz"]
    assert geo.volunits == "MeV cm^2 [dir]^3"
    vecs = geo.transform(g["parts"].copy())
    assert vecs.flags.c_contiguous  # consumers upload / stage it without a strided copy
    assert np.array_equal(vecs, g["Geometry_transform"])
    assert np.array_equal(geo.std(vecs=vecs, weights=g["w"]), g["Geometry_std"])
    assert np.array_equal(geo.inverse_transform(vecs.copy()), g["Geometry_inverse"])
    # jacobian: one value per particle (product over metrics)
    jac = geo.jac(g["parts"].copy())
    assert jac.shape == (len(g["parts"]),)
    assert np.allclose(jac, 1 / g["parts"][:, 0])


def test_geometry_xml_roundtrip():
    from xml.etree.ElementTree import Element, fromstring, tostring

    from kdsource_b200 import geom as G

    geo = G.Geometry([G.Lethargy(12.5), G.SurfXY(-3, 3, -np.inf, np.inf, 1.5),
                      G.Isotrop(keep_zdir=True), G.Decade()], trasl=[1, 2, 3], rot=[0, 0, 0.5])
    el = Element("Geom")
    geo.save(el)
    txt = tostring(el).decode()
    assert 'order="4"' in txt and '<params nps="5">-3 3 -inf inf 1.5</params>' in txt
    assert '<params nps="3">0 0 1</params>' in txt and '<params nps="0" />' in txt
    back = G.Geometry.load(fromstring(txt))
    assert [m.__class__.__name__ for m in back.ms] == ["Lethargy", "SurfXY", "Isotrop", "Decade"]
    assert back.ms[0].E0 == 12.5 and back.ms[1].z == 1.5 and back.ms[2].keep_zdir == 1
    assert np.allclose(back.trasl, [1, 2, 3]) and np.allclose(back.rot.as_rotvec(), [0, 0, 0.5])
    for name in G.METRIC_ORDER:
        assert name in G._metrics
    with pytest.raises(Exception, match="Invalid metricname"):
        G.Geometry.load(fromstring('<Geom order="1"><Nope><dim>1</dim><params nps="0"/></Nope>'
                                   '<trasl/><rot/></Geom>'))
    with pytest.raises(Exception, match="Invalid metric tree"):
        G.Lethargy.load(fromstring('<Lethargy><dim>2</dim><params nps="1">10</params></Lethargy>'))
    # the four factory geometries
    assert G.GeomFlat().dim == 6 and G.GeomFlatTemp().dim == 7
    assert G.GeomGuide(3, 5).dim == 5 and G.GeomActiv().dim == 7


# ------------------------------------------------------- PList / KDSource XML
@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    shutil.copy(FIXTURE, tmp_path / "samples.mcpl.gz")
    shutil.copy(FIXTURE_XML, tmp_path / "source.xml")
    monkeypatch.chdir(tmp_path)
    return tmp_path


def test_plist(workdir):
    from kdsource_b200.api import PList

    pl = PList("samples.mcpl.gz")
    assert pl.N == 9 and pl.params_set
    parts, ws = pl.get()
    assert parts.shape == (9, 8) and ws.shape == (9,)
    assert pl.N_eff == pytest.approx(ws.sum() ** 2 / (ws ** 2).sum())
    p2, w2 = pl.get(N=4, skip=3)
    assert np.array_equal(p2, parts[3:7]) and np.array_equal(w2, ws[3:7])
    assert PList("samples.mcpl.gz", pt="p", set_params=False).get()[0].shape == (0, 8)
    with pytest.raises(Exception, match="Empty particle list"):
        PList("samples.mcpl.gz", pt="p")
    tr = PList("samples.mcpl.gz", trasl=[1, 2, 3], rot=[0, 0, np.pi / 2], switch_x2z=True,
               set_params=False)
    pt, _ = tr.get()
    x, y, z = parts[:, 1] + 1, parts[:, 2] + 2, parts[:, 3] + 3
    rx, ry, rz = -y, x, z  # rotation by +90 deg about z
    assert np.allclose(pt[:, 1:4], np.stack((ry, rz, rx), axis=1))
    with pytest.raises(ValueError):
        PList("samples.mcpl.gz", trasl=[1, 2], set_params=False)


def test_kdsource_save_load_xml(workdir):
    import kdsource_b200.api as kds

    pl = kds.PList("samples.mcpl.gz")
    geo = kds.geom.Geometry([kds.geom.Lethargy(10), kds.geom.SurfXY(), kds.geom.PolarMu()])
    s = kds.KDSource(pl, geo, bw=0.5, J=3.0)
    with pytest.raises(Exception, match="Must fit"):
        s.evaluate(np.zeros((1, 8)))
    with pytest.raises(Exception, match="importance must have len"):
        s.fit(importance=[1, 2])
    s.fit(importance=[3, 1, 1, 1, 1])
    parts, ws = pl.get()
    std = geo.std(vecs=geo.transform(parts.copy()), weights=ws)
    assert np.allclose(s.scaling, std / [3, 1, 1, 1, 1])
    assert s.N_eff == pytest.approx(pl.N_eff)
    xml = s.save("out.xml", adjust_N=True)
    txt = open(xml).read()
    for tag in ("<J units=\"1/s\">3.0</J>", "<kernel>g</kernel>", "<pt>n</pt>", "<x2z>0</x2z>",
                'order="3"', "<Lethargy>", "<PolarMu>", '<BW variable="0">0.5</BW>'):
        assert tag in txt, tag
    s2 = kds.load("out.xml")
    assert np.allclose(s2.scaling, s.scaling, rtol=1e-8) and s2.kde.bw == 0.5 and s2.J == 3.0
    # adaptive bandwidth -> float32 side file, absolute path in <BW>
    s3 = kds.KDSource(pl, geo, bw=np.linspace(0.1, 0.9, 9))
    s3.fit()
    s3.save("ad.xml", bwfile="ad_bws")
    assert np.array_equal(np.fromfile("ad_bws", dtype="float32"),
                          np.linspace(0.1, 0.9, 9).astype("float32"))
    s4 = kds.load("ad.xml")
    assert np.allclose(s4.kde.bw, np.linspace(0.1, 0.9, 9), rtol=1e-7)
    # the reference's own fixture XML loads (positional schema, no <kernel>)
    ref = kds.load("source.xml")
    assert ref.kernel == "gaussian" and ref.kde.bw == pytest.approx(0.7372364669810315)
    assert [m.__class__.__name__ for m in ref.geom.ms] == ["Lethargy", "SurfXY", "PolarMu"]
    assert np.allclose(ref.scaling, [0.696868144, 9.6744063, 5.59415602, 9.52596915e-03,
                                     93.5705146])
    with pytest.raises(ValueError):
        kds.KDSource(pl, geo, bw=np.ones((3, 3)))
    with pytest.raises(Exception, match="Invalid bw_method"):
        kds.KDSource(pl, geo, bw=0.5).fit(bw_method="nope")


def test_resample_argument_validation(workdir):
    from kdsource_b200.api import resample_to_mcpl

    with pytest.raises(RuntimeError, match="Missing file"):
        resample_to_mcpl("nope.xml", "o.mcpl.gz", 10, rng=1)
    with pytest.raises(RuntimeError, match="Invalid"):
        resample_to_mcpl("source.xml", "o.mcpl.gz", 0, rng=1)
    with pytest.raises(RuntimeError, match="does not end with"):
        resample_to_mcpl("source.xml", "o.txt", 10, rng=1)
    open("exists.mcpl.gz", "wb").close()
    with pytest.raises(RuntimeError, match="Already exist"):
        resample_to_mcpl("source.xml", "exists.mcpl.gz", 10, rng=1)


# ------------------------------------------------------------ kde host logic
def test_kde_host_helpers(golden_kde):
    from sklearn.model_selection import KFold

    from kdsource_b200 import kde

    assert kde.bw_silv(6, 99011.96) == float(golden_kde["bw_silv_6_99011.96"])
    for N, K in [(23, 5), (700, 10), (150, 150)]:
        fb = kde.kfold_bounds(N, K)
        for f, (_, test) in enumerate(KFold(n_splits=K).split(np.zeros((N, 1)))):
            assert test[0] == fb[f] and test[-1] + 1 == fb[f + 1]
    for N, B in [(23, 4), (2400, 3)]:
        assert list(np.diff(kde.array_split_bounds(N, B))) == [
            len(c) for c in np.array_split(np.arange(N), B)]
    w = np.array([1.0, 2.0, 0.0, 1.0])
    v = np.random.default_rng(0).normal(size=(4, 3))
    assert kde.optimize_bw("silv", v, w) == kde.bw_silv(3, 16 / 6)
    with pytest.raises(Exception, match="Invalid bw_method"):
        kde.optimize_bw("nope", v, w)
    with pytest.raises(TypeError):
        kde.bw_knn(v, None)
    assert set(kde.bw_methods) == {"auto", "silv", "knn", "mlcv"}


def test_api_surface():
    import kdsource_b200.api as kds

    for name in ("KDSource", "load", "PList", "Geometry", "Metric", "bw_silv", "write_mcpl",
                 "resample_to_mcpl", "geom", "kde", "plist", "savessv", "appendssv",
                 "convert2mcpl", "join2mcpl"):
        assert hasattr(kds, name), name
    from kdsource_b200.utils import pdg2pt, pt2pdg

    assert pt2pdg("n") == 2112 and pt2pdg("p") == 22 and pt2pdg("e") == 11 and pt2pdg("x") == 0
    assert pdg2pt(2112) == "n" and pdg2pt(5) == "0"


def test_geom_struct_for_sampler():
    from kdsource_b200 import geom as G
    from kdsource_b200.sampler import make_geom_struct

    geo = G.Geometry([G.Lethargy(10), G.SurfXY(-1, 1, -2, 2, 0.5), G.Isotrop(keep_zdir=True)],
                     trasl=[1, 2, 3])
    gs = make_geom_struct(geo, [0.5, 2, 3, 0.1, 0.1, 0.1], "gaussian")
    assert gs.order == 3 and gs.kernel == b"g" and gs.has_trasl == 1 and gs.has_rot == 0
    assert [gs.ms[i].type for i in range(3)] == [1, 4, 9]
    assert list(gs.ms[1].params)[:5] == [-1, 1, -2, 2, 0.5]
    assert gs.ms[1].scaling[1] == 3.0 and gs.ms[2].nps == 3 and gs.ms[2].params[2] == 1.0
    with pytest.raises(ValueError):
        make_geom_struct(geo, [1, 2, 3], "g")


def test_gzip_table_builder_inflates_with_zlib():
    """Host half of the device gzip (kdsb200_gz_build_table): the dynamic-Huffman block
    header and the canonical codes must form a deflate stream zlib accepts -- packed
    here bit by bit in Python and inflated with zlib (no GPU involved)."""
    import ctypes
    import zlib

    from kdsource_b200 import _lib

    L = _lib.load()
    rng = np.random.default_rng(0)

    def roundtrip(data, rec=0, tail=0):
        """Pack the stream the way the device encoder does: literals, and in record mode
        one (length tail, distance rec) match where a record's tail repeats."""
        raw_in = np.frombuffer(data, np.uint8)
        hist = np.zeros(258, dtype=np.uint64)
        symbols = []  # byte values, 256 = end of block, 257 = tail match
        if rec:
            recs = raw_in.reshape(-1, rec)
            for i, r in enumerate(recs):
                m = i > 0 and np.array_equal(r[rec - tail:], recs[i - 1][rec - tail:])
                symbols += list(r[:rec - tail] if m else r) + ([257] if m else [])
        else:
            symbols = list(raw_in)
        hist[:258] = np.bincount(np.array(symbols + [257], dtype=np.int64), minlength=258)
        hist[257] -= 1
        hist[256] = 0
        T = _lib.GzTable()
        assert L.kdsb200_gz_build_table(hist.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                        rec, tail, ctypes.byref(T)) == 0
        lens = np.array(T.len[:])
        assert lens.min() >= 1 and lens.max() <= 15          # every symbol encodable
        if not T.match_nbits:
            assert np.sum(2.0 ** -lens) == pytest.approx(1.0)    # complete prefix code
        else:
            assert (T.rec_bytes, T.tail_bytes) == (rec, tail) and T.match_nbits <= 32
        bits = [(T.hdr[j >> 5] >> (j & 31)) & 1 for j in range(T.hdr_bits)]
        for b in symbols + [256]:
            if b == 257:
                bits += [(T.match_bits >> k) & 1 for k in range(T.match_nbits)]
            else:
                bits += [(T.code[b] >> k) & 1 for k in range(T.len[b])]
        bits += [0] * (-len(bits) % 8)
        raw = np.packbits(np.array(bits, dtype=np.uint8), bitorder="little").tobytes()
        assert zlib.decompress(raw, wbits=-15) == data
        return len(raw)

    roundtrip(rng.integers(0, 256, 4000, dtype=np.uint8).tobytes())
    assert roundtrip(bytes([7]) * 3000) < 600
    roundtrip(b"")
    # natural Huffman depths beyond 15 bits must be limited
    roundtrip(np.concatenate([np.full(2 ** k, k, dtype=np.uint8) for k in range(17)]).tobytes())
    # MCPL-3 records: constant weight / pdg / z / time fields compress, mantissas do not
    rec = np.zeros((500, 9), dtype=np.float32)
    rec[:, :6] = rng.normal(size=(500, 6))
    rec[:, 7] = 1.0
    raw = rec.view(np.uint8).copy()
    raw[:, 32:36] = np.frombuffer(np.int32(2112).tobytes(), np.uint8)
    plain = roundtrip(raw.tobytes())
    assert plain < 0.8 * raw.size
    # record mode: the constant (time, weight, pdgcode) tail becomes one match per record
    assert roundtrip(raw.tobytes(), rec=36, tail=12) < 0.9 * plain
    # other record layouts: long tails and distances with extra bits on both codes
    rec2 = rng.integers(0, 256, (300, 100), dtype=np.uint8)
    rec2[:, 40:] = rec2[0, 40:]
    rec2[17, 99] ^= 1  # one record whose tail differs
    assert roundtrip(rec2.tobytes(), rec=100, tail=60) < 0.55 * rec2.size
    roundtrip(rec2.tobytes(), rec=100, tail=4)


def test_plot_methods_host_logic(workdir):
    """Argument handling of the plot back-ends that runs before any kernel: fitted check,
    variable names, the [vec0, vec1] sample mask and its empty-range error (reference
    kdsource.py:447-471, 552-571), and the kernel names / roughness table."""
    import kdsource_b200.api as kds
    from kdsource_b200.kdsource import R

    pl = kds.PList("samples.mcpl.gz")
    geo = kds.geom.Geometry([kds.geom.Lethargy(10), kds.geom.SurfXY(), kds.geom.PolarMu()])
    for kern, r in (("gaussian", 1 / (2 * np.sqrt(np.pi))), ("epa", 3 / 5), ("box", 1 / 3)):
        s = kds.KDSource(pl, geo, bw=0.5, kernel=kern)
        assert s.R == pytest.approx(r) and s.R == R[kern[0]] and s.kde.kernel == kern
        for call in (lambda: s.plot_integr("x", np.linspace(-1, 1, 3)),
                     lambda: s.plot_E(np.logspace(-3, 0, 3)),
                     lambda: s.plot_t(np.linspace(0, 1, 3)),
                     lambda: s.plot2D_integr(["x", "y"], [np.zeros(2), np.zeros(2)]),
                     lambda: s.plot_point("x", np.zeros(3), np.zeros(7)),
                     lambda: s.plot2D_point(["x", "y"], [np.zeros(2), np.zeros(2)], np.zeros(7))):
            with pytest.raises(Exception, match="Must fit"):
                call()
    s = kds.KDSource(pl, geo, bw=0.5)
    s.fit()
    data = s.kde.data
    assert s._range_mask(None, None).all() and len(s._range_mask(None, None)) == len(data)
    lo = np.full(5, -np.inf)
    lo[1] = np.median(data[:, 1] * s.scaling[1])
    m = s._range_mask(lo, None)
    assert np.array_equal(m, data[:, 1] >= lo[1] / s.scaling[1]) and 0 < m.sum() < len(data)
    hi = np.full(5, np.inf)
    hi[1] = lo[1]
    assert not (s._range_mask(lo, hi) & ~(data[:, 1] * s.scaling[1] == lo[1])).any()
    with pytest.raises(Exception, match="No particles"):
        s.plot_E(np.logspace(-3, 0, 3), vec0=np.full(5, 1e9), vec1=np.full(5, 2e9))
    with pytest.raises(Exception, match="No particles"):
        s.plot2D_integr(["x", "y"], [np.zeros(2), np.zeros(2)], vec0=np.full(5, 1e9))
    with pytest.raises(KeyError):
        s.plot_integr("no_such_variable", np.zeros(3))
    with pytest.raises(ValueError):
        kds.KDSource(pl, geo, bw=np.zeros((2, 2)))

"""CPU oracle for KDSource's kernel-density hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kdsource_b200/`` may import this
module; the allowed users are ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

Two layers:

* numpy restatements of the reference's Python arithmetic
  (``python/kdsource/kde.py``, ``kdsource.py``, ``geom.py``), each citing the
  lines it follows; and
* ctypes bindings to ``oracle/kds_oracle.c`` (the same arithmetic in C with
  OpenMP, plus the C-side sampler restatement), and to ``oracle/_ref`` -- the
  reference's own C sources compiled in place -- used to pin the restatement.

Parity status: see the header of ``kds_oracle.c``.  The Gaussian kernel sum
restates KDEpy (absent from the reference tree): parity unpinned against KDEpy
itself (cross-checked against scikit-learn's and scipy's estimators);
everything around it -- including the reference's whole ``KDSource`` object --
is pinned by the reference's own code run in place
(``tests/golden/make_golden.py``) or by ``oracle/_ref``.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_u64_p = ctypes.POINTER(ctypes.c_uint64)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)

METRIC_NAMES = [
    "Energy", "Lethargy", "Wavelength", "Vol", "SurfXY", "SurfR", "SurfR2",
    "SurfCircle", "Guide", "Isotrop", "Polar", "PolarMu", "Time", "Decade",
]  # order of _metric_names, clib/src/geom.c:13-16
METRIC_ID = {n: i for i, n in enumerate(METRIC_NAMES)}


class OrcMetric(ctypes.Structure):
    _fields_ = [
        ("type", ctypes.c_int),
        ("dim", ctypes.c_int),
        ("scaling", ctypes.c_float * 6),
        ("nps", ctypes.c_int),
        ("params", ctypes.c_double * 8),
    ]


class OrcGeom(ctypes.Structure):
    _fields_ = [
        ("order", ctypes.c_int),
        ("ms", OrcMetric * 8),
        ("kernel", ctypes.c_char),
        ("has_trasl", ctypes.c_int),
        ("has_rot", ctypes.c_int),
        ("trasl", ctypes.c_double * 3),
        ("rot", ctypes.c_double * 3),
    ]


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def lib():
    """The C restatement (built by ``make -C oracle``)."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libkds_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle not built: run `make -C oracle`")
        L = ctypes.CDLL(path)
        L.orc_bw_silv.restype = ctypes.c_double
        L.orc_bw_silv.argtypes = [ctypes.c_int, ctypes.c_double]
        L.orc_rand_type_scripted.restype = ctypes.c_double
        L.orc_sizeof_geom.restype = ctypes.c_size_t
        assert L.orc_sizeof_geom() == ctypes.sizeof(OrcGeom)
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own C sources compiled in place (oracle/_ref), or None."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libkdsref.so")
        if not os.path.exists(path):
            return None
        R = ctypes.CDLL(path)
        R.ref_rand_type.restype = ctypes.c_double
        R.ref_metric_name.restype = ctypes.c_char_p
        _REF = R
    return _REF


# ----------------------------------------------------------------------
# A. Python-side hot path, numpy restatement
# ----------------------------------------------------------------------

def bw_silv(dim, N_eff):
    """Silverman's rule, python/kdsource/kde.py:18-38."""
    return (4 / (2 + dim)) ** (1 / (4 + dim)) * N_eff ** (-1 / (4 + dim))


def normalise_weights(weights, N):
    """KDEpy ``fit``: weights normalised to sum 1, None -> uniform."""
    if weights is None:
        return np.full(N, 1.0 / N)
    w = np.asarray(weights, dtype=np.float64)
    return w / w.sum()


def kde_sum(data, weights, bw, points, cutoff=None, chunk=256, kernel="gaussian",
            cutoff_rel=None):
    """Kernel sum (KDEpy Naive/TreeKDE.evaluate with norm=2), the arithmetic behind
    python/kdsource/kdsource.py:239 and kde.py:146-151:

        out[m] = sum_i w_i/sum(w) (2 pi)^(-d/2) h_i^-d exp(-|q_m-x_i|^2/(2 h_i^2))

    ``cutoff``: hard support radius (None = exact sum); ``cutoff_rel``: a hard radius
    ``cutoff_rel * h_i`` per source instead.  ``kernel`` "epa"/"box"
    (kdsource.py:60-65): KDEpy's radial finite-support kernels, published form
    (KDEpy is not in the tree -- restated from its documentation, pinned against
    scikit-learn's KernelDensity in tests/test_oracle.py): bandwidth = standard
    deviation of the 1-D kernel, so the support radius is R = h/sqrt(var) with
    var = 1/5 (epa), 1/3 (box); K = (d+2)/(2 V_d R^d)(1 - r^2/R^2) resp. 1/(V_d R^d)
    for r < R, V_d the volume of the unit d-ball.
    """
    if kernel in ("epa", "e", "box", "b"):
        from math import gamma

        data = np.asarray(data, dtype=np.float64)
        pts = np.asarray(points, dtype=np.float64)
        N, d = data.shape
        epa = kernel[0] == "e"
        wn = normalise_weights(weights, N)
        R = np.broadcast_to(np.asarray(bw, dtype=np.float64), (N,)) * np.sqrt(5.0 if epa else 3.0)
        Vd = np.pi ** (d / 2) / gamma(d / 2 + 1)
        coef = wn * ((d + 2) / 2 if epa else 1.0) / (Vd * R ** d)
        out = np.empty(len(pts))
        for s in range(0, len(pts), chunk):
            q = pts[s:s + chunk]
            r2 = np.zeros((len(q), N))
            for k in range(d):
                t = q[:, k:k + 1] - data[None, :, k]
                r2 += t * t
            u = r2 / R ** 2
            K = np.where(u < 1.0, coef * (1.0 - u) if epa else coef, 0.0)
            out[s:s + chunk] = K.sum(axis=1)
        return out
    data = np.asarray(data, dtype=np.float64)
    pts = np.asarray(points, dtype=np.float64)
    N, d = data.shape
    wn = normalise_weights(weights, N)
    h = np.broadcast_to(np.asarray(bw, dtype=np.float64), (N,))
    coef = wn * (2 * np.pi) ** (-d / 2) * h ** (-d)
    ih2 = 0.5 / h ** 2
    out = np.empty(len(pts))
    for s in range(0, len(pts), chunk):
        q = pts[s:s + chunk]
        r2 = np.zeros((len(q), N))
        for k in range(d):
            t = q[:, k:k + 1] - data[None, :, k]
            r2 += t * t
        K = coef * np.exp(-r2 * ih2)
        if cutoff is not None:
            K = np.where(r2 <= cutoff * cutoff, K, 0.0)
        if cutoff_rel is not None:  # per-source hard radius r_i = cutoff_rel * h_i
            K = np.where(r2 <= (cutoff_rel * h) ** 2, K, 0.0)
        out[s:s + chunk] = K.sum(axis=1)
    return out


def kde_sum_c(data, weights, bw, points, cutoff=None, nthreads=0):
    """Same as :func:`kde_sum`, evaluated by kds_oracle.c (OpenMP)."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    pts = np.ascontiguousarray(points, dtype=np.float64)
    N, d = data.shape
    wn = np.ascontiguousarray(normalise_weights(weights, N))
    h = np.ascontiguousarray(np.broadcast_to(np.asarray(bw, np.float64), (N,)))
    out = np.empty(len(pts))
    lib().orc_kde_sum(
        _p(data, c_double_p), _p(wn, c_double_p), _p(h, c_double_p),
        ctypes.c_int64(N), ctypes.c_int(d), _p(pts, c_double_p),
        ctypes.c_int64(len(pts)),
        ctypes.c_double(cutoff if cutoff is not None else -1.0),
        _p(out, c_double_p), ctypes.c_int(nthreads))
    return out


def treekde_radius(bw_max, atol=1e-3):
    """Support radius KDEpy's TreeKDE uses for the Gaussian kernel
    (``practical_support(max bw, atol)``): the root of the 1-D pdf
    N(0, bw) = atol on [0, 8 bw] (brentq, xtol=1e-3) plus xtol
    (from KDEpy>=1.1.12; not in the reference tree)."""
    from scipy.optimize import brentq

    def f(x):
        return np.exp(-0.5 * (x / bw_max) ** 2) / (bw_max * np.sqrt(2 * np.pi)) - atol

    xtol = 1e-3
    return brentq(f, 0.0, 8 * bw_max, xtol=xtol) + xtol


def kfold_bounds(N, K):
    """sklearn KFold(n_splits=K) without shuffle (kde.py:134-136): contiguous
    folds, the first N % K folds one longer.  Returns K+1 boundaries."""
    sizes = np.full(K, N // K, dtype=np.int64)
    sizes[: N % K] += 1
    return np.concatenate(([0], np.cumsum(sizes)))


def cv_score(bw, data, weights=None, n_splits=10):
    """_kde_cv_score, python/kdsource/kde.py:106-153 (Gaussian kernel, raw test
    weights, mean over folds of the mean over test points)."""
    data = np.asarray(data, dtype=np.float64)
    N = len(data)
    if np.ndim(bw) == 1 and len(bw) < N:
        raise ValueError("If bw is array, must have same len as data")
    fb = kfold_bounds(N, n_splits)
    scores = []
    for f in range(n_splits):
        t0, t1 = fb[f], fb[f + 1]
        train = np.r_[0:t0, t1:N]
        bw_tr = np.asarray(bw)[train] if np.ndim(bw) == 1 else bw
        w_tr = None if weights is None else np.asarray(weights)[train]
        dens = kde_sum(data[train], w_tr, bw_tr, data[t0:t1])
        with np.errstate(divide="ignore"):
            if weights is not None:
                scores.append(np.mean(np.asarray(weights)[t0:t1] * np.log(dens)))
            else:
                scores.append(np.mean(np.log(dens)))
    return np.mean(scores)


def kde_sum_tree(data, weights, bw, points, eps=1e-3):
    """KDEpy ``TreeKDE.evaluate`` as the reference runs it, restated from SURVEY appendix D
    (KDEpy>=1.1.12 is not in the tree): scipy ``cKDTree`` over the samples, one support
    radius from the largest bandwidth (:func:`treekde_radius`), approximate ball query
    ``query_ball_point(q, r, p=2, eps=eps*sqrt(N))`` per query point, Gaussian terms
    summed over the returned neighbours only.  This is the *algorithm* whose CPU time
    the reference arm of bench.py reports; its values differ from the exact sum by the
    truncation (SURVEY: 2.4 % median)."""
    from scipy.spatial import cKDTree

    data = np.ascontiguousarray(data, dtype=np.float64)
    pts = np.ascontiguousarray(points, dtype=np.float64)
    N, d = data.shape
    wn = normalise_weights(weights, N)
    h = np.broadcast_to(np.asarray(bw, dtype=np.float64), (N,))
    coef = wn * (2 * np.pi) ** (-d / 2) * h ** (-d)
    ih2 = 0.5 / h ** 2
    tree = cKDTree(data)
    r = treekde_radius(float(np.max(h)), atol=eps)
    out = np.zeros(len(pts))
    for m, idx in enumerate(tree.query_ball_point(pts, r, p=2, eps=eps * np.sqrt(N))):
        if idx:
            idx = np.asarray(idx)
            diff = data[idx] - pts[m]
            out[m] = np.sum(coef[idx] * np.exp(-np.einsum("ij,ij->i", diff, diff) * ih2[idx]))
    return out


def cv_score_tree(bw, data, weights=None, n_splits=10):
    """:func:`cv_score` with the TreeKDE-like truncated sum (one fold = one tree build +
    one evaluate of the test points, exactly the work of kde.py:134-151)."""
    data = np.asarray(data, dtype=np.float64)
    N = len(data)
    fb = kfold_bounds(N, n_splits)
    scores = []
    for f in range(n_splits):
        t0, t1 = fb[f], fb[f + 1]
        train = np.r_[0:t0, t1:N]
        bw_tr = np.asarray(bw)[train] if np.ndim(bw) == 1 else bw
        w_tr = None if weights is None else np.asarray(weights)[train]
        dens = kde_sum_tree(data[train], w_tr, bw_tr, data[t0:t1])
        with np.errstate(divide="ignore"):
            if weights is not None:
                scores.append(np.mean(np.asarray(weights)[t0:t1] * np.log(dens)))
            else:
                scores.append(np.mean(np.log(dens)))
    return np.mean(scores)


def mlcv_scores_tree(data, weights, seed, grid, n_splits=10, n_jobs=-1):
    """The reference's grid search as it runs on a CPU (kde.py:204-211): one process per
    grid point (joblib, ``n_jobs=-1``), each scoring ``grid[g] * seed`` with the
    TreeKDE-like estimator."""
    from joblib import Parallel, delayed

    seed = np.asarray(seed, dtype=np.float64)
    return np.array(Parallel(n_jobs=n_jobs)(
        delayed(cv_score_tree)(g * seed, data, weights, n_splits) for g in grid))


def mlcv_scores_c(data, weights, seed, grid, n_splits=10, nthreads=0):
    """CV scores for every grid factor (kde.py:204-211) via kds_oracle.c."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    N, d = data.shape
    seedv = np.ascontiguousarray(np.broadcast_to(np.asarray(seed, np.float64), (N,)))
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    n_splits = min(n_splits, N)
    out = np.empty(len(grid))
    lib().orc_mlcv_scores(
        _p(data, c_double_p), _p(w, c_double_p), _p(seedv, c_double_p),
        ctypes.c_int64(N), ctypes.c_int(d), _p(grid, c_double_p),
        ctypes.c_int(len(grid)), ctypes.c_int(n_splits), _p(out, c_double_p),
        ctypes.c_int(nthreads))
    return out


def bw_mlcv(data, weights=None, n_splits=10, seed=None, grid=None, use_c=True):
    """bw_mlcv, python/kdsource/kde.py:156-226 (without the plot)."""
    data = np.asarray(data, dtype=np.float64)
    if seed is None:
        N_eff = (len(data) if weights is None
                 else np.sum(weights) ** 2 / np.sum(np.asarray(weights) ** 2))
        seed = bw_silv(data.shape[1], N_eff)
    if grid is None:
        grid = np.logspace(-0.1, 0.1, 10)
    bw_grid = np.reshape(grid, (-1, *np.ones(np.ndim(seed), dtype=int))) * seed
    if n_splits > len(data):
        n_splits = len(data)
    if use_c:
        scores = mlcv_scores_c(data, weights, seed, grid, n_splits)
    else:
        scores = np.array([cv_score(b, data, weights, n_splits) for b in bw_grid])
    idx_best = int(np.argmax(scores))
    if idx_best in (0, len(bw_grid) - 1):
        raise Exception(
            "Maximum not found in bw range selected. Move grid and try again.")
    return bw_grid[idx_best], scores


def array_split_bounds(N, batches):
    """np.array_split boundaries (kde.py:91): first N % batches chunks one longer."""
    sizes = np.full(batches, N // batches, dtype=np.int64)
    sizes[: N % batches] += 1
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)


def knn(data, kk, batch_off=None, nthreads=0):
    """Brute-force (distance, index)-ordered kk nearest neighbours inside each
    contiguous batch (self included), the arithmetic of kde.py:93-97."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    N, d = data.shape
    if batch_off is None:
        batch_off = np.array([0, N], dtype=np.int64)
    batch_off = np.ascontiguousarray(batch_off, dtype=np.int64)
    dist = np.empty((N, kk))
    idx = np.empty((N, kk), dtype=np.int64)
    lib().orc_knn(_p(data, c_double_p), ctypes.c_int64(N), ctypes.c_int(d),
                  _p(batch_off, c_i64_p), ctypes.c_int(len(batch_off) - 1),
                  ctypes.c_int(kk), _p(dist, c_double_p), _p(idx, c_i64_p),
                  ctypes.c_int(nthreads))
    return dist, idx


def bw_knn(data, weights=None, K_eff=100, k=None, batch_size=10000):
    """bw_knn, python/kdsource/kde.py:41-103 (including the inverted N_eff
    branch: with weights N_eff = N; without weights the reference raises)."""
    data = np.asarray(data, dtype=np.float64)
    N, dim = data.shape
    batches = int(max(1, np.round(N / batch_size)))
    if weights is None:
        raise TypeError("reference bw_knn fails for weights=None (kde.py:70-74)")
    N_eff = N
    if k is None:
        k_float = batch_size * K_eff / N_eff
        k = int(max(1, round(k_float)))
        f_k = k_float / k
    else:
        f_k = 1.0
    off = array_split_bounds(N, batches)
    dist, _ = knn(data, k + 1, off)
    return dist[:, -1] * (f_k) ** (1 / dim)


def bw_auto(data, weights, grid=None, Nmlcv=1e4, Nbatch=1e4, Nknn=10, n_splits=10):
    """bw_auto, python/kdsource/kde.py:229-297."""
    data = np.asarray(data, dtype=np.float64)
    N, dim = data.shape
    if Nbatch == -1:
        Nbatch = N
    BW_knn = bw_knn(data, weights=weights, k=Nknn, batch_size=Nbatch)
    if Nmlcv == -1:
        Nmlcv = N
    Nmlcv = int(Nmlcv)
    seed = BW_knn[:Nmlcv]
    BW_mlcv, _ = bw_mlcv(data[:Nmlcv], weights=weights, seed=seed, grid=grid,
                         n_splits=n_splits)
    out = BW_knn * BW_mlcv[0] / BW_knn[0]
    out *= bw_silv(dim, len(BW_knn)) / bw_silv(dim, len(BW_mlcv))
    return out


def evaluate(parts_vecs, jacs, kde_data, kde_weights, kde_bw, scaling, J, N_eff,
             dim, R=1 / (2 * np.sqrt(np.pi))):
    """KDSource.evaluate arithmetic, python/kdsource/kdsource.py:236-248, given
    already-transformed query vectors and jacobians."""
    scaling = np.asarray(scaling, dtype=np.float64)
    evals = 1 / np.prod(scaling) * kde_sum(kde_data, kde_weights, kde_bw,
                                           parts_vecs / scaling)
    errs = np.sqrt(evals * R ** dim / (N_eff * np.mean(kde_bw) * np.prod(scaling)))
    evals = evals * J * jacs
    errs = errs * J * jacs
    return evals, errs


# ----------------------------------------------------------------------
# B/C. C-side hot path (sampler), via kds_oracle.c
# ----------------------------------------------------------------------

def make_geom(metrics, kernel="g", trasl=None, rot=None):
    """metrics: list of (name, scaling list, params list)."""
    g = OrcGeom()
    g.order = len(metrics)
    for i, (name, scaling, params) in enumerate(metrics):
        m = g.ms[i]
        m.type = METRIC_ID[name]
        m.dim = len(scaling)
        for j, s in enumerate(scaling):
            m.scaling[j] = np.float32(s)
        m.nps = len(params)
        for j, p in enumerate(params):
            m.params[j] = float(p)
    g.kernel = kernel.encode()[:1]
    g.has_trasl = int(trasl is not None)
    g.has_rot = int(rot is not None)
    if trasl is not None:
        for j in range(3):
            g.trasl[j] = float(trasl[j])
    if rot is not None:
        for j in range(3):
            g.rot[j] = float(rot[j])
    return g


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (ctypes.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (ctypes.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def weights_cdf(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    cdf = np.empty(len(w), dtype=np.uint64)
    lib().orc_weights_cdf(_p(w, c_double_p), ctypes.c_int64(len(w)),
                          ctypes.c_double(float(np.sum(w))), _p(cdf, c_u64_p))
    return cdf


def resample(src8, weights, bw, geom, nout, seed=0, first=0, pdg=2112,
             ext_uniforms=None, perturb=True, want=("idx", "parts", "rec"), contract=2):
    """This project's resampling contract (prefix-sum + binary search pick,
    per-particle Philox stream, reference perturbation maps)."""
    src8 = np.ascontiguousarray(src8, dtype=np.float64)
    N = len(src8)
    cdf = weights_cdf(weights)
    bwv = None
    bws = 0.0
    if np.ndim(bw) == 1:
        bwv = np.ascontiguousarray(bw, dtype=np.float32)
    else:
        bws = float(bw)
    slots = 0
    if ext_uniforms is not None:
        ext_uniforms = np.ascontiguousarray(ext_uniforms, dtype=np.float64)
        slots = ext_uniforms.shape[1]
    idx = np.empty(nout, dtype=np.int64) if "idx" in want else None
    parts = np.empty((nout, 8)) if "parts" in want else None
    rec = np.empty((nout, 36), dtype=np.uint8) if "rec" in want else None
    lib().orc_resample(
        _p(src8, c_double_p), _p(cdf, c_u64_p), _p(bwv, c_float_p),
        ctypes.c_double(bws), ctypes.c_int64(N), ctypes.byref(geom),
        ctypes.c_int32(pdg), ctypes.c_uint64(seed), ctypes.c_uint64(first),
        ctypes.c_uint64(nout), _p(ext_uniforms, c_double_p), ctypes.c_int(slots),
        ctypes.c_int(int(perturb)), ctypes.c_int(int(contract)), _p(rec, c_u8_p),
        _p(idx, c_i64_p),
        _p(parts, c_double_p))
    return {"idx": idx, "parts": parts, "rec": rec, "cdf": cdf}


def metric_perturb_scripted(name, scaling, params, kernel, bw, part8, uniforms,
                            use_ref=False):
    """One metric perturbation with a scripted uniform stream.  use_ref=True
    runs the reference's own compiled X_perturb (oracle/_ref)."""
    scaling = np.ascontiguousarray(scaling, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    p = np.array(part8, dtype=np.float64)
    L = ref_lib() if use_ref else lib()
    fn = L.ref_metric_perturb if use_ref else L.orc_metric_perturb_scripted
    fn.restype = ctypes.c_int
    used = fn(ctypes.c_int(METRIC_ID[name]), ctypes.c_int(len(scaling)),
              _p(scaling, c_double_p), ctypes.c_int(len(params)),
              _p(params, c_double_p) if len(params) else None,
              ctypes.c_char(kernel.encode()[:1]), ctypes.c_double(bw),
              _p(p, c_double_p), _p(uniforms, c_double_p),
              ctypes.c_int(len(uniforms)))
    return p, used


def rand_type_scripted(kernel, uniforms, use_ref=False):
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    used = ctypes.c_int(0)
    L = ref_lib() if use_ref else lib()
    fn = L.ref_rand_type if use_ref else L.orc_rand_type_scripted
    fn.restype = ctypes.c_double
    v = fn(ctypes.c_char(kernel.encode()[:1]), _p(uniforms, c_double_p),
           ctypes.c_int(len(uniforms)), ctypes.byref(used))
    return v, used.value


def rotv(v, rv, inverse, use_ref=False):
    v = np.array(v, dtype=np.float64)
    rv = np.ascontiguousarray(rv, dtype=np.float64)
    L = ref_lib() if use_ref else lib()
    fn = L.ref_rotv if use_ref else L.orc_rotv
    fn.restype = None
    fn(_p(v, c_double_p), _p(rv, c_double_p), ctypes.c_int(int(inverse)))
    return v


def ref_sampler(parts8, weights, pdg, metrics, kernel, bw, nout, seed=1, pt="n",
                bwfile=None, plist_trasl=None, plist_rot=None, x2z=False,
                geom_trasl=None, geom_rot=None, uniforms=None, nskip=1000):
    """Run the REFERENCE's own C sampler (oracle/_ref) over an in-memory list:
    KDS_create -> KDS_rand_w_mean(1000) -> nout x KDS_rand_sample2, the call
    sequence of clib/src/resamplemcpl.c:35-43."""
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref not built")
    parts8 = np.ascontiguousarray(parts8, dtype=np.float64)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    pdg = np.ascontiguousarray(np.broadcast_to(pdg, (len(parts8),)), dtype=np.int32)
    types = np.array([METRIC_ID[m[0]] for m in metrics], dtype=np.int32)
    dims = np.array([len(m[1]) for m in metrics], dtype=np.int32)
    scal = np.array(sum([list(m[1]) for m in metrics], []), dtype=np.float64)
    npss = np.array([len(m[2]) for m in metrics], dtype=np.int32)
    pars = np.array(sum([list(m[2]) for m in metrics], []) + [0.0], dtype=np.float64)
    out8 = np.empty((nout, 8))
    outw = np.empty(nout)

    def v3(a):
        return None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    pt_, pr_, gt_, gr_ = v3(plist_trasl), v3(plist_rot), v3(geom_trasl), v3(geom_rot)
    uni = None if uniforms is None else np.ascontiguousarray(uniforms, np.float64)
    R.ref_sampler_run.restype = ctypes.c_int
    wraps = R.ref_sampler_run(
        _p(parts8, c_double_p), _p(weights, c_double_p), _p(pdg, c_i32_p),
        ctypes.c_int64(len(parts8)), ctypes.c_char(pt.encode()),
        ctypes.c_int(len(metrics)), _p(types, c_int_p), _p(dims, c_int_p),
        _p(scal, c_double_p), _p(npss, c_int_p), _p(pars, c_double_p),
        ctypes.c_char(kernel.encode()[:1]), ctypes.c_double(float(bw)),
        bwfile.encode() if bwfile else None, _p(pt_, c_double_p),
        _p(pr_, c_double_p), ctypes.c_int(int(x2z)), _p(gt_, c_double_p),
        _p(gr_, c_double_p), ctypes.c_uint64(seed), _p(uni, c_double_p),
        ctypes.c_int(0 if uni is None else len(uni)), ctypes.c_int64(nout),
        ctypes.c_int(nskip), _p(out8, c_double_p), _p(outw, c_double_p))
    return out8, outw, wraps


def ref_sampler_cli(parts8, weights, pdg, metrics, kernel, bw, nout, seed=1, pt="n",
                    bwfile=None, nskip=1000):
    """The reference's C sampler driven like ``kdsource-resample`` drives it
    (python/kdsource/resample.py:33-43, _chooks.py:29-41): every uniform is a ctypes
    callback into Python's ``random.Random(seed).random``.  Returns (parts, weights)."""
    import random

    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref not built")
    parts8 = np.ascontiguousarray(parts8, dtype=np.float64)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    pdg = np.ascontiguousarray(np.broadcast_to(pdg, (len(parts8),)), dtype=np.int32)
    types = np.array([METRIC_ID[m[0]] for m in metrics], dtype=np.int32)
    dims = np.array([len(m[1]) for m in metrics], dtype=np.int32)
    scal = np.array(sum([list(m[1]) for m in metrics], []), dtype=np.float64)
    npss = np.array([len(m[2]) for m in metrics], dtype=np.int32)
    pars = np.array(sum([list(m[2]) for m in metrics], []) + [0.0], dtype=np.float64)
    out8 = np.empty((nout, 8))
    outw = np.empty(nout)
    cb = ctypes.CFUNCTYPE(ctypes.c_double)(random.Random(seed).random)
    R.ref_sampler_run_stateless.restype = ctypes.c_int
    R.ref_sampler_run_stateless(
        cb, _p(parts8, c_double_p), _p(weights, c_double_p), _p(pdg, c_i32_p),
        ctypes.c_int64(len(parts8)), ctypes.c_char(pt.encode()), ctypes.c_int(len(metrics)),
        _p(types, c_int_p), _p(dims, c_int_p), _p(scal, c_double_p), _p(npss, c_int_p),
        _p(pars, c_double_p), ctypes.c_char(kernel.encode()[:1]), ctypes.c_double(float(bw)),
        bwfile.encode() if bwfile else None, ctypes.c_int64(nout), ctypes.c_int(nskip),
        _p(out8, c_double_p), _p(outw, c_double_p))
    return out8, outw


def hilbert_keys(data, lo, hi, bits):
    """Hilbert-curve index of every row of `data` over the box [lo, hi] quantised to `bits`
    per dimension (Skilling, "Programming the Hilbert curve", 2004: axes -> transpose,
    then bit interleave).  Checker for the ordering the sorted-tile layout uses; the
    ordering has no counterpart in the reference (it only changes the summation order)."""
    data = np.asarray(data, dtype=np.float64)
    n, D = data.shape
    span = np.asarray(hi, dtype=np.float64) - np.asarray(lo, dtype=np.float64)
    scale = np.where(span > 0, float(1 << bits) / np.where(span > 0, span, 1.0), 0.0)
    X = np.clip(np.floor((data - lo) * scale), 0, (1 << bits) - 1).astype(np.uint64)
    M = np.uint64(1) << np.uint64(bits - 1)
    one = np.uint64(1)
    Q = M
    while Q > one:
        P = Q - one
        for i in range(D):
            bit = (X[:, i] & Q) != 0
            X[bit, 0] ^= P
            t = (X[:, 0] ^ X[:, i]) & P
            t[bit] = 0
            X[:, 0] ^= t
            X[:, i] ^= t
        Q >>= one
    for i in range(1, D):
        X[:, i] ^= X[:, i - 1]
    t = np.zeros(n, dtype=np.uint64)
    Q = M
    while Q > one:
        sel = (X[:, D - 1] & Q) != 0
        t[sel] ^= Q - one
        Q >>= one
    for i in range(D):
        X[:, i] ^= t
    key = np.zeros(n, dtype=np.uint64)
    for b in range(bits - 1, -1, -1):
        for i in range(D):
            key = (key << one) | ((X[:, i] >> np.uint64(b)) & one)
    return key


# ----------------------------------------------------------------------
# synthetic data generator (Verification.ipynb cell 4, seeded)
# ----------------------------------------------------------------------

def synthetic_particles(N, seed=12345, wseed=777, z_spread=0.0):
    """Two-cluster correlated neutron list: columns [ekin,x,y,z,ux,uy,uz,t],
    weights ~ max(N(1,0.1), 1e-99)."""
    rng = np.random.default_rng(seed)
    n1 = N // 2
    n2 = N - n1
    u = np.concatenate((rng.normal(5, 1, n1), rng.normal(9, 1, n2)))
    E = 10 * np.exp(-u)
    x = np.concatenate((rng.normal(10, 10, n1), rng.normal(-10, 10, n2)))
    y = rng.normal(0, 10, N)
    z = rng.normal(0, z_spread, N) if z_spread > 0 else np.zeros(N)
    mu = np.sqrt(rng.uniform(0, 1, N))
    phi = rng.uniform(-np.pi, np.pi, N)
    dxy = np.sqrt(1 - mu ** 2)
    parts = np.stack((E, x, y, z, dxy * np.cos(phi), dxy * np.sin(phi), mu,
                      np.zeros(N)), axis=1)
    perm = rng.permutation(N)
    parts = parts[perm]
    w = np.maximum(np.random.default_rng(wseed).normal(1, 0.1, N), 1e-99)
    return parts, w

"""The C sampling interface (include/kdsource_compat.h) served by libkdsb200.so:
host-only pieces on CPU, sampling on the GPU."""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
LIB = os.path.join(ROOT, "kdsource_b200", "libkdsb200.so")


class McplParticle(ctypes.Structure):
    _fields_ = [("ekin", ctypes.c_double), ("polarisation", ctypes.c_double * 3),
                ("position", ctypes.c_double * 3), ("direction", ctypes.c_double * 3),
                ("time", ctypes.c_double), ("weight", ctypes.c_double),
                ("pdgcode", ctypes.c_int32), ("userflags", ctypes.c_uint32)]


class PListS(ctypes.Structure):
    _fields_ = [("pt", ctypes.c_char), ("pdgcode", ctypes.c_int), ("filename", ctypes.c_char_p),
                ("file", ctypes.c_void_p), ("trasl", ctypes.POINTER(ctypes.c_double)),
                ("rot", ctypes.POINTER(ctypes.c_double)), ("x2z", ctypes.c_int),
                ("part", ctypes.POINTER(McplParticle)), ("npts", ctypes.c_longlong)]


class GeomS(ctypes.Structure):
    _fields_ = [("ord", ctypes.c_int), ("ms", ctypes.c_void_p), ("bwfilename", ctypes.c_char_p),
                ("bwfile", ctypes.c_void_p), ("bw", ctypes.c_double), ("kernel", ctypes.c_char)]


class KDSourceS(ctypes.Structure):
    _fields_ = [("J", ctypes.c_double), ("kernel", ctypes.c_char),
                ("plist", ctypes.POINTER(PListS)), ("geom", ctypes.POINTER(GeomS)),
                ("impl", ctypes.c_void_p)]


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    shutil.copy(os.path.join(GOLDEN, "ref_fixture_samples.mcpl.gz"), tmp_path / "samples.mcpl.gz")
    shutil.copy(os.path.join(GOLDEN, "ref_fixture_source.xml"), tmp_path / "source.xml")
    monkeypatch.chdir(tmp_path)
    return tmp_path


def _lib():
    L = ctypes.CDLL(LIB)
    L.KDS_open.restype = ctypes.POINTER(KDSourceS)
    L.KDS_open.argtypes = [ctypes.c_char_p]
    L.KDS_destroy.argtypes = [ctypes.POINTER(KDSourceS)]
    L.KDS_w_mean.restype = ctypes.c_double
    L.KDS_w_mean.argtypes = [ctypes.POINTER(KDSourceS), ctypes.c_int, ctypes.c_void_p]
    L.KDS_sample2.argtypes = [ctypes.POINTER(KDSourceS), ctypes.POINTER(McplParticle),
                              ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int]
    return L


def test_compat_symbols_exported():
    L = ctypes.CDLL(LIB)
    for name in ("KDS_create", "KDS_open", "KDS_rand_sample2", "KDS_rand_sample", "KDS_sample2",
                 "KDS_sample", "KDS_rand_w_mean", "KDS_w_mean", "KDS_destroy", "MS_create",
                 "MS_open", "MS_rand_sample2", "MS_rand_sample", "MS_sample2", "MS_sample",
                 "MS_rand_w_mean", "MS_w_mean", "MS_destroy", "PList_create", "PList_destroy",
                 "Metric_create", "Metric_destroy", "Geom_create", "Geom_destroy", "pt2pdg",
                 "pdg2pt", "kdsource_resample_to_mcpl", "kdsource_write_mcpl", "E_perturb",
                 "Let_perturb", "Vol_perturb", "SurfXY_perturb", "Isotrop_perturb",
                 "PolarMu_perturb", "_n_metrics", "_metric_names", "_metric_perturbs"):
        assert hasattr(L, name), name
    assert ctypes.c_int.in_dll(L, "_n_metrics").value == 14
    names = (ctypes.c_char_p * 14).in_dll(L, "_metric_names")
    assert [n.decode() for n in names][:5] == ["Energy", "Lethargy", "Wavelength", "Vol", "SurfXY"]
    L.pt2pdg.argtypes = [ctypes.c_char]
    assert L.pt2pdg(b"n") == 2112 and L.pt2pdg(b"p") == 22 and L.pt2pdg(b"x") == 0


def test_kds_open_parses_reference_fixture(workdir):
    """KDS_open on the reference's own XML (no <kernel> tag, PolarMu metric)."""
    L = _lib()
    k = L.KDS_open(b"source.xml")
    assert k.contents.J == 1.0 and k.contents.kernel == b"g"
    assert k.contents.plist.contents.npts == 9 and k.contents.plist.contents.pdgcode == 2112
    assert k.contents.plist.contents.pt == b"n" and k.contents.plist.contents.x2z == 0
    assert k.contents.geom.contents.ord == 3
    assert k.contents.geom.contents.bw == pytest.approx(0.7372364669810315)
    first = k.contents.plist.contents.part.contents
    assert first.ekin == pytest.approx(0.0612344, rel=1e-5) and first.pdgcode == 2112
    L.KDS_destroy(k)


def test_kdsource_write_mcpl_c_entry(tmp_path):
    from kdsource_b200 import mcpl

    L = ctypes.CDLL(LIB)
    n = 50
    rng = np.random.default_rng(0)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    arrs = [rng.uniform(0.1, 5, n), rng.normal(size=n), rng.normal(size=n), rng.normal(size=n),
            d[:, 0].copy(), d[:, 1].copy(), d[:, 2].copy(), rng.uniform(0, 3, n),
            rng.uniform(0.5, 2, n)]
    pdg = np.full(n, 2112, dtype=np.int32)
    dp = ctypes.POINTER(ctypes.c_double)
    out = str(tmp_path / "c_out.mcpl")
    L.kdsource_write_mcpl.restype = None
    L.kdsource_write_mcpl(out.encode(), ctypes.c_uint64(n), ctypes.c_uint(2),
                          *[a.ctypes.data_as(dp) for a in arrs],
                          pdg.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None, None, None, None)
    f = mcpl.MCPLFile(out + ".gz")
    assert f.nparticles == n and f.hdr.opt_universalpdgcode == 2112 and f.hdr.particlesize == 32
    assert f.hdr.srcname == "KDSource MCPL writer"
    pb = f.read_block()
    assert np.allclose(pb.ekin, arrs[0], rtol=1e-6) and np.allclose(pb.weight, arrs[8], rtol=1e-6)
    assert np.allclose(np.stack((pb.ux, pb.uy, pb.uz), 1), d, atol=1e-6)
    # double precision + per-particle pdg + userflags
    uf = np.arange(n, dtype=np.uint32)
    out2 = str(tmp_path / "c_out2.mcpl.gz")
    L.kdsource_write_mcpl(out2.encode(), ctypes.c_uint64(n), ctypes.c_uint(1),
                          *[a.ctypes.data_as(dp) for a in arrs],
                          pdg.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None, None, None,
                          uf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    pb2 = mcpl.MCPLFile(out2).read_block()
    assert np.array_equal(pb2.x, arrs[1]) and np.array_equal(pb2.userflags, uf)
    assert np.allclose(np.stack((pb2.ux, pb2.uy, pb2.uz), 1), d, atol=1e-15)


@pytest.mark.gpu
def test_resample_fixture_like_reference_test(workdir):
    """python/kdsource/test.py:64-80: 117 particles from the embedded fixture;
    pdgcode 2112, z == 0, time == 0, weight == 1 -- through the Python API and
    through the C entry point kdsource_resample_to_mcpl."""
    from kdsource_b200 import mcpl
    from kdsource_b200.api import resample_to_mcpl

    resample_to_mcpl("source.xml", "foo_tmp_testfile.mcpl.gz", 117, rng=12345, force=True)
    L = ctypes.CDLL(LIB)
    cb = ctypes.CFUNCTYPE(ctypes.c_double)(np.random.default_rng(3).random)
    L.kdsource_resample_to_mcpl.restype = None
    L.kdsource_resample_to_mcpl(cb, b"source.xml", b"c_entry.mcpl.gz", ctypes.c_uint64(117))
    for fn in ("foo_tmp_testfile.mcpl.gz", "c_entry.mcpl.gz"):
        f = mcpl.MCPLFile(fn)
        assert f.nparticles == 117 and f.hdr.srcname == "KDSource resample"
        assert f.hdr.particlesize == 36
        pb = f.read_block()
        assert len(pb) == 117
        assert np.all(pb.pdgcode == 2112) and np.all(pb.z == 0) and np.all(pb.time == 0)
        assert np.all(pb.weight == 1.0)
        assert np.allclose(pb.ux ** 2 + pb.uy ** 2 + pb.uz ** 2, 1, atol=1e-5)
        assert np.all(pb.uz > 0) and np.all(pb.ekin > 0) and np.all(pb.ekin <= 10)
    # same seed -> same file; the callable path is honoured too
    resample_to_mcpl("source.xml", "again.mcpl.gz", 117, rng=12345)
    a = mcpl.MCPLFile("foo_tmp_testfile.mcpl.gz").read_raw(117)
    b = mcpl.MCPLFile("again.mcpl.gz").read_raw(117)
    assert a.tobytes() == b.tobytes()
    import random

    resample_to_mcpl("source.xml", "callable.mcpl.gz", 40, rng=random.Random(5).random)
    assert mcpl.MCPLFile("callable.mcpl.gz").nparticles == 40


@pytest.mark.gpu
def test_downstream_c_program(workdir, oracle):
    """Compile a host program against kdsource_compat.h, as the reference's CI does."""
    exe = str(workdir / "main_compat")
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "main_compat.c"), "-o", exe,
                    "-L", os.path.dirname(LIB), "-l:libkdsb200.so",
                    "-Wl,-rpath," + os.path.dirname(LIB)], check=True)
    n = 20000
    env = dict(os.environ, KDSB200_SEED="7")
    r = subprocess.run([exe, "source.xml", str(n)], check=True, capture_output=True, text=True,
                       env=env)
    lines = r.stdout.splitlines()
    assert "Successfully sampled particles." in lines[-1]
    assert any(ln.startswith("npts 9 J 1 kernel g order 3") for ln in lines)
    P = np.array([[float(v) for v in ln.split()[1:]] for ln in lines if ln.startswith("P ")])
    assert P.shape == (n, 10)
    assert np.all(P[:, 9] == 2112) and np.all(P[:, 8] == 1.0)
    assert np.all(P[:, 3] == 0) and np.all(P[:, 7] == 0)
    # sequential mode returns the list itself from the cursor (1000 % 9 = 1)
    from kdsource_b200 import mcpl

    pb = mcpl.MCPLFile("samples.mcpl.gz").read_block()
    S = np.array([[float(v) for v in ln.split()[1:]] for ln in lines if ln.startswith("S ")])
    assert np.allclose(S[:, 0], pb.ekin[[1, 2, 3]], rtol=1e-6)
    assert np.allclose(S[:, 3], pb.weight[[1, 2, 3]], rtol=1e-6)
    wc = float([ln for ln in lines if ln.startswith("w_crit")][0].split()[1])
    assert wc == pytest.approx(np.mean(pb.weight[np.arange(1000) % 9]), rel=1e-6)
    # distribution: same contract as the oracle with the same seed -> identical picks
    parts = np.stack((pb.ekin, pb.x, pb.y, pb.z, pb.ux, pb.uy, pb.uz, pb.time), 1)
    sc = [6.96868144e-01, 9.67440630e+00, 5.59415602e+00, 9.52596915e-03, 9.35705146e+01]
    geom = oracle.make_geom([("Lethargy", sc[:1], [10.0]),
                             ("SurfXY", sc[1:3], [-np.inf, np.inf, -np.inf, np.inf, 0.0]),
                             ("PolarMu", sc[3:5], [])])
    ref = oracle.resample(parts, pb.weight, 0.7372364669810315, geom, n, seed=7)["parts"]
    scale = np.array([1.0, 30, 30, 1, 1, 1, 1, 1])
    err = np.abs(P[:, :8] - ref) / (np.abs(ref) + scale)
    assert np.quantile(err, 0.999) < 1e-5


# ---------------------------------------------------------------------------
# host side of the C interface (no GPU): every prototype of the reference's four
# headers is exported, and the single-particle functions follow the reference
# ---------------------------------------------------------------------------
_REFERENCE_PROTOTYPES = [
    # kdsource.h.in:53-92
    "KDS_create", "KDS_open", "KDS_rand_sample2", "KDS_rand_sample", "KDS_rand_w_mean",
    "KDS_sample2", "KDS_sample", "KDS_w_mean", "KDS_destroy", "MS_create", "MS_open",
    "MS_rand_sample2", "MS_rand_sample", "MS_rand_w_mean", "MS_sample2", "MS_sample",
    "MS_w_mean", "MS_destroy",
    # geom.h:37-97
    "Metric_create", "Metric_copy", "Metric_destroy", "Geom_create", "Geom_copy", "Geom_perturb",
    "Geom_next", "Geom_destroy", "E_perturb", "Let_perturb", "wl_perturb", "t_perturb",
    "Dec_perturb", "Vol_perturb", "SurfXY_perturb", "SurfR_perturb", "SurfR2_perturb",
    "SurfCircle_perturb", "Guide_perturb", "Isotrop_perturb", "Polar_perturb", "PolarMu_perturb",
    "_n_metrics", "_metric_names", "_metric_perturbs",
    # plist.h:36-41
    "PList_create", "PList_copy", "PList_get", "PList_next", "PList_destroy",
    # utils.h:22-31 (interp and the H10 dose tables belong to the out-of-scope tally code)
    "kds_rand_norm", "kds_rand_epan", "kds_rand_box", "kds_rand_type", "traslv", "rotv",
    "pt2pdg", "pdg2pt",
    # ctypes entry points: resamplemcpl.c:17, writemcpl.c:15
    "kdsource_resample_to_mcpl", "kdsource_write_mcpl",
]
_PERTURB_FN = {"Energy": "E_perturb", "Lethargy": "Let_perturb", "Wavelength": "wl_perturb",
               "Vol": "Vol_perturb", "SurfXY": "SurfXY_perturb", "SurfR": "SurfR_perturb",
               "SurfR2": "SurfR2_perturb", "SurfCircle": "SurfCircle_perturb",
               "Guide": "Guide_perturb", "Isotrop": "Isotrop_perturb", "Polar": "Polar_perturb",
               "PolarMu": "PolarMu_perturb", "Time": "t_perturb", "Decade": "Dec_perturb"}
RNG_T = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_void_p)


def test_every_reference_prototype_is_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], check=True, capture_output=True,
                         text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    missing = [n for n in _REFERENCE_PROTOTYPES if n not in exported]
    assert not missing, missing
    # and every function the shipped header declares
    import re

    hdr = open(os.path.join(ROOT, "include", "kdsource_compat.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)                 # comments
    hdr = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", hdr)            # macros
    hdr = re.sub(r"typedef[^;{]*(\{[^}]*\})?[^;]*;", "", hdr)        # typedefs, structs
    hdr = re.sub(r"struct \w+ \{[^}]*\};", "", hdr)
    declared = set(re.findall(r"\b(\w+)\s*\([^;{()]*\)\s*;", hdr))
    declared -= {"KDS_DECL_PERTURB"}
    declared |= set(re.findall(r"KDS_DECL_PERTURB\((\w+)\)", hdr))
    assert len(declared) > 50
    assert not [n for n in declared if n not in exported]


class _Script:
    """A kds_rng_fct_t that replays a list of uniforms and counts the draws."""

    def __init__(self, u):
        self.u, self.n = list(u), 0
        self.cb = RNG_T(self._next)

    def _next(self, _state):
        v = self.u[self.n % len(self.u)]
        self.n += 1
        return v


def _part_struct(v8):
    p = McplParticle()
    p.ekin = v8[0]
    for k in range(3):
        p.position[k] = v8[1 + k]
        p.direction[k] = v8[4 + k]
    p.time, p.weight, p.pdgcode = v8[7], 1.0, 2112
    return p


def _part_vec(p):
    return np.array([p.ekin, *p.position, *p.direction, p.time])


def _metric(L, name, scaling, params):
    L.Metric_create.restype = ctypes.c_void_p
    L.Metric_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p,
                                ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    sc = (ctypes.c_double * len(scaling))(*scaling)
    pa = (ctypes.c_double * max(len(params), 1))(*params)
    fn = ctypes.cast(getattr(L, _PERTURB_FN[name]), ctypes.c_void_p)
    return L.Metric_create(len(scaling), sc, fn, len(params), pa)


@pytest.mark.parametrize("name", sorted(_PERTURB_FN))
def test_host_perturb_functions_follow_the_reference(oracle, name):
    """X_perturb called on the host, one particle at a time with the caller's rng: the
    same result and the same number of uniforms as the oracle restatement (which is
    bit-equal to the reference's compiled geom.c, tests/test_oracle.py)."""
    from test_oracle import _CASES, _random_part

    L = ctypes.CDLL(LIB)
    sc, pa = _CASES[name]
    m = _metric(L, name, sc, pa)
    fn = getattr(L, _PERTURB_FN[name])
    fn.restype = ctypes.c_int
    fn.argtypes = [RNG_T, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(McplParticle),
                   ctypes.c_double, ctypes.c_char]
    rng = np.random.default_rng(hash(name) % 977)
    for kern in "geb":
        for _ in range(60):
            part = _random_part(rng, name)
            u = rng.uniform(size=64)
            ref, used = oracle.metric_perturb_scripted(name, sc, pa, kern, 0.6, part, u)
            s = _Script(u)
            p = _part_struct(part)
            assert fn(s.cb, None, m, ctypes.byref(p), 0.6, kern.encode()) == 0
            assert s.n == used
            assert np.allclose(_part_vec(p), ref, rtol=1e-13, atol=1e-15)
    L.Metric_destroy.argtypes = [ctypes.c_void_p]
    L.Metric_destroy(m)


def test_host_deviates_translation_rotation(oracle):
    L = ctypes.CDLL(LIB)
    rng = np.random.default_rng(4)
    for fname, kern in (("kds_rand_norm", "g"), ("kds_rand_epan", "e"), ("kds_rand_box", "b")):
        fn = getattr(L, fname)
        fn.restype = ctypes.c_double
        fn.argtypes = [RNG_T, ctypes.c_void_p]
        L.kds_rand_type.restype = ctypes.c_double
        L.kds_rand_type.argtypes = [RNG_T, ctypes.c_void_p, ctypes.c_char]
        for _ in range(100):
            u = rng.uniform(size=32)
            ref, used = oracle.rand_type_scripted(kern, u)
            s = _Script(u)
            assert fn(s.cb, None) == ref and s.n == used
            s = _Script(u)
            assert L.kds_rand_type(s.cb, None, kern.encode()) == ref and s.n == used
    s = _Script([0.3])
    assert L.kds_rand_type(s.cb, None, b"x") == 0.0 and s.n == 0  # unknown kernel: no draw
    dp = ctypes.POINTER(ctypes.c_double)
    L.rotv.restype = dp
    L.rotv.argtypes = [dp, dp, ctypes.c_int]
    L.traslv.restype = dp
    L.traslv.argtypes = [dp, dp, ctypes.c_int]
    for _ in range(30):
        v, rv = rng.normal(size=3), rng.normal(size=3)
        for inv in (0, 1):
            a = v.copy()
            L.rotv(a.ctypes.data_as(dp), rv.ctypes.data_as(dp), inv)
            assert np.allclose(a, oracle.rotv(v, rv, inv), rtol=1e-14, atol=1e-15)
            b = v.copy()
            L.traslv(b.ctypes.data_as(dp), rv.ctypes.data_as(dp), inv)
            assert np.array_equal(b, v - rv if inv else v + rv)
    z = np.array([1.0, 2.0, 3.0])
    L.rotv(z.ctypes.data_as(dp), np.zeros(3).ctypes.data_as(dp), 0)  # zero rotation: untouched
    assert np.array_equal(z, [1.0, 2.0, 3.0])


def test_plist_cursor_functions(workdir):
    """PList_create / get / next / copy on the reference's 9-particle fixture
    (plist.c:15-117): cursor semantics, wrap flag, transformations."""
    from kdsource_b200 import mcpl

    pb = mcpl.MCPLFile("samples.mcpl.gz").read_block()
    L = ctypes.CDLL(LIB)
    dp = ctypes.POINTER(ctypes.c_double)
    L.PList_create.restype = ctypes.POINTER(PListS)
    L.PList_create.argtypes = [ctypes.c_char, ctypes.c_char_p, dp, dp, ctypes.c_int]
    L.PList_copy.restype = ctypes.POINTER(PListS)
    L.PList_copy.argtypes = [ctypes.POINTER(PListS)]
    L.PList_get.argtypes = [ctypes.POINTER(PListS), ctypes.POINTER(McplParticle)]
    L.PList_next.argtypes = [ctypes.POINTER(PListS), ctypes.c_int]
    L.PList_destroy.argtypes = [ctypes.POINTER(PListS)]
    tr = np.array([1.0, -2.0, 0.5])
    rot = np.array([0.0, 0.0, np.pi / 2])
    pl = L.PList_create(b"n", b"samples.mcpl.gz", tr.ctypes.data_as(dp), rot.ctypes.data_as(dp), 1)
    assert pl.contents.npts == 9 and pl.contents.part.contents.ekin == pytest.approx(pb.ekin[0])
    part = McplParticle()
    seen, wraps = [], []
    for i in range(12):
        assert L.PList_get(pl, ctypes.byref(part)) == 0
        seen.append(part.ekin)
        if i == 0:
            # +trasl, rotate 90 degrees about z, then (x,y,z) -> (y,z,x)  (plist.c:69-89)
            x, y, z = pb.x[0] + 1.0, pb.y[0] - 2.0, pb.z[0] + 0.5
            assert np.allclose(list(part.position), [x, z, -y], atol=1e-12)
            assert np.allclose(list(part.direction), [pb.ux[0], pb.uz[0], -pb.uy[0]], atol=1e-12)
        wraps.append(L.PList_next(pl, 1))
    assert np.allclose(seen, pb.ekin[np.arange(12) % 9], rtol=1e-15)
    assert wraps == [0] * 8 + [1] + [0] * 3  # the 9th advance rewinds (plist.c:96-100)
    cp = L.PList_copy(pl)  # same file, cursor of its own at the first particle
    assert cp.contents.part.contents.ekin == pytest.approx(pb.ekin[0])
    L.PList_next(cp, 1)
    assert cp.contents.part.contents.ekin == pytest.approx(pb.ekin[1])
    assert pl.contents.part.contents.ekin == pytest.approx(pb.ekin[12 % 9])
    L.PList_destroy(pl)
    assert cp.contents.part.contents.ekin == pytest.approx(pb.ekin[1])  # list survives the original
    L.PList_destroy(cp)


def test_geom_next_copy_and_host_perturb(tmp_path, oracle):
    """Geom_create with a bandwidth file, Geom_next (geom.c:160-177), Geom_copy /
    Metric_copy, and Geom_perturb (geom.c:139-158: -trasl, rot^-1, metrics, rot, +trasl)
    against the same steps composed from the oracle's pieces."""
    L = ctypes.CDLL(LIB)
    bwf = tmp_path / "bws"
    np.array([0.5, 0.25, 0.75], dtype=np.float32).tofile(bwf)
    ms = [_metric(L, "Lethargy", [0.7], [10.0]),
          _metric(L, "Vol", [1.0, 2.0, 0.5], [-50, 50, -30, 30, -10, 10]),
          _metric(L, "Isotrop", [0.2, 0.2, 0.2], [0, 0, 1])]
    arr = (ctypes.c_void_p * 3)(*ms)
    tr = np.array([0.3, -0.2, 0.1])
    rv = np.array([0.2, -0.1, 0.4])
    dp = ctypes.POINTER(ctypes.c_double)

    class GeomFull(ctypes.Structure):
        _fields_ = GeomS._fields_ + [("trasl", dp), ("rot", dp),
                                     ("bws", ctypes.POINTER(ctypes.c_float)),
                                     ("nbws", ctypes.c_longlong), ("bwpos", ctypes.c_longlong)]

    L.Geom_create.restype = ctypes.POINTER(GeomFull)
    L.Geom_create.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_char_p,
                              ctypes.c_char, dp, dp]
    L.Geom_copy.restype = ctypes.POINTER(GeomFull)
    L.Geom_copy.argtypes = [ctypes.POINTER(GeomFull)]
    L.Geom_next.argtypes = [ctypes.POINTER(GeomFull), ctypes.c_int]
    L.Geom_perturb.argtypes = [RNG_T, ctypes.c_void_p, ctypes.POINTER(GeomFull),
                               ctypes.POINTER(McplParticle)]
    L.Geom_destroy.argtypes = [ctypes.POINTER(GeomFull)]
    g = L.Geom_create(3, arr, 9.0, str(bwf).encode(), b"g", tr.ctypes.data_as(dp),
                      rv.ctypes.data_as(dp))
    assert g.contents.bw == 0.5  # first value read at creation (geom.c:96)
    assert [L.Geom_next(g, 1) for _ in range(3)] == [0, 0, 1]  # third call rewinds
    assert g.contents.bw == 0.5
    assert L.Geom_next(g, 1) == 0 and g.contents.bw == 0.25
    cp = L.Geom_copy(g)
    assert cp.contents.ord == 3 and cp.contents.kernel == b"g" and cp.contents.bw == 0.5
    assert np.allclose([cp.contents.trasl[k] for k in range(3)], tr)
    # constant bandwidth: Geom_next is a no-op
    g2 = L.Geom_create(0, None, 0.4, None, b"e", None, None)
    assert L.Geom_next(g2, 0) == 0 and g2.contents.bw == 0.4
    L.Geom_destroy(g2)
    # Geom_perturb on the host
    rng = np.random.default_rng(12)
    names = [("Lethargy", [0.7], [10.0]), ("Vol", [1.0, 2.0, 0.5], [-50, 50, -30, 30, -10, 10]),
             ("Isotrop", [0.2, 0.2, 0.2], [0, 0, 1])]
    for _ in range(40):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        part = np.array([rng.uniform(0.01, 9), *rng.uniform(-3, 3, 3), *d, 0.7])
        u = rng.uniform(size=96)
        s = _Script(u)
        p = _part_struct(part)
        assert L.Geom_perturb(s.cb, None, g, ctypes.byref(p)) == 0
        ref = part.copy()
        ref[1:4] = oracle.rotv(ref[1:4] - tr, rv, 1)
        ref[4:7] = oracle.rotv(ref[4:7], rv, 1)
        used = 0
        for nm, sc, pa in names:
            ref, k = oracle.metric_perturb_scripted(nm, sc, pa, "g", g.contents.bw, ref, u[used:])
            used += k
        ref[1:4] = oracle.rotv(ref[1:4], rv, 0) + tr
        ref[4:7] = oracle.rotv(ref[4:7], rv, 0)
        assert s.n == used
        assert np.allclose(_part_vec(p), ref, rtol=1e-12, atol=1e-13)
    L.Geom_destroy(cp)
    L.Geom_destroy(g)


@pytest.mark.gpu
def test_multisource_and_bias_c_program(workdir, oracle):
    """MS_open -> MS_w_mean -> MS_sample2 (kdsource.c:375-473) from a gcc-compiled host
    program: source choice follows ws, weights carry (J_i/J)/(ws_i/sum ws), each
    source's particles follow its own density (KS against the oracle sampler); then
    KDS_sample2 with a WeightFun bias (kdsource.c:294-325) and a host-side list walk."""
    from scipy import stats

    import kdsource_b200.api as kds
    from kdsource_b200 import geom as kgeom
    from kdsource_b200 import mcpl

    # second source: photons, its own list, J = 3
    parts, w = oracle.synthetic_particles(4000, seed=31, wseed=32)
    mcpl.write_mcpl_file("photons.mcpl.gz", ekin=parts[:, 0], x=parts[:, 1], y=parts[:, 2],
                         z=parts[:, 3], ux=parts[:, 4], uy=parts[:, 5], uz=parts[:, 6],
                         time=parts[:, 7], weight=w, pdgcode=22, srcname="t")
    g = kgeom.Geometry([kgeom.Lethargy(10), kgeom.SurfXY(), kgeom.Isotrop(keep_zdir=True)])
    src = kds.KDSource(kds.PList("photons.mcpl.gz", pt="p"), g, bw="silv", J=3.0)
    src.fit()
    src.save("photons.xml", adjust_N=False)
    exe = str(workdir / "main_ms")
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "main_ms.c"), "-o", exe,
                    "-L", os.path.dirname(LIB), "-l:libkdsb200.so",
                    "-Wl,-rpath," + os.path.dirname(LIB)], check=True)
    n = 30000
    srand_env = dict(os.environ, KDSB200_SEED="11")
    r = subprocess.run([exe, "source.xml", "photons.xml", "1.0", "2.5", str(n)], check=True,
                       capture_output=True, text=True, env=srand_env)
    lines = r.stdout.splitlines()
    assert "Successfully sampled particles." in lines[-1]
    head = [ln for ln in lines if ln.startswith("MS ")][0].split()
    assert int(head[2]) == 2 and float(head[4]) == 4.0 and float(head[6]) == 1.0 \
        and float(head[7]) == 3.5
    pb = mcpl.MCPLFile("samples.mcpl.gz").read_block()
    wc = float([ln for ln in lines if ln.startswith("w_crit")][0].split()[1])
    wm0 = np.mean(pb.weight[np.arange(1000) % 9])
    wm1 = np.mean(w[:1000])
    assert wc == pytest.approx((1.0 * wm0 + 2.5 * wm1) / 3.5, rel=1e-6)  # kdsource.c:454-462
    P = np.array([[float(v) for v in ln.split()[1:]] for ln in lines if ln.startswith("P ")])
    assert P.shape == (n, 10)
    is_n, is_p = P[:, 9] == 2112, P[:, 9] == 22
    assert np.all(is_n | is_p)
    # source choice ~ ws / sum ws (binomial, 5 sigma)
    frac = is_p.mean()
    assert abs(frac - 2.5 / 3.5) < 5 * np.sqrt((2.5 / 3.5) * (1 / 3.5) / n)
    # weight correction (J_i / J) / (ws_i / sum ws), kdsource.c:431-436
    assert np.allclose(P[is_n, 8], (1.0 / 4.0) / (1.0 / 3.5), rtol=1e-12)
    assert np.allclose(P[is_p, 8], (3.0 / 4.0) / (2.5 / 3.5), rtol=1e-12)
    # each source's particles follow that source: KS against the oracle sampler
    sc = src.scaling
    geom_p = oracle.make_geom([("Lethargy", sc[:1], [10.0]),
                               ("SurfXY", sc[1:3], [-np.inf, np.inf, -np.inf, np.inf, 0.0]),
                               ("Isotrop", sc[3:6], [0, 0, 1])])
    ref = oracle.resample(parts, w, float(src.kde.bw), geom_p, 20000, seed=99)["parts"]
    for col in (0, 1, 2, 6):
        assert stats.ks_2samp(P[is_p, col], ref[:, col]).pvalue > 1e-3, col
    scn = [6.96868144e-01, 9.67440630e+00, 5.59415602e+00, 9.52596915e-03, 9.35705146e+01]
    geom_n = oracle.make_geom([("Lethargy", scn[:1], [10.0]),
                               ("SurfXY", scn[1:3], [-np.inf, np.inf, -np.inf, np.inf, 0.0]),
                               ("PolarMu", scn[3:5], [])])
    pn = np.stack((pb.ekin, pb.x, pb.y, pb.z, pb.ux, pb.uy, pb.uz, pb.time), 1)
    refn = oracle.resample(pn, pb.weight, 0.7372364669810315, geom_n, 20000, seed=98)["parts"]
    for col in (0, 1, 2, 6):
        assert stats.ks_2samp(P[is_n, col], refn[:, col]).pvalue > 1e-3, col
    # bias: picks ~ w * bias, weight = 1 / bias
    bias = np.where(pb.ekin > 0.05, 3.0, 0.5)
    wb = float([ln for ln in lines if ln.startswith("wb_mean")][0].split()[1])
    cur = (1000 + np.arange(9)) % 9  # the cursor after MS_w_mean(…, 1000, …)
    assert wb == pytest.approx(np.mean(pb.weight[cur] * bias[cur]), rel=1e-6)
    B = np.array([[float(v) for v in ln.split()[1:]] for ln in lines if ln.startswith("B ")])
    assert B.shape == (n, 2)
    idx = np.argmin(np.abs(B[:, :1] - pb.ekin[None, :]), axis=1)
    assert np.allclose(B[:, 0], pb.ekin[idx], rtol=1e-6)  # unperturbed list entries
    assert np.allclose(B[:, 1], 1.0 / bias[idx], rtol=1e-12)
    expect = pb.weight * bias / np.sum(pb.weight * bias)
    counts = np.bincount(idx, minlength=9)
    assert stats.chisquare(counts, expect * n).pvalue > 1e-3
    # host walk: the list in order from the cursor PList_create left (entry 0), perturbed
    # energies differ, wrap flags are 0 inside the list
    W = np.array([[float(v) for v in ln.split()[1:]] for ln in lines if ln.startswith("W ")])
    assert np.allclose(W[:, 0], pb.ekin[:5], rtol=1e-6) and np.all(W[:, 1] != W[:, 0])
    assert np.all(W[:, 1] <= 10.0) and np.all(W[:, 3] == 0) and np.all(W[:, 4] == 0)

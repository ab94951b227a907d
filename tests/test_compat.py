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

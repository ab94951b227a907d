#!/usr/bin/env python
"""Benchmark of the KDSource kernel-density hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   (CPU reference arm)

Headline workload (BASELINE.json configs[1]): MLCV bandwidth optimisation on a
1e6-particle 6-D (E,x,y,u,v,w) list, 32-point bandwidth grid, 10 folds.  One
"step" is one full grid search = 32 * sum_f n_f (N - n_f) = 2.88e13 Gaussian
kernel evaluations.  `value` = kernel evaluations / s over all ranks with the
inputs resident in HBM; `e2e` is the same search through the public API
(`kdsource_b200.kde.mlcv_scores`) from host numpy arrays, host<->device copies
included.  With N > 1 ranks the test rows are sharded (strong scaling: the
configuration's list size is fixed) and the G partial scores are all-reduced.

Secondary numbers for the other sub-paths (evaluate, kNN, resample) ride in
the `sub` object of the same JSON line, each with its own roofline.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "kde_kernel_pair_evals_per_s"
UNIT = "evals/s"
N_MLCV = 1_000_000
G_MLCV = 32
K_MLCV = 10
D_MLCV = 6
CPU_SAMPLE_N = 24_000  # bounded CPU sample of the same workload


def synthetic_scaled(N, seed=12345, wseed=777):
    """Synthetic correlated neutron list in scaled KDE variables."""
    from kdsource_b200 import geom as kgeom

    rng = np.random.default_rng(seed)
    n1 = N // 2
    u = np.concatenate((rng.normal(5, 1, n1), rng.normal(9, 1, N - n1)))
    x = np.concatenate((rng.normal(10, 10, n1), rng.normal(-10, 10, N - n1)))
    y = rng.normal(0, 10, N)
    mu = np.sqrt(rng.uniform(0, 1, N))
    phi = rng.uniform(-np.pi, np.pi, N)
    dxy = np.sqrt(1 - mu ** 2)
    parts = np.stack((10 * np.exp(-u), x, y, np.zeros(N), dxy * np.cos(phi), dxy * np.sin(phi),
                      mu, np.zeros(N)), axis=1)
    parts = parts[rng.permutation(N)]
    w = np.maximum(np.random.default_rng(wseed).normal(1, 0.1, N), 1e-99)
    g = kgeom.Geometry([kgeom.Lethargy(10), kgeom.SurfXY(), kgeom.Isotrop(keep_zdir=True)])
    vecs = g.transform(parts.copy())
    scaling = g.std(vecs=vecs, weights=w)
    return parts, w, g, scaling, vecs / scaling


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)),
                "power_w_max": float(np.max(pw)), "samples": len(sm),
                "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


def mlcv_config(n_mlcv):
    """The `config` object of the headline line -- the same for both arms."""
    return {"workload": "mlcv_1e6x6d_g32_k10" if n_mlcv == N_MLCV else
            "mlcv_{}x6d_g32_k10".format(n_mlcv),
            "N": n_mlcv, "d": D_MLCV, "G": G_MLCV, "n_splits": K_MLCV,
            "l2": "flushed between timed steps (256 MiB write, untimed)",
            "sharding": "test rows over ranks, each rank uploads its rows and the list is "
                        "all-gathered over NVLink, device reduction + NCCL all-reduce of the G "
                        "partial log-likelihoods inside the timed step"}


def cpu_mlcv_baseline(data, w, bw, grid, nthreads=0):
    """Oracle port (C + OpenMP, all host threads) on a bounded sample."""
    from oracle import kds_oracle as O

    n = min(CPU_SAMPLE_N, len(data))
    fb = O.kfold_bounds(n, K_MLCV)
    nf = np.diff(fb)
    evals = float(np.sum(nf * (n - nf))) * len(grid)
    t0 = time.perf_counter()
    # explicit thread count: torchrun exports OMP_NUM_THREADS=1
    O.mlcv_scores_c(data[:n], w[:n], bw, grid, K_MLCV, nthreads=nthreads or (os.cpu_count() or 1))
    dt = time.perf_counter() - t0
    return evals / dt, dt, n


def cpu_mlcv_tree_baseline(n=100_000):
    """The reference's ALGORITHM as it runs on a CPU (kde.py:134-151, 204-211): one joblib
    process per grid point, each building a kd-tree over the training folds and summing the
    truncated Gaussian over an approximate ball query per test point
    (oracle.mlcv_scores_tree) -- at N = 1e5 on all cores, bounded to ONE
    of the ten folds per grid point (a fold costs the same as any other), reported in
    dense-equivalent evaluations/s.  In six dimensions the ball query visits most of the
    tree, so this is slower than the dense OpenMP port, not faster."""
    from kdsource_b200 import kde
    from oracle import kds_oracle as O

    cores = os.cpu_count() or 1
    _, w, _, _, data = synthetic_scaled(n)
    bw = kde.bw_silv(D_MLCV, np.sum(w) ** 2 / np.sum(w ** 2))
    grid = np.logspace(-0.3, 0.3, G_MLCV)
    g = grid[::max(1, len(grid) // min(cores, len(grid)))][:cores]
    fb = O.kfold_bounds(n, K_MLCV)
    n0 = int(fb[1] - fb[0])
    evals = float(n0) * (n - n0) * len(g)
    from joblib import Parallel, delayed

    def one_fold(gf):
        train = np.r_[fb[1]:n]
        return O.kde_sum_tree(data[train], w[train], gf * bw, data[:fb[1]])

    t0 = time.perf_counter()
    Parallel(n_jobs=-1)(delayed(one_fold)(gf) for gf in g)
    dt = time.perf_counter() - t0
    return {"value": evals / dt, "unit": UNIT + " (dense-equivalent)", "cores": cores,
            "kind": "port", "N": n,
            "sample": "cKDTree ball query + numpy, fold 0 of 10 for {} grid points in {} joblib "
                      "processes, N = {} ({:.1f} s)".format(len(g), cores, n, dt)}


def run_reference(args):
    """`--impl reference`: the CPU implementation of the path on the host cores.
    The reference's MLCV is Python + KDEpy + joblib, none of which can run on the
    GPU box (KDEpy is not installable; /root/reference does not travel), so this
    arm times restatements of it.  `value` is the FASTER of the two: the dense O(N^2)
    sum in C + OpenMP on all cores (size-independent rate, timed on the first 24 000
    particles each step).  `cpu_baseline.reference_algorithm` is the reference's own
    algorithm (kd-tree ball queries, one joblib process per grid point) timed once at
    N = 1e5 on all cores -- in six dimensions it is the slower one."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from kdsource_b200 import kde

    cores = os.cpu_count() or 1
    _, w, _, _, data = synthetic_scaled(CPU_SAMPLE_N)
    bw = kde.bw_silv(D_MLCV, np.sum(w) ** 2 / np.sum(w ** 2))
    grid = np.logspace(-0.3, 0.3, G_MLCV)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, n = cpu_mlcv_baseline(data, w, bw, grid)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = ("dense fp64 sum (C + OpenMP, {} threads) over the first {} particles of the 1e6 "
              "list, G={}, K={}, per step".format(cores, CPU_SAMPLE_N, G_MLCV, K_MLCV))
    try:
        tree = cpu_mlcv_tree_baseline()
    except Exception as e:  # joblib workers unavailable: keep the dense baseline
        tree = {"unavailable": str(e)[:200]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": mlcv_config(args.n or N_MLCV),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample, "reference_algorithm": tree},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def time_events(t, fn, reps):
    """Per-call CUDA-event times (ms) of fn() on the current stream."""
    out = []
    for _ in range(reps):
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


def sub_benchmarks(t, peaks, quick, cpu=True):
    """Secondary sub-paths on one GPU: evaluate, kNN, resample; each with the device
    rate + roofline, the end-to-end rate through the host API and (cpu=True) the CPU
    reference of that sub-path on a bounded sample."""
    cores = os.cpu_count() or 1
    import ctypes

    from kdsource_b200 import _lib, kde
    from kdsource_b200.sampler import DeviceSampler, make_geom_struct
    from kdsource_b200.treekde import TreeKDE

    L = _lib.load()
    sub = {}
    flush = t.empty(256 << 20, dtype=t.uint8, device="cuda")
    # --- evaluate: N=1e5 sources (config 1), M=1e6 queries -> 1e11 pairs per launch
    N, M = 100_000, (200_000 if quick else 1_000_000)
    parts, w, g, scaling, data = synthetic_scaled(N)
    _, _, _, _, q = synthetic_scaled(M, seed=54321, wseed=1)
    bw = kde.bw_silv(D_MLCV, np.sum(w) ** 2 / np.sum(w ** 2))

    def eval_leg(kd, d_qq, Mq, lane_ops, reps=5):
        """Device rate of one TreeKDE configuration: dense-equivalent and executed pairs/s
        (they differ when tiles are culled), roofline on the executed pairs."""
        dv = kd._device_state()
        for _ in range(3):
            kd.evaluate_device(d_qq, Mq, dv)
        ts = []
        for _ in range(reps):
            flush.zero_()
            ts += time_events(t, lambda: kd.evaluate_device(d_qq, Mq, dv), 1)
        ms_ = float(np.median(ts))
        dense = float(dv["N"]) * Mq
        executed = dense
        if dv["dtype"] == "float32":
            executed = float(kd.last_tiles_visited.item()) * 512.0 * 1024.0
        return {"kernel": dv["dtype"], "cutoff": kd.cutoff, "ms": ms_,
                "pairs_per_s": dense / (ms_ * 1e-3),
                "executed_pairs_per_s": executed / (ms_ * 1e-3),
                "visited_fraction": executed / dense,
                "roofline": {"bound": "fp32", "achieved": executed * lane_ops / (ms_ * 1e-3) / 1e9,
                             "peak": peaks["ffma2"] / 1e9, "unit": "Glane-op/s",
                             "frac": executed * lane_ops / (ms_ * 1e-3) / peaks["ffma2"],
                             "per_unit": "{} FP32 lane-ops + 1 MUFU per executed pair (d=6); "
                                         "includes ordering and packing the queries".format(lane_ops),
                             "traffic": None}}

    d_q = _lib.to_dev(q)
    # the default path of the public API (dtype="auto"): ordered tiles, tile-local coordinates
    k = TreeKDE(bw=bw).fit(data, w)
    leg = eval_leg(k, d_q, M, 13)
    sub["evaluate"] = {
        "workload": "N=1e5 sources x M={:g} queries, d=6, Silverman bw, dtype='auto' (the "
                    "default)".format(M), **leg}
    ms = leg["ms"]
    # the other kernels on the same shape
    for name, lane in (("float32x2", 25), ("float32_dense", 13)):
        kk = TreeKDE(bw=bw, dtype=name).fit(data, w)
        sub["evaluate"][name] = eval_leg(kk, d_q, M, lane, reps=3)
        del kk
    if not quick:
        # a list of 1e6 (Silverman 0.23 of the spread, where the unordered fp32 kernel is past
        # its 1e-5 domain): the default path, exact and with KDEpy's hard support radius
        N2 = 1_000_000
        _, w2, _, _, data2 = synthetic_scaled(N2)
        bw2 = kde.bw_silv(D_MLCV, np.sum(w2) ** 2 / np.sum(w2 ** 2))
        for name, cut in (("auto_N1e6", None), ("auto_N1e6_cutoff_treekde", "treekde")):
            kk = TreeKDE(bw=bw2, cutoff=cut).fit(data2, w2)
            sub["evaluate"][name] = eval_leg(kk, d_q, M, 13, reps=3)
            del kk
        del data2, w2
    cpu_jobs = []  # the CPU references run after every GPU measurement (their OpenMP / joblib
    #                workers would otherwise compete with the host side of the e2e runs)

    def cpu_evaluate():
        from oracle import kds_oracle as O

        mq = 20_000  # bounded sample: 1e5 sources x 2e4 queries = 2e9 pairs
        t0 = time.perf_counter()
        O.kde_sum_c(data, w, bw, q[:mq], nthreads=cores)
        dt = time.perf_counter() - t0
        sub["evaluate"]["cpu_baseline"] = {
            "value": N * mq / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "dense fp64 sum, 1e5 sources x {} queries ({:.1f} s)".format(mq, dt)}

    cpu_jobs.append(cpu_evaluate)

    def host_api_seconds(fn, reps=3):
        """Median wall time of fn() through the host API after one warm-up call (the first
        call allocates the pinned staging buffers)."""
        import gc

        gc.collect()
        t.cuda.empty_cache()  # start from the allocator state a fresh process would have
        fn()
        ts = []
        for _ in range(reps):
            t.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            t.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts))

    # e2e through KDSource-level API: host (M,6) fp64 in, densities out
    Me = M
    dt = host_api_seconds(lambda: k.evaluate(q[:Me]))
    sub["evaluate"]["e2e"] = {"value": N * Me / dt, "unit": "pairs/s",
                              "h2d_bytes_per_step": Me * 6 * 8, "d2h_bytes_per_step": Me * 8,
                              "api": "TreeKDE.evaluate(host array), M={}; median of 3 calls "
                                     "after one warm-up".format(Me)}
    # --- kNN: k=10, batches of 1e4 (reference default), fp64 exact mode
    Nk = 200_000 if quick else 1_000_000
    _, wk, _, _, dk = synthetic_scaled(Nk, seed=99)
    off = kde.array_split_bounds(Nk, int(round(Nk / 10000)))
    d_data = _lib.to_dev(dk)
    d_off = _lib.to_dev(off, np.int64)
    res = {}
    for prec, use64 in (("f64", 1), ("f32", 0)):
        scratch = _lib.empty(Nk * 6, t.float64 if use64 else t.float32)
        d_kth = _lib.empty(Nk, t.float64)

        def kn():
            _lib.check(L.kdsb200_knn(_lib.ptr(d_data), Nk, 6, _lib.ptr(d_off), len(off) - 1,
                                     int(np.max(np.diff(off))), 11, use64, _lib.ptr(scratch),
                                     _lib.ptr(d_kth), None, None, _lib.stream()))

        for _ in range(2):
            kn()
        ms = float(np.median(time_events(t, kn, 3)))
        pk = float(np.sum(np.diff(off).astype(np.float64) ** 2))
        # distance arithmetic only (the top-k selection is amortised over the pairs):
        # fp64 mode 3d-1 = 17 DADD/DMUL per pair (no contraction: bit-exact with numpy),
        # fp32 mode 2d = 12 packed lane-ops per pair
        ops, peak, unit = ((17.0, peaks["dfma"], "fp64 lane-op/s") if use64
                           else (12.0, peaks["ffma2"], "fp32 lane-op/s"))
        res[prec] = {"pairs_per_s": pk / (ms * 1e-3), "ms": ms,
                     "roofline": {"bound": "fp64" if use64 else "fp32",
                                  "achieved": pk * ops / (ms * 1e-3) / 1e9, "peak": peak / 1e9,
                                  "unit": "G" + unit, "frac": pk * ops / (ms * 1e-3) / peak,
                                  "traffic": None}}
    # fp64 pipe: d subtractions + d multiplies + (d-1) additions per pair, no FMA
    # contraction (bit-exact with numpy/sklearn): 3d-1 DP ops per pair
    sub["knn"] = {"workload": "N={:g}, k=10, batch_size=1e4, d=6".format(Nk), **res,
                  "per_unit": "17 fp64 ops (f64 mode) / 12 fp32 lane-ops (f32 mode) per pair"}
    dt = host_api_seconds(lambda: kde.knn_batched(dk, 11, off))
    sub["knn"]["e2e"] = {"value": pk / dt, "unit": "pairs/s", "h2d_bytes_per_step": Nk * 6 * 8,
                         "d2h_bytes_per_step": Nk * 8,
                         "api": "kde.knn_batched(host array); median of 3 calls after one warm-up"}

    def cpu_knn():
        from sklearn.neighbors import NearestNeighbors

        nb_s = 20  # bounded sample: the first 20 batches, sklearn as kde.py:93-97 calls it
        t0 = time.perf_counter()
        for b in range(nb_s):
            blk = dk[off[b]:off[b + 1]]
            NearestNeighbors(n_neighbors=11, n_jobs=-1).fit(blk).kneighbors(blk)
        dt = time.perf_counter() - t0
        sub["knn"]["cpu_baseline"] = {
            "value": float(np.sum(np.diff(off[:nb_s + 1]).astype(np.float64) ** 2)) / dt,
            "unit": "pairs/s (dense-equivalent)", "cores": cores, "kind": "reference",
            "sample": "sklearn NearestNeighbors(k+1, n_jobs=-1) kd-tree on the first {} batches "
                      "({:.1f} s)".format(nb_s, dt)}

    cpu_jobs.append(cpu_knn)
    # --- resample: N=1e7 adaptive-bw sources (config 4), 1e8 particles per step
    Ns = 1_000_000 if quick else 10_000_000
    nout = 10_000_000 if quick else 100_000_000
    ps, ws_, gs_, sc_, _ = synthetic_scaled(Ns, seed=7)
    bws = np.random.default_rng(3).uniform(0.2, 0.5, Ns).astype(np.float32)
    smp = DeviceSampler(ps, ws_, bws, make_geom_struct(gs_, sc_, "g"))
    out = _lib.empty(nout * 36, t.uint8)

    def rs():
        smp.sample_device(nout, seed=1, first=0, out=out)

    for _ in range(2):
        rs()
    ms = float(np.median(time_events(t, rs, 3)))
    bytes_alg = 76.0 * nout
    traffic, traffic_pp = None, None
    tp = os.path.join(ROOT, "profiles", "traffic_resample.json")
    if os.path.exists(tp) and not quick:
        with open(tp) as fh:
            tj = json.load(fh)
        traffic, traffic_pp = tj.get("dram_bytes_per_launch"), tj.get("dram_bytes_per_particle")
    sub["resample"] = {
        "workload": "N={:g} sources, adaptive bw, {:g} particles -> MCPL-3 records".format(Ns, nout),
        "particles_per_s": nout / (ms * 1e-3), "ms": ms,
        "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": bytes_alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "per_unit": "76 B per particle (40 read + 36 written)", "traffic": traffic,
                     "traffic_note": "ncu: {} B of DRAM traffic per particle -- a random pick "
                                     "costs a whole 128-byte HBM line; at the measured rate that "
                                     "is {:.0f} GB/s".format(traffic_pp, (traffic_pp or 0) * nout /
                                                            (ms * 1e-3) / 1e9)}}
    # device gzip of those records (gz_size + gz_scan + gz_encode kernels): algorithmic bytes
    # = 2 reads of the records + the compressed stream written
    nb = nout * 36
    member = int(L.kdsb200_gz_member_bytes())
    hist = _lib.empty(258, t.int64)
    _lib.check(L.kdsb200_gz_histogram(_lib.ptr(out), min(nb, 64 << 20) // 36 * 36, member, 36, 12,
                                      _lib.ptr(hist), _lib.stream()))
    hh = hist.cpu().numpy().astype(np.uint64)
    table = _lib.GzTable()
    _lib.check(L.kdsb200_gz_build_table(hh.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), 36, 12,
                                        ctypes.byref(table)))
    cap = (nb + nb // 4 + (1 << 20)) // 4 * 4
    gz_out = _lib.empty(cap, t.uint8)
    gz_scr = _lib.empty(int(L.kdsb200_gz_scratch_bytes(nb, member)), t.uint8)
    gz_tot = _lib.empty(1, t.int64)

    def gz():
        _lib.check(L.kdsb200_gz_compress(_lib.ptr(out), nb, member, ctypes.byref(table),
                                         _lib.ptr(gz_scr), _lib.ptr(gz_out), cap, _lib.ptr(gz_tot),
                                         _lib.stream()))

    for _ in range(2):
        gz()
    ms = float(np.median(time_events(t, gz, 3)))
    zbytes = int(gz_tot.cpu()[0])
    sub["resample"]["gzip_device"] = {
        "particles_per_s": nout / (ms * 1e-3), "ms": ms, "ratio": zbytes / nb,
        "roofline": {"bound": "hbm", "achieved": (2.0 * nb + zbytes) / (ms * 1e-3) / 1e9,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": (2.0 * nb + zbytes) / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "per_unit": "72 B read (two passes) + {:.1f} B written per particle; "
                                 "the encoder zeroes its output itself".format(zbytes / nout),
                     "traffic": None}}
    del gz_out, gz_scr
    sub["resample"]["e2e"] = resample_e2e(ps, ws_, bws, gs_, sc_, quick)
    if cpu:
        for job in cpu_jobs:
            job()
        sub["resample"]["cpu_baseline"] = cpu_resample_baseline(ps, ws_, bws, sc_, cores)
    return sub


def cpu_resample_baseline(ps, ws_, bws, sc_, cores):
    """The reference's own C sampler (clib/src compiled in place -> oracle/_ref) over
    an in-memory list with an inlined xorshift generator, single thread (the
    reference has no threaded sampler); oracle port if _ref was not built."""
    import tempfile

    from oracle import kds_oracle as O

    n = 10_000_000
    Nl = min(len(ps), 1_000_000)
    metrics = [("Lethargy", sc_[:1], [10.0]),
               ("SurfXY", sc_[1:3], [-np.inf, np.inf, -np.inf, np.inf, 0.0]),
               ("Isotrop", sc_[3:6], [0, 0, 1])]
    if O.ref_lib() is not None:
        with tempfile.NamedTemporaryFile(suffix="_bws") as fh:
            bws[:Nl].astype(np.float32).tofile(fh)
            fh.flush()
            t0 = time.perf_counter()
            O.ref_sampler(ps[:Nl], ws_[:Nl], 2112, metrics, "g", 1.0, n, seed=5, bwfile=fh.name)
            dt = time.perf_counter() - t0
        kind = "reference"
        what = "KDS_rand_sample2 loop of the reference C library"
        # the same loop as `kdsource-resample` really drives it: one ctypes callback into
        # random.Random(seed).random per uniform (python/kdsource/resample.py:33-43)
        ncli = 1_000_000
        with tempfile.NamedTemporaryFile(suffix="_bws") as fh:
            bws[:Nl].astype(np.float32).tofile(fh)
            fh.flush()
            t0 = time.perf_counter()
            O.ref_sampler_cli(ps[:Nl], ws_[:Nl], 2112, metrics, "g", 1.0, ncli, seed=5,
                              bwfile=fh.name)
            dtc = time.perf_counter() - t0
        cli = {"value": ncli / dtc, "unit": "particles/s", "cores": 1, "kind": "reference",
               "sample": "the same C loop with the CLI's per-uniform Python callback "
                         "(random.Random(seed).random through ctypes), {} particles "
                         "({:.1f} s)".format(ncli, dtc)}
    else:
        t0 = time.perf_counter()
        O.resample(ps[:Nl], ws_[:Nl], bws[:Nl], O.make_geom(metrics), n, seed=5)
        dt = time.perf_counter() - t0
        kind = "port"
        what = "oracle C restatement"
        cli = None
    return {"value": n / dt, "unit": "particles/s", "cores": 1, "kind": kind,
            "sample": "{}: {} particles from a {}-particle in-memory list, inlined xorshift "
                      "generator, no file output ({:.1f} s)".format(what, n, Nl, dt),
            "cli_faithful": cli}


def resample_e2e(ps, ws_, bws, gs_, sc_, quick):
    """kdsource-resample end to end: XML + MCPL list on disk -> resample_to_mcpl ->
    gzip-compressed MCPL-3 file (the call of python/kdsource/resample.py:2-49)."""
    import shutil
    import tempfile

    import kdsource_b200.api as kds
    from kdsource_b200 import mcpl

    Nl = min(len(ps), 1_000_000)
    nout = 2_000_000 if quick else 100_000_000
    nout_host = 2_000_000 if quick else 20_000_000  # host zlib comparison: bounded
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="kdsb_bench_", dir=base)
    cwd = os.getcwd()
    try:
        os.chdir(tmp)
        pp = ps[:Nl]
        mcpl.write_mcpl_file("list.mcpl.gz", ekin=pp[:, 0], x=pp[:, 1], y=pp[:, 2], z=pp[:, 3],
                             ux=pp[:, 4], uy=pp[:, 5], uz=pp[:, 6], time=pp[:, 7],
                             weight=ws_[:Nl], pdgcode=2112, srcname="synthetic")
        src = kds.KDSource(kds.PList("list.mcpl.gz"), gs_, bw=bws[:Nl].astype(np.float64))
        src.fit(scaling=sc_)
        xml = src.save("src.xml", adjust_N=False)
        res = {}
        # warm-up: pinned output buffers, the writer's worker threads, file-system state
        kds.resample_to_mcpl(xml, "out.mcpl.gz", 1 << 22, rng=1, force=True)
        for name, level, n in (("gpu_gzip", None, nout), ("host_zlib_level1", 1, nout_host)):
            t0 = time.perf_counter()
            kds.resample_to_mcpl(xml, "out.mcpl.gz", n, rng=1, force=True, compresslevel=level)
            dt = time.perf_counter() - t0
            res[name] = {"particles": n, "particles_per_s": n / dt, "s": dt,
                         "file_bytes": os.path.getsize("out.mcpl.gz")}
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    return {"value": res["gpu_gzip"]["particles_per_s"], "unit": "particles/s",
            "api": "resample_to_mcpl(xml, out.mcpl.gz, {}) incl. XML + {}-particle MCPL read, "
                   "upload, generation, device gzip, D2H and file write to {}".format(
                       nout, Nl, base or "tmp"),
            "d2h_bytes_per_step": res["gpu_gzip"]["file_bytes"], "h2d_bytes_per_step": Nl * 76,
            "gpu_gzip": res["gpu_gzip"], "host_zlib_level1": res["host_zlib_level1"]}


def run_ours(args):
    import torch
    import torch.distributed as td

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    from kdsource_b200 import _lib, dist, kde
    from tools.calib import calibrate

    n_mlcv = args.n or N_MLCV
    parts, w, g, scaling, data = synthetic_scaled(n_mlcv)
    bw = kde.bw_silv(D_MLCV, np.sum(w) ** 2 / np.sum(w ** 2))
    grid = np.logspace(-0.3, 0.3, G_MLCV)
    t = torch
    cal = calibrate()
    hbm, hbm_src = measured_peaks()
    peaks = {"mufu": cal["mufu_ex2_lane_ops_per_s"], "ffma2": cal["ffma2_lane_ops_per_s"],
             "ffma": cal["ffma_lane_ops_per_s"], "dfma": cal["dfma_lane_ops_per_s"],
             "hbm_gbs": hbm}

    plan = kde.MLCVPlan(data, w, bw, grid, n_splits=K_MLCV)
    flush = t.empty(256 << 20, dtype=t.uint8, device="cuda")
    for _ in range(args.warmup):
        plan.launch()
        plan.collect(allreduce=True)
    t.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    dist.barrier()
    t.cuda.synchronize()
    step_ms = []
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed steps (untimed)
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        plan.launch()
        # device reduction, NCCL all-reduce of the G partial sums and their download
        scores = plan.collect(allreduce=True)
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    dist.barrier()
    t.cuda.synchronize()
    clk = clocks.stop() if rank == 0 else None
    tot_ms = float(np.sum(step_ms))
    if world > 1:
        tt = t.tensor([tot_ms], device="cuda", dtype=t.float64)
        td.all_reduce(tt, op=td.ReduceOp.MAX)
        tot_ms = float(tt.item())
    # whole-job evaluations: all test rows x train sources x grid
    fb = kde.kfold_bounds(n_mlcv, K_MLCV)
    nf = np.diff(fb).astype(np.float64)
    evals_step = float(np.sum(nf * (n_mlcv - nf))) * G_MLCV
    value = evals_step * args.steps / (tot_ms * 1e-3)

    # end to end through the public API: host arrays in, scores out
    e2e_t = []
    for i in range(max(1, min(args.steps, 2))):
        dist.barrier()
        t.cuda.synchronize()
        t0 = time.perf_counter()
        sc2 = kde.mlcv_scores(data, w, bw, grid, n_splits=K_MLCV)
        t.cuda.synchronize()
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_t))
    if world > 1:
        tt = t.tensor([e2e_s], device="cuda", dtype=t.float64)
        td.all_reduce(tt, op=td.ReduceOp.MAX)
        e2e_s = float(tt.item())
    assert np.allclose(sc2, scores, rtol=1e-9, equal_nan=True)

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    # roofline of the dominant kernel (mlcv_pairs_kernel): one exponential per (row,
    # source, g) slot.  A fraction x of the slots evaluates 2^x on the FMA pipe
    # (ex2_poly2_scaled), so the SFU bound is mufu_rate / (1 - x).
    L = _lib.load()
    gtile = int(L.kdsb200_mlcv_gtile(G_MLCV))
    poly_slots = int(L.kdsb200_mlcv_poly_slots(G_MLCV))
    poly_x = poly_slots / float(gtile)
    sfu_peak = peaks["mufu"] / (1.0 - poly_x)
    local_evals = plan.n_evals  # rows of rank 0
    k_ms = float(np.mean(step_ms))
    # slots the kernel executes: tiles inside a CTA's own fold are skipped, partial tiles
    # and padding rows/sources are not
    slots = plan.n_slots
    achieved = slots / (k_ms * 1e-3)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_mlcv.json")
    if os.path.exists(tp) and world == 1 and n_mlcv == N_MLCV:
        with open(tp) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    sub = {}
    cpu = None
    if world == 1:
        if not args.no_sub:
            import contextlib

            with contextlib.redirect_stdout(sys.stderr):  # the API prints progress lines
                sub = sub_benchmarks(t, peaks, args.quick)
        v, dt, n = cpu_mlcv_baseline(data, w, bw, grid)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "first {} particles, G={}, K={} ({:.1f} s)".format(n, G_MLCV, K_MLCV, dt)}
        if not args.no_sub:
            try:
                cpu["reference_algorithm"] = cpu_mlcv_tree_baseline()
            except Exception as e:  # joblib workers unavailable: keep the dense baseline
                cpu["reference_algorithm"] = {"unavailable": str(e)[:200]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": mlcv_config(n_mlcv),
        "clocks": clk,
        "e2e": {"value": evals_step / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(plan.h2d_bytes), "d2h_bytes_per_step":
                int(getattr(plan, "d2h_bytes", 0)), "s_per_step": e2e_s},
        "gpu_launches": int(plan.n_launches * args.steps),
        "roofline": {"bound": "sfu", "achieved": achieved / 1e9, "peak": sfu_peak / 1e9,
                     "unit": "Gop/s", "frac": achieved / sfu_peak, "traffic": traffic,
                     "kernel": "mlcv_pairs_kernel<6,{},PM> with {}/{} polynomial slots".format(gtile, poly_slots, gtile),
                     "per_unit": "1 exponential + 2 FP32 lane-ops per (row, source, g) slot; "
                                 "{:.3g} of the exponentials run on the FMA pipe".format(poly_x),
                     "peak_source": "calibrated live: ex2.approx loop {:.4g} Gop/s on {} SMs, "
                                    "divided by (1 - {:.3g})".format(peaks["mufu"] / 1e9,
                                                                    cal["sm"], poly_x),
                     "frac_of_pure_sfu": achieved / peaks["mufu"],
                     "useful_frac": local_evals / slots},
        "calibration": cal,
        "hbm_peak_source": hbm_src,
        "scores_argmax": int(np.nanargmax(scores)),
        "rows_recomputed_fp64": int(plan.n_flagged()),
        "cpu_baseline": cpu,
        "sub": sub,
    }
    print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def run_mode(args):
    """`--mode evaluate|knn|resample`: the other three sub-paths as N-GPU benchmarks of their
    own (strong scaling of a fixed configuration, one JSON line, same timing rules as the
    headline: warm-up, CUDA events, barrier + synchronize on both sides, max over ranks)."""
    import contextlib

    import torch
    import torch.distributed as td

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    from kdsource_b200 import _lib, dist, kde
    from kdsource_b200.sampler import DeviceSampler, make_geom_struct
    from kdsource_b200.treekde import TreeKDE

    t = torch
    L = _lib.load()
    hbm, _ = measured_peaks()
    flush = t.empty(256 << 20, dtype=t.uint8, device="cuda")

    def timed_steps(fn):
        for _ in range(args.warmup):
            fn()
        dist.barrier()
        t.cuda.synchronize()
        ms = []
        for _ in range(args.steps):
            flush.zero_()
            ms += time_events(t, fn, 1)
        dist.barrier()
        t.cuda.synchronize()
        tot = t.tensor([float(np.sum(ms))], device="cuda", dtype=t.float64)
        if world > 1:
            td.all_reduce(tot, op=td.ReduceOp.MAX)
        return float(tot.item()) / args.steps

    def timed_e2e(fn):
        dist.barrier()
        t.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        t.cuda.synchronize()
        dt = t.tensor([time.perf_counter() - t0], device="cuda", dtype=t.float64)
        if world > 1:
            td.all_reduce(dt, op=td.ReduceOp.MAX)
        return float(dt.item())

    with contextlib.redirect_stdout(sys.stderr):
        if args.mode == "evaluate":
            N, M = 1_000_000, 8_000_000
            _, w, _, _, data = synthetic_scaled(N)
            q = synthetic_scaled(M, seed=54321, wseed=1)[4]
            bw = kde.bw_silv(D_MLCV, np.sum(w) ** 2 / np.sum(w ** 2))
            k = TreeKDE(bw=bw).fit(data, w)
            dev = k._device_state()
            lo, hi = dist.shard_range(M, rank, world)
            d_q = _lib.to_dev(q[lo:hi])
            ms = timed_steps(lambda: k.evaluate_device(d_q, hi - lo, dev))
            units, unit, metric = float(N) * M, "pairs/s", METRIC
            e2e_s = timed_e2e(lambda: k.evaluate(q))
            h2d, d2h = (hi - lo) * 48, (hi - lo) * 8
            work = "evaluate_1e6x8e6_silverman_auto"
            roof = {"bound": "fp32", "achieved": units * 13 / world / (ms * 1e-3) / 1e9,
                    "peak": None, "unit": "Glane-op/s", "frac": None,
                    "per_unit": "13 FP32 lane-ops + 1 MUFU per pair (d=6), rank 0", "traffic": None}
            launches = 40  # ordering + packing of the queries + the pair kernel + scatter
        elif args.mode == "knn":
            N = 10_000_000
            _, w, _, _, data = synthetic_scaled(N, seed=99)
            off = kde.array_split_bounds(N, 1000)
            b_lo, b_hi = dist.shard_range(1000, rank, world)
            # the list resident on every rank; knn_batched shards the batches over the ranks
            # itself and all-gathers the k-th distances (inside the timed step)
            d_all = _lib.to_dev(data)
            ms = timed_steps(lambda: kde.knn_batched(d_all, 11, off))
            units = float(np.sum(np.diff(off).astype(np.float64) ** 2))
            unit, metric = "pairs/s", "knn_distance_pairs_per_s"
            e2e_s = timed_e2e(lambda: kde.bw_knn(data, w, k=10, batch_size=10000))
            h2d, d2h = int(off[b_hi] - off[b_lo]) * 48, N * 8
            work = "knn_1e7_k10_batch1e4_f64"
            roof = {"bound": "fp64", "achieved": units * 17 / world / (ms * 1e-3) / 1e9,
                    "peak": None, "unit": "Gfp64 lane-op/s", "frac": None,
                    "per_unit": "17 fp64 ops per pair (no FMA contraction), rank 0; the step "
                                "includes the all-gather of the k-th distances",
                    "traffic": None}
            launches = 2
        else:
            Ns, nout = 10_000_000, 400_000_000
            ps, ws_, gs_, sc_, _ = synthetic_scaled(Ns, seed=7)
            bws = np.random.default_rng(3).uniform(0.2, 0.5, Ns).astype(np.float32)
            smp = DeviceSampler(ps, ws_, bws, make_geom_struct(gs_, sc_, "g"))
            lo, hi = dist.shard_range(nout, rank, world)
            out = _lib.empty((hi - lo) * 36, t.uint8)
            ms = timed_steps(lambda: smp.sample_device(hi - lo, seed=1, first=lo, out=out))
            del out
            units, unit, metric = float(nout), "particles/s", "resampled_particles_per_s"
            res = resample_e2e(ps, ws_, bws, gs_, sc_, False) if world == 1 else None
            e2e_s = (res["gpu_gzip"]["s"] * nout / res["gpu_gzip"]["particles"]) if res else None
            h2d, d2h = 76_000_000, int(22.5 * nout / world)
            work = "resample_1e7src_4e8out_adaptive_bw"
            roof = {"bound": "hbm", "achieved": 76.0 * nout / world / (ms * 1e-3) / 1e9,
                    "peak": hbm, "unit": "GB/s",
                    "frac": 76.0 * nout / world / (ms * 1e-3) / 1e9 / hbm,
                    "per_unit": "76 B per particle (40 read + 36 written), rank 0", "traffic": None}
            launches = 1
    if rank == 0:
        line = {"metric": metric, "value": units / (ms * 1e-3), "unit": unit, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64" if args.mode == "knn" else "f32", "data": "synthetic",
                "config": {"workload": work, "mode": args.mode,
                           "l2": "flushed between timed steps (256 MiB write, untimed)"},
                "e2e": {"value": units / e2e_s if e2e_s else None, "unit": unit,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches * args.steps, "roofline": roof}
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="override the MLCV list size (debug)")
    ap.add_argument("--quick", action="store_true", help="smaller secondary benchmarks")
    ap.add_argument("--no-sub", action="store_true", help="skip secondary benchmarks")
    ap.add_argument("--mode", default="mlcv", choices=["mlcv", "evaluate", "knn", "resample"],
                    help="mlcv (headline, BASELINE configs[1]) or one of the other sub-paths")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode != "mlcv":
        run_mode(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

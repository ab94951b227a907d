"""Resampler memory-path probe: rate of the fp32 Philox kernel with the window table (0) vs
the pick table (1) at several list sizes (RS_TABLES=0,1; RS_IDX_ONLY=1 adds a run without
record output)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_scaled, time_events  # noqa: E402
from kdsource_b200 import _lib  # noqa: E402
from kdsource_b200.sampler import DeviceSampler, make_geom_struct  # noqa: E402

t = _lib.torch()
L = _lib.load()
sizes = [int(float(a)) for a in (sys.argv[1:] or ["1e5", "1e6", "1e7"])]
for Ns in sizes:
    nout = 100_000_000
    ps, ws_, gs_, sc_, _ = synthetic_scaled(Ns, seed=7)
    bws = np.random.default_rng(3).uniform(0.2, 0.5, Ns).astype(np.float32)
    out = _lib.empty(nout * 36, t.uint8)
    ref_idx = None
    for table in [int(a) for a in os.environ.get("RS_TABLES", "0,1").split(",")]:
        smp = DeviceSampler(ps, ws_, bws, make_geom_struct(gs_, sc_, "g"), pick_table=bool(table))
        idx = smp.sample_device(1 << 20, seed=1, want_idx=True, want_records=False)["idx"].cpu().numpy()
        if ref_idx is None:
            ref_idx = idx
        assert np.array_equal(idx, ref_idx), "pick table changed the indices"
        for _ in range(2):
            smp.sample_device(nout, seed=1, out=out)
        ms = float(np.median(time_events(t, lambda: smp.sample_device(nout, seed=1, out=out), 5)))
        print("N=%g table=%d bits=%d: %.3f ms  %.3e particles/s  %.1f%% of 76B roofline" % (
            Ns, table, smp.pick_bits if table else smp.guide_bits, ms,
            nout / ms * 1e3, 76 * nout / ms * 1e3 / 6549.8e9 * 100), flush=True)
        if os.environ.get("RS_IDX_ONLY"):
            f = lambda: smp.sample_device(nout, seed=1, want_records=False, want_idx=True)
            f()
            ms = float(np.median(time_events(t, f, 5)))
            print("   idx only (8 B written per particle): %.3f ms  %.3e particles/s" % (ms, nout / ms * 1e3))
        del smp

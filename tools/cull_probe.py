"""Tile-culling probe (GPU box): visited fraction and rates of the ordered-tile kernel in
cutoff mode for growing list / query sizes.   python tools/cull_probe.py [N ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from kdsource_b200 import _lib, kde  # noqa: E402
from kdsource_b200.treekde import TreeKDE  # noqa: E402


def main():
    t = _lib.torch()
    sizes = [int(float(a)) for a in sys.argv[1:]] or [1_000_000, 10_000_000]
    for N in sizes:
        M = N
        _, w, _, _, data = bench.synthetic_scaled(N)
        _, _, _, _, q = bench.synthetic_scaled(M, seed=54321, wseed=1)
        bw = kde.bw_silv(6, np.sum(w) ** 2 / np.sum(w ** 2))
        for cutoff in ("treekde", None):
            if cutoff is None and N > 2_000_000:
                continue
            k = TreeKDE(bw=bw, cutoff=cutoff).fit(data, w)
            t0 = time.perf_counter()
            dev = k._device_state()
            t.cuda.synchronize()
            t_pack = time.perf_counter() - t0
            chunk = 1 << 22
            visited, ms = 0, 0.0
            for s in range(0, M, chunk):
                d_q = _lib.to_dev(q[s:s + chunk])
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record()
                k.evaluate_device(d_q, len(d_q), dev)
                e1.record()
                e1.synchronize()
                ms += e0.elapsed_time(e1)
                visited += int(k.last_tiles_visited.item())
            dense = float(N) * M
            executed = visited * 512.0 * 1024.0
            print(json.dumps({
                "N": N, "M": M, "bw": bw, "cutoff": cutoff, "radius": dev["cutoff"],
                "tile_extent": dev["tile_extent"], "dtype": dev["dtype"], "pack_s": t_pack,
                "eval_s": ms * 1e-3, "visited_frac": executed / dense,
                "executed_pairs_per_s": executed / (ms * 1e-3),
                "dense_equiv_pairs_per_s": dense / (ms * 1e-3)}), flush=True)
            del k, dev


if __name__ == "__main__":
    main()

import sys, time, cProfile, pstats
sys.path.insert(0, ".")
import numpy as np
from bench import synthetic_scaled
from kdsource_b200 import _lib, kde
from kdsource_b200.treekde import TreeKDE
t = _lib.torch()
N, M = 100_000, 1_000_000
_, w, _, _, data = synthetic_scaled(N)
q = synthetic_scaled(M, seed=54321, wseed=1)[4]
bw = kde.bw_silv(6, np.sum(w) ** 2 / np.sum(w ** 2))
k = TreeKDE(bw=bw).fit(data, w)
k.evaluate(q); k.evaluate(q)
t.cuda.synchronize(); t0 = time.perf_counter(); k.evaluate(q); t.cuda.synchronize(); print("evaluate", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); k.evaluate(q); t.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
off = kde.array_split_bounds(M, 100)
kde.knn_batched(q, 11, off); 
t0 = time.perf_counter(); kde.knn_batched(q, 11, off); print("knn", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); kde.knn_batched(q, 11, off); t.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

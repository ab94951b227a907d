"""2-rank check: sharded kNN / MLCV / resample equal the single-rank results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
from bench import synthetic_scaled  # noqa: E402
from kdsource_b200 import dist, kde  # noqa: E402
from kdsource_b200.sampler import DeviceSampler, make_geom_struct  # noqa: E402

rank, world = dist.rank_world()
parts, w, g, scaling, data = synthetic_scaled(30000)
grid = np.logspace(-0.3, 0.3, 12)
sc = kde.mlcv_scores(data, w, 0.4, grid, 10)
plan = kde.MLCVPlan(data, w, 0.4, grid, 10, rows=(0, len(data)))
plan.launch()
full = plan.collect()
assert np.allclose(sc, full, rtol=1e-12), (sc, full)
bw = kde.bw_knn(data, w, k=10, batch_size=5000)
off = kde.array_split_bounds(len(data), 6)
os.environ["KDS_SINGLE"] = "1"
smp = DeviceSampler(parts, w, 0.3, make_geom_struct(g, scaling, "g"))
lo, hi = dist.shard_range(100000, rank, world)
mine = smp.sample(hi - lo, seed=5, first=lo)["records"]
allrec = smp.sample(100000, seed=5, first=0)["records"]
assert np.array_equal(mine, allrec[lo:hi])
print("rank", rank, "of", world, "ok: mlcv sharded == full, knn", bw[:3], "resample split ok")
td.destroy_process_group()

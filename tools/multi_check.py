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
from kdsource_b200.treekde import TreeKDE  # noqa: E402
import tempfile  # noqa: E402
from kdsource_b200 import mcpl  # noqa: E402
from kdsource_b200.resample import resample_to_mcpl  # noqa: E402
import kdsource_b200.api as kds  # noqa: E402

k = TreeKDE(bw=0.35).fit(data, w)
q = data[:5001] + 0.01
sharded = k.evaluate(q)
whole = k.evaluate(q, shard=False)
assert np.allclose(sharded, whole, rtol=1e-12)
# multi-rank resample_to_mcpl: same file as a single-rank run (compared after decompression)
tmp = "/tmp/kds_multi"
os.makedirs(tmp, exist_ok=True)
if rank == 0:
    mcpl.write_mcpl_file(tmp + "/list.mcpl.gz", ekin=parts[:, 0], x=parts[:, 1], y=parts[:, 2],
                         z=parts[:, 3], ux=parts[:, 4], uy=parts[:, 5], uz=parts[:, 6],
                         time=parts[:, 7], weight=w, pdgcode=2112)
dist.barrier()
pl = kds.PList(tmp + "/list.mcpl.gz")
src = kds.KDSource(pl, g, bw=0.3)
src.fit()
# KDSource.evaluate shards the query points; same densities as the unsharded kernel sum
qp = parts[:4001].copy()
ev, er = src.evaluate(qp)
dens = src.kde.evaluate(g.transform(qp.copy()) / src.scaling, shard=False)
assert np.allclose(ev, dens / np.prod(src.scaling) * src.J * g.jac(qp.copy()), rtol=1e-12)
if rank == 0:
    src.save(tmp + "/src.xml", adjust_N=False)
dist.barrier()
resample_to_mcpl(tmp + "/src.xml", tmp + "/multi.mcpl.gz", 300001, rng=9, force=True)
if rank == 0:
    import gzip
    a = gzip.open(tmp + "/multi.mcpl.gz").read()
    smp2, _ = __import__("kdsource_b200.resample", fromlist=["open_source"]).open_source(tmp + "/src.xml")
    ref = mcpl.build_header(300001, "KDSource resample") + smp2.sample(300001, seed=9)["records"].tobytes()
    assert a == ref, (len(a), len(ref))
dist.barrier()
print("rank", rank, "of", world, "ok: mlcv sharded == full, knn", bw[:3], "resample split ok, evaluate sharded ok, multi-rank mcpl ok")
td.destroy_process_group()

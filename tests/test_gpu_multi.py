"""N > 1 on real GPUs: the sharded sub-paths give the single-GPU results (runs only where
the box has at least two GPUs; the host-side sharding logic is covered on CPU with gloo in
test_dist_cpu.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_ranks_equal_one_rank():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    # tools/multi_check.py asserts: MLCV scores with sharded test rows == full plan, kNN
    # bandwidths, resample ranges == slices of the whole output, sharded TreeKDE / KDSource
    # evaluate == unsharded, multi-rank resample_to_mcpl file == single-rank stream
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "multi_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok: mlcv sharded == full") == 2

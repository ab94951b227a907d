import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


def _ensure_built():
    """Build the oracle and the CUDA library if they are missing (cross-compiles
    without a GPU); the GPU box uses the prebuilt files shipped with the snapshot."""
    need = [os.path.join(ROOT, "oracle", "libkds_oracle.so"),
            os.path.join(ROOT, "kdsource_b200", "libkdsb200.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g

        g.build()


_ensure_built()


@pytest.fixture(scope="session")
def golden_kde():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "ref_kde.npz"))


@pytest.fixture(scope="session")
def golden_geom():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "ref_geom.npz"))


@pytest.fixture(scope="session")
def golden_kdsource():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "ref_kdsource.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import kds_oracle

    return kds_oracle


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

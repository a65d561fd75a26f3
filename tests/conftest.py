import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("CUDA unavailable")
    return torch.device("cuda:0")


def canonical_tie_order_equal(ids_a, ids_b, tile_ranges, depths):
    """True when two sorted-id lists agree up to the order inside groups of equal
    (tile, depth) -- the only freedom torch.argsort (unstable, binning.py:223) has."""
    ids_a = np.asarray(ids_a); ids_b = np.asarray(ids_b)
    if ids_a.shape != ids_b.shape:
        return False
    depths = np.asarray(depths)
    r = np.asarray(tile_ranges).reshape(-1, 2)
    if not np.array_equal(depths[ids_a] + 0.0, depths[ids_b] + 0.0):
        return False
    for s, e in r:
        if e - s <= 1:
            continue
        a, b = ids_a[s:e], ids_b[s:e]
        if np.array_equal(a, b):
            continue
        d = depths[a] + 0.0
        # groups of equal depth are contiguous because the list is depth-sorted
        bounds = np.flatnonzero(np.diff(d) != 0) + 1
        for ga, gb in zip(np.split(a, bounds), np.split(b, bounds)):
            if not np.array_equal(np.sort(ga), np.sort(gb)):
                return False
    return True

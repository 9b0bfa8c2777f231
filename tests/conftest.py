import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"

# most tests hand plain (pageable) numpy arrays to the host-pointer API; the library's one-time note about that on
# stderr is checked once (tests/test_gpu_host_staging.py) and silenced everywhere else
import os  # noqa: E402

os.environ.setdefault("LUMACU_QUIET", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return json.loads((GOLDEN_DIR / "golden.json").read_text())


@pytest.fixture(scope="session")
def small_cases():
    return np.load(GOLDEN_DIR / "small_cases.npz")


@pytest.fixture(scope="session")
def po():
    """The TEST-ONLY checkers (oracle/): C restatement + the compiled reference when present."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def lumalib():
    import lumahdrv_b200
    lumahdrv_b200.build_library()
    return lumahdrv_b200


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit equality of float arrays with all NaNs considered equal."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    an, bn = np.isnan(a), np.isnan(b)
    if not np.array_equal(an, bn):
        return False
    return bool(np.array_equal(a.view(np.uint32)[~an], b.view(np.uint32)[~bn]))


def max_ulp(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    ok = ~(np.isnan(a) & np.isnan(b))
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)[ok]
    return int(d.max()) if d.size else 0

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (test infrastructure)"""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def capi():
    """the product C ABI through ctypes; builds the library in-tree if it is stale and nvcc exists"""
    from librec_b200 import _build, capi as c
    try:
        _build.build()
    except RuntimeError:
        if not os.path.exists(_build.LIB_PATH):
            raise
    c.load()
    return c


@pytest.fixture(scope="session")
def c1(O):
    """config C1: the seeded ml-100k 0.8 split (committed fixture, see tests/golden/make_golden.py)"""
    z = np.load(os.path.join(GOLDEN, "ml100k_seed1_split.npz"))
    full = O.Csr(int(z["U"]), int(z["I"]), z["rowptr"].astype(np.int64), z["col"].astype(np.int32), z["val"].astype(np.float64))
    flags = z["flags"]
    tr, te = full.select(flags == 1), full.select(flags == 0)
    with open(os.path.join(GOLDEN, "oracle_c1.json")) as f:
        pins = json.load(f)
    rng_state = (int(z["rng_seed"]), int(z["rng_have"]), float(z["rng_nextg"]))
    return {"full": full, "train": tr, "test": te, "pins": pins, "rng_state": rng_state}


def rng_csr(O, U, I, density, seed, values=(1.0, 2.0, 3.0, 4.0, 5.0)):
    """random CSR with ascending columns per row"""
    rng = np.random.default_rng(seed)
    mask = rng.random((U, I)) < density
    rows, cols = np.nonzero(mask)
    rowptr = np.zeros(U + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    val = rng.choice(np.asarray(values, np.float64), size=rows.shape[0])
    return O.Csr(U, I, np.cumsum(rowptr), cols.astype(np.int32), val)

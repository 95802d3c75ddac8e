"""At-scale statistical parity: the fast kernels at FULL concurrency (item-run tiles, staleness- and
curvature-aware step) against the oracle's sequential restatement of the reference loop on the benchmark shapes, held-out
RMSE / MAE on a seeded 80/20 split of the synthetic matrix; BPR by AUC / Precision@10 (SURVEY.md 8(d) gate).
Tolerances are written in tests/parity_scale.py (TOL) and in bpr_parity's arguments."""
import numpy as np
import pytest

import parity_scale

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
def test_c2_biasedmf_heldout_rmse_matches_the_sequential_oracle(capi, O):
    r = parity_scale.rating_parity(capi, O, "c2", epochs=10)
    print("C2 parity:", r)
    assert r["rollbacks"] == 0
    assert abs(r["d_rmse"]) <= r["tol"] and abs(r["d_mae"]) <= r["tol"], r


@pytest.mark.timeout(600)
def test_c4_prefix_pmf_heldout_rmse_matches_the_sequential_oracle(capi, O):
    r = parity_scale.rating_parity(capi, O, "c4p", epochs=10)
    print("C4-prefix parity:", r)
    assert r["rollbacks"] == 0
    assert abs(r["d_rmse"]) <= r["tol"] and abs(r["d_mae"]) <= r["tol"], r


@pytest.mark.timeout(600)
def test_bpr_c1_auc_and_precision_match_the_oracle_bpr(capi, O, c1):
    """bpr-test.properties (k=10, lr 0.01, reg 0.01, 50 iterations) on the binarised seeded ml-100k split"""
    r = parity_scale.bpr_parity(capi, O, c1["train"], c1["test"], k=10, lr=0.01, reg=0.01, epochs=50)
    print("BPR C1 parity:", r)
    assert r["ok"], r


@pytest.mark.timeout(600)
def test_c2_group_kernel_heldout_rmse_matches_the_sequential_oracle(capi, O, monkeypatch):
    """the order-faithful user-group kernel (LRK_SGD_GROUP=1) on the same case"""
    monkeypatch.setenv("LRK_SGD_GROUP", "1")
    r = parity_scale.rating_parity(capi, O, "c2", epochs=10)
    print("C2 parity (group kernel):", r)
    assert r["rollbacks"] == 0
    assert abs(r["d_rmse"]) <= 1e-3 and abs(r["d_mae"]) <= 1e-3, r

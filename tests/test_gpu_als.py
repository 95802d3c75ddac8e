"""GPU: WRMF and eALS on the device (csrc/als.cuh; recommender/cf/ranking/WRMFRecommender.java:74-166, EALSRecommender.java:114-214,
SURVEY.md 8f row N3) against the oracle's restatements (pinned by the pure-Python replays of tests/test_oracle_als.py).  Both models are
deterministic and the kernels keep the reference's floating-point operation order in fp64, so the bar is BIT equality of the factors."""
import numpy as np
import pytest

from conftest import rng_csr

pytestmark = pytest.mark.gpu


def _wrmf_weights(O, val, coef=4.0):
    lut = {float(v): O.lib().lro_wrmf_weight(float(v), coef) for v in np.unique(val)}
    return np.array([lut[float(v)] for v in val])


def _csr_with_gaps(O, U, I, density, seed):
    """random CSR in which user 4 and item 3 have no entries"""
    rng = np.random.default_rng(seed)
    mask = rng.random((U, I)) < density
    mask[4, :] = False
    mask[:, 3] = False
    rows, cols = np.nonzero(mask)
    rowptr = np.zeros(U + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    val = rng.choice(np.array([1.0, 2.0, 3.0, 4.0, 5.0]), size=rows.shape[0])
    return O.Csr(U, I, np.cumsum(rowptr), cols.astype(np.int32), val)


def _run_wrmf(O, capi, tr, val, k, P, Q, reg_u, reg_i, epochs):
    oP, oQ = P.copy(), Q.copy()
    with capi.Handle(capi.MODEL_WRMF, k, seed=1) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, val)
        h.set_factors(P, Q)
        losses = [h.sgd_epoch(0.0, reg_u, reg_i, 0.0, e + 1) for e in range(epochs)]
        gP, gQ, _, _ = h.get_factors()
    for _ in range(epochs):
        O.lib().lro_wrmf_epoch(tr.U, tr.I, tr.rowptr, tr.col, val, k, oP, oQ, reg_u, reg_i)
    assert losses == [0.0] * epochs                       # trainModel never assigns `loss`
    return gP, gQ, oP, oQ


@pytest.mark.parametrize("k", [1, 5, 16, 20, 40, 64, 70, 96, 100])
def test_wrmf_factors_bit_identical_small(O, capi, k):
    """every TILE instantiation of the solve kernel (k = 1 .. 100), rows with 0 .. ~20 entries, three iterations"""
    tr = _csr_with_gaps(O, 57, 41, 0.25, 3 + k)
    val = _wrmf_weights(O, tr.val)
    rng = np.random.default_rng(k)
    P, Q = rng.normal(0, 0.1, (tr.U, k)), rng.normal(0, 0.1, (tr.I, k))
    gP, gQ, oP, oQ = _run_wrmf(O, capi, tr, val, k, P, Q, 0.01, 0.02, 3)
    assert np.array_equal(gP, oP) and np.array_equal(gQ, oQ)
    assert np.isfinite(gP).all()


def test_wrmf_singular_system_returns_the_half_built_inverse(O, capi):
    """reg 0 and a zero factor column make A singular: DenseMatrix.inverse() gives up at the empty pivot column and returns the
    inverse as it stands (DenseMatrix.java:393-394); the device must reproduce exactly that matrix"""
    tr = rng_csr(O, 31, 23, 0.3, 9)
    k = 4
    rng = np.random.default_rng(1)
    P, Q = rng.normal(0, 0.1, (tr.U, k)), rng.normal(0, 0.1, (tr.I, k))
    Q[:, 2] = 0.0
    val = _wrmf_weights(O, tr.val)
    gP, gQ, oP, oQ = _run_wrmf(O, capi, tr, val, k, P, Q, 0.0, 0.0, 1)
    assert np.array_equal(gP, oP, equal_nan=True) and np.array_equal(gQ, oQ, equal_nan=True)


def test_wrmf_c1_factors_and_lists_bit_identical(O, capi, c1):
    """wrmf-test.properties (k=20, reg 0.01, coefficient 4) on the C1 train split: two iterations, factors bit-identical, and the
    top-10 lists of the learned model equal the oracle's recommendRank"""
    tr = c1["train"]
    k = 20
    val = _wrmf_weights(O, tr.val)
    rng = np.random.default_rng(7)
    P, Q = rng.normal(0, 0.1, (tr.U, k)), rng.normal(0, 0.1, (tr.I, k))
    oP, oQ = P.copy(), Q.copy()
    with capi.Handle(capi.MODEL_WRMF, k, seed=1) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, val)
        h.set_factors(P, Q)
        for e in range(2):
            h.sgd_epoch(0.0, 0.01, 0.01, 0.0, e + 1)
        gP, gQ, _, _ = h.get_factors()
        users = np.arange(tr.U, dtype=np.int32)
        items, scores, counts = h.topn(10, users=users)
    for _ in range(2):
        O.lib().lro_wrmf_epoch(tr.U, tr.I, tr.rowptr, tr.col, val, k, oP, oQ, 0.01, 0.01)
    assert np.array_equal(gP, oP) and np.array_equal(gQ, oQ)
    oi, os_, oc = O.recommend_rank(1, tr.U, tr.I, k, oP, oQ, None, None, 0.0, tr, 10)
    assert np.array_equal(counts, oc) and np.array_equal(items, oi) and np.array_equal(scores, os_)
    # the model learned the train matrix: observed entries score far above the rest
    S = gP[:50] @ gQ.T
    seen = np.zeros_like(S, bool)
    for u in range(50):
        seen[u, tr.col[tr.rowptr[u]:tr.rowptr[u + 1]]] = True
    assert S[seen].mean() > S[~seen].mean() + 0.3


@pytest.mark.parametrize("heavy", [0, 8])
@pytest.mark.parametrize("judge", [0, 1, 2])
@pytest.mark.parametrize("k", [4, 40])
def test_eals_factors_bit_identical_small(O, capi, monkeypatch, judge, k, heavy):
    """heavy = 8 sends every row with more than 8 entries down the CTA-per-row kernel (default threshold 512)"""
    if heavy:
        monkeypatch.setenv("LRK_EALS_HEAVY", str(heavy))
    tr = _csr_with_gaps(O, 61, 37, 0.3, 10 * judge + k)
    conf = np.zeros(tr.I)
    O.lib().lro_eals_confidences(tr.U, tr.I, tr.rowptr, tr.col, 0.4, 128.0, judge, conf)
    val = np.array([O.lib().lro_eals_weight(float(v), 1.0, judge) for v in tr.val])
    rng = np.random.default_rng(k + judge)
    P, Q = np.zeros((tr.U, k)), rng.normal(0, 0.1, (tr.I, k))
    oP, oQ = P.copy(), Q.copy()
    with capi.Handle(capi.MODEL_EALS, k, seed=1) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, val)
        with pytest.raises(capi.LibrecException):
            h.set_factors(P, Q)
            h.sgd_epoch(0.0, 0.01, 0.02, 0.0, 1)           # no confidences yet
        h.set_matrix("eals.confidences", conf)
        assert np.array_equal(h.get_matrix("eals.confidences", (tr.I,)), conf)
        h.set_factors(P, Q)
        for e in range(3):
            assert h.sgd_epoch(0.0, 0.01, 0.02, 0.0, e + 1) == 0.0
        gP, gQ, _, _ = h.get_factors()
    for _ in range(3):
        O.lib().lro_eals_epoch(tr.U, tr.I, tr.rowptr, tr.col, val, k, oP, oQ, conf, 0.01, 0.02)
    assert np.array_equal(gP, oP) and np.array_equal(gQ, oQ)
    assert np.isfinite(gP).all() and np.abs(gP).max() > 0


@pytest.mark.parametrize("heavy", [0, 100])
def test_eals_c1_k200_bit_identical(O, capi, c1, monkeypatch, heavy):
    """eals-test.properties: k=200, reg 0.01, judge 1, coefficient 1 -- one iteration on the C1 train split, then the lists
    (heavy = 100: a third of the users and the popular items take the CTA-per-row kernel)"""
    if heavy:
        monkeypatch.setenv("LRK_EALS_HEAVY", str(heavy))
    tr = c1["train"]
    k = 200
    conf = np.ones(tr.I)
    val = np.array([O.lib().lro_eals_weight(float(v), 1.0, 1) for v in tr.val])
    rng = np.random.default_rng(21)
    P, Q = np.zeros((tr.U, k)), rng.normal(0, 0.1, (tr.I, k))
    oP, oQ = P.copy(), Q.copy()
    with capi.Handle(capi.MODEL_EALS, k, seed=1) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, val)
        h.set_matrix("eals.confidences", conf)
        h.set_factors(P, Q)
        h.sgd_epoch(0.0, 0.01, 0.01, 0.0, 1)
        gP, gQ, _, _ = h.get_factors()
        users = np.arange(0, tr.U, 3, dtype=np.int32)
        items, scores, counts = h.topn(10, users=users)
    O.lib().lro_eals_epoch(tr.U, tr.I, tr.rowptr, tr.col, val, k, oP, oQ, conf, 0.01, 0.01)
    assert np.array_equal(gP, oP) and np.array_equal(gQ, oQ)
    oi, os_, oc = O.recommend_rank(1, tr.U, tr.I, k, oP, oQ, None, None, 0.0, tr, 10, users=users)
    assert np.array_equal(counts, oc) and np.array_equal(items, oi) and np.array_equal(scores, os_)


def test_als_argument_errors(capi):
    with pytest.raises(capi.LibrecException):
        capi.Handle(capi.MODEL_WRMF, 113, seed=1)
    with pytest.raises(capi.LibrecException):
        capi.Handle(capi.MODEL_EALS, 8, seed=1, devices=[0, 1])

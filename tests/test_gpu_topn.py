"""GPU: prediction / top-N through the C ABI must be bit-identical to the oracle for identical factors."""
import numpy as np
import pytest

from conftest import rng_csr

pytestmark = pytest.mark.gpu


def _setup(capi, O, model, U, I, k, seed, density=0.05, scale=0.1, topn_path=1):
    rng = np.random.default_rng(seed)
    P = rng.normal(0, scale, (U, k)); Q = rng.normal(0, scale, (I, k))
    biased = model == capi.MODEL_BIASEDMF
    bu = rng.normal(0, scale, U) if biased else None
    bi = rng.normal(0, scale, I) if biased else None
    tr = rng_csr(O, U, I, density, seed + 1)
    h = capi.Handle(model, k, topn_path=topn_path)
    h.set_train_csr(U, I, tr.rowptr, tr.col, tr.val)
    h.set_factors(P, Q, bu, bi, 3.53)
    return h, tr, P, Q, bu, bi


def _omodel(capi, O, model):
    return {capi.MODEL_BIASEDMF: O.BIASEDMF, capi.MODEL_PMF: O.PMF, capi.MODEL_BPR: O.BPR}[model]


@pytest.mark.parametrize("model", [0, 1, 2])
@pytest.mark.parametrize("U,I,k,N", [(70, 333, 20, 10), (33, 1000, 64, 10), (40, 257, 128, 7), (9, 31, 3, 50), (64, 96, 200, 10)])
def test_topn_exact_bit_identical(O, capi, model, U, I, k, N):
    h, tr, P, Q, bu, bi = _setup(capi, O, model, U, I, k, seed=U + k)
    with h:
        items, scores, counts = h.topn(N)
    oi, os_, oc = O.recommend_rank(_omodel(capi, O, model), U, I, k, P, Q, bu, bi, 3.53, tr, N)
    assert np.array_equal(counts, oc) and np.array_equal(items, oi)
    assert np.array_equal(scores.view(np.int64), os_.view(np.int64))      # bit-exact fp64


def test_topn_ties_follow_java_heap_order(O, capi):
    U, I, k, N = 50, 400, 8, 10
    rng = np.random.default_rng(4)
    P = rng.integers(-2, 3, (U, k)).astype(np.float64)      # small integers -> massive score ties
    Q = rng.integers(-1, 2, (I, k)).astype(np.float64)
    Q[::7] = 0.0; P[3] = 0.0; Q[5, 0] = -0.0
    tr = rng_csr(O, U, I, 0.1, 8)
    with capi.Handle(capi.MODEL_PMF, k, topn_path=1) as h:
        h.set_train_csr(U, I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q)
        items, scores, counts = h.topn(N)
    oi, os_, oc = O.recommend_rank(O.PMF, U, I, k, P, Q, None, None, 0.0, tr, N)
    assert np.array_equal(counts, oc) and np.array_equal(items, oi)
    assert np.array_equal(scores.view(np.int64), os_.view(np.int64))


def test_topn_edge_cases(O, capi):
    U, I, k = 6, 40, 4
    rng = np.random.default_rng(0)
    P = rng.normal(size=(U, k)); Q = rng.normal(size=(I, k))
    Q[7, 1] = np.nan                                           # NaN score dropped (MatrixRecommender.java:186)
    rowptr = [0, 40, 40, 78, 79, 79, 79]                       # user0: everything trained; user2: all but 2 items
    col = list(range(40)) + [c for c in range(40) if c not in (3, 30)] + [39]
    tr = O.Csr(U, I, rowptr, col, np.ones(len(col)))
    with capi.Handle(capi.MODEL_PMF, k, topn_path=1) as h:
        h.set_train_csr(U, I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q)
        full = h.topn(10)
        sub = h.topn(10, users=[5, 2, 2, 0])
        noex = h.topn(45, exclude_train=False)
        with pytest.raises(capi.LibrecException):
            h.topn(0)
    oi, os_, oc = O.recommend_rank(O.PMF, U, I, k, P, Q, None, None, 0.0, tr, 10)
    assert oc.tolist()[:3] == [0, 10, 2]
    for got, exp in zip(full, (oi, os_, oc)):
        assert np.array_equal(got, exp, equal_nan=True)
    si, ss, sc = O.recommend_rank(O.PMF, U, I, k, P, Q, None, None, 0.0, tr, 10, users=[5, 2, 2, 0])
    assert np.array_equal(sub[0], si) and np.array_equal(sub[1], ss) and np.array_equal(sub[2], sc)
    ni, ns, nc = O.recommend_rank(O.PMF, U, I, k, P, Q, None, None, 0.0, None, 45)
    assert np.array_equal(noex[0], ni) and np.array_equal(noex[1], ns) and nc.tolist() == [39] * U


def test_predict_pairs_and_eval_rating_bit_identical(O, capi):
    U, I, k = 300, 500, 20
    h, tr, P, Q, bu, bi = _setup(capi, O, capi.MODEL_BIASEDMF, U, I, k, seed=5, scale=0.7)
    te = rng_csr(O, U, I, 0.03, 77)
    rng = np.random.default_rng(1)
    us = rng.integers(0, U, 4000).astype(np.int32); its = rng.integers(0, I, 4000).astype(np.int32)
    with h:
        got = h.predict_pairs(us, its)
        rmse, mae, pred = h.eval_rating(U, te.rowptr, te.col, te.val, 1.0, 5.0, want_pred=True)
    exp = np.zeros(4000)
    O.lib().lro_predict_pairs(O.BIASEDMF, k, P, Q, bu.ctypes.data, bi.ctypes.data, 3.53, us, its, 4000, exp)
    assert np.array_equal(got.view(np.int64), exp.view(np.int64))
    ormse, omae, opred = O.eval_rating(O.BIASEDMF, te, k, P, Q, bu, bi, 3.53, 1.0, 5.0, want_pred=True)
    assert np.array_equal(pred.view(np.int64), opred.view(np.int64))            # bounded predictions bit-exact
    assert abs(rmse - ormse) <= 1e-13 * ormse and abs(mae - omae) <= 1e-13 * omae   # tree vs sequential fp64 sum


def test_topn_after_training_uses_current_factors(O, capi, c1):
    tr = c1["train"]
    O.lib().lro_rng_set_state(*c1["rng_state"])
    P, Q, bu, bi = O.mf_setup(tr.U, tr.I, 20, True)
    mu = c1["pins"]["global_mean"]
    with capi.Handle(capi.MODEL_BIASEDMF, 20, topn_path=1) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q, bu, bi, mu)
        for it in range(5):
            h.sgd_epoch(0.002, 0.01, 0.01, 0.01, it + 1)
        items, scores, counts = h.topn(10)
        gP, gQ, gbu, gbi = h.get_factors()
    oi, os_, oc = O.recommend_rank(O.BIASEDMF, tr.U, tr.I, 20, gP, gQ, gbu, gbi, mu, tr, 10)
    assert np.array_equal(items, oi) and np.array_equal(scores.view(np.int64), os_.view(np.int64)) and np.array_equal(counts, oc)


# ---- tensor-core candidate path (tcgen05 / TMA) must still deliver the reference's lists bit for bit ----
@pytest.mark.parametrize("model,U,I,k,N,scale", [
    (1, 700, 9000, 128, 10, 0.1),      # PMF, two K blocks
    (2, 300, 5000, 64, 10, 0.1),       # BPR, one K block
    (0, 513, 4097, 20, 10, 0.3),       # BiasedMF: item bias folded into two extra K columns (Kp = 64)
    (0, 260, 3000, 100, 15, 0.1),      # BiasedMF k=100 -> Kp = 128 (two K blocks), N = 15
    (1, 1000, 20000, 128, 1, 0.05),    # N = 1
    (2, 600, 9000, 64, 20, 0.1),       # N = 20: K' = 26
    (0, 520, 8200, 30, 26, 0.2),       # N = 26: the largest N of the tensor-core path (K' = 32)
])
def test_topn_tensor_core_path_bit_identical(O, capi, model, U, I, k, N, scale):
    h, tr, P, Q, bu, bi = _setup(capi, O, model, U, I, k, seed=U + I + k, density=0.01, scale=scale, topn_path=2)
    with h:
        items, scores, counts = h.topn(N)
        stats = h.topn_stats()
        sub_users = np.random.default_rng(0).integers(0, U, 300).astype(np.int32)
        sub = h.topn(N, users=sub_users)
        noex = h.topn(N, exclude_train=False, nq=128)
    oi, os_, oc = O.recommend_rank(_omodel(capi, O, model), U, I, k, P, Q, bu, bi, 3.53, tr, N)
    assert np.array_equal(counts, oc) and np.array_equal(items, oi)
    assert np.array_equal(scores.view(np.int64), os_.view(np.int64))
    # random Gaussian factors: the certificate should hold for (almost) every user
    assert stats["fast_users"] + stats["fallback_users"] == U and stats["fast_users"] >= 0.9 * U, stats
    si, ss, sc = O.recommend_rank(_omodel(capi, O, model), U, I, k, P, Q, bu, bi, 3.53, tr, N, users=sub_users)
    assert np.array_equal(sub[0], si) and np.array_equal(sub[1].view(np.int64), ss.view(np.int64)) and np.array_equal(sub[2], sc)
    ni, ns, nc = O.recommend_rank(_omodel(capi, O, model), U, I, k, P, Q, bu, bi, 3.53, None, N, users=np.arange(128))
    assert np.array_equal(noex[0], ni) and np.array_equal(noex[1].view(np.int64), ns.view(np.int64)) and np.array_equal(noex[2], nc)


def test_topn_tensor_core_path_degenerate_rows_fall_back(O, capi):
    """ties / zero rows / heavy train rows cannot be certified -> exact kernel re-does them, result still exact"""
    U, I, k, N = 300, 4000, 16, 10
    rng = np.random.default_rng(4)
    P = rng.integers(-2, 3, (U, k)).astype(np.float64)
    Q = rng.integers(-1, 2, (I, k)).astype(np.float64)
    P[3] = 0.0; Q[::7] = 0.0
    P[100:] = rng.normal(0, 0.1, (U - 100, k))
    Q2 = Q.copy(); Q2[2000:] = rng.normal(0, 0.1, (I - 2000, k))
    tr = rng_csr(O, U, I, 0.05, 8)
    with capi.Handle(capi.MODEL_PMF, k, topn_path=2) as h:
        h.set_train_csr(U, I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q2)
        items, scores, counts = h.topn(N)
        stats = h.topn_stats()
    oi, os_, oc = O.recommend_rank(O.PMF, U, I, k, P, Q2, None, None, 0.0, tr, N)
    assert np.array_equal(counts, oc) and np.array_equal(items, oi)
    assert np.array_equal(scores.view(np.int64), os_.view(np.int64))
    assert stats["fallback_users"] >= 1


def test_topn_tensor_core_after_training(O, capi, c1):
    """learned (non-Gaussian) factors: ml-100k BPR-style ranking through the tensor-core path"""
    tr = c1["train"]
    O.lib().lro_rng_set_state(*c1["rng_state"])
    P, Q, bu, bi = O.mf_setup(tr.U, tr.I, 64, True)
    mu = c1["pins"]["global_mean"]
    with capi.Handle(capi.MODEL_BIASEDMF, 64, topn_path=2) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q, bu, bi, mu)
        for it in range(30):
            h.sgd_epoch(0.01, 0.01, 0.01, 0.01, it + 1)
        items, scores, counts = h.topn(10)
        stats = h.topn_stats()
        gP, gQ, gbu, gbi = h.get_factors()
    oi, os_, oc = O.recommend_rank(O.BIASEDMF, tr.U, tr.I, 64, gP, gQ, gbu, gbi, mu, tr, 10)
    assert np.array_equal(items, oi) and np.array_equal(scores.view(np.int64), os_.view(np.int64)) and np.array_equal(counts, oc)
    print("ml-100k learned factors:", stats)


def test_topn_second_sweep_replaces_exact_fallback(O, capi, monkeypatch):
    """K' barely above N (and enough users that the catalogue is not chunked): many rows fail the margin test after the
    first sweep; they are swept again from the threshold their first result implies (topn_tc.cuh) and must come out
    bit-identical without touching the exact kernel"""
    monkeypatch.setenv("LRK_TC_KEEP", "11")
    U, I, k, N = 38000, 20000, 32, 10
    rng = np.random.default_rng(77)
    P = rng.normal(0, 0.1, (U, k)); Q = rng.normal(0, 0.1, (I, k))
    edges = np.linspace(0, I, 9).astype(np.int64)                     # 8 train items per user: one per stratum -> ascending
    cols = (edges[:-1][None, :] + (rng.random((U, 8)) * np.diff(edges)[None, :]).astype(np.int64)).astype(np.int32)
    tr = O.Csr(U, I, np.arange(U + 1, dtype=np.int64) * 8, np.ascontiguousarray(cols.reshape(-1)), np.ones(U * 8))
    with capi.Handle(capi.MODEL_PMF, k, topn_path=2) as h:
        h.set_train_csr(U, I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q)
        items, scores, counts = h.topn(N)
        stats = h.topn_stats()
    sample = np.linspace(0, U - 1, 1500).astype(np.int32)
    oi, os_, oc = O.recommend_rank(O.PMF, U, I, k, P, Q, None, None, 0.0, tr, N, users=sample)
    assert np.array_equal(counts[sample], oc) and np.array_equal(items[sample], oi)
    assert np.array_equal(scores[sample].view(np.int64), os_.view(np.int64))
    assert stats["resweep_users"] > 100 and stats["fallback_users"] <= stats["resweep_users"] // 4, stats
    assert 0.0 < stats["sweep_error_over_bound"] < 1.0, stats


@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("pscale,qscale", [(1e-3, 1e-3), (300.0, 2e-4), (1e-6, 50.0)])
def test_topn_tensor_core_operand_scaling(O, capi, model, pscale, qscale):
    """fp16 operands after exact power-of-two scaling: tiny (reference init N(0, 0.001^2)) and huge factor magnitudes
    must neither underflow nor overflow in the sweep; the certificate self-check stays inside its bound"""
    U, I, k, N = 600, 12000, 64, 10
    rng = np.random.default_rng(int(pscale * 1e7) + model)
    P = rng.normal(0, pscale, (U, k)); Q = rng.normal(0, qscale, (I, k))
    P[::5] *= 100.0                                                   # rows of very different magnitude
    biased = model == capi.MODEL_BIASEDMF
    bu = rng.normal(0, pscale * qscale, U) if biased else None
    bi = rng.normal(0, pscale * qscale * 3, I) if biased else None
    tr = rng_csr(O, U, I, 0.003, 5)
    with capi.Handle(model, k, topn_path=2) as h:
        h.set_train_csr(U, I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q, bu, bi, 3.53)
        items, scores, counts = h.topn(N)
        stats = h.topn_stats()
    oi, os_, oc = O.recommend_rank(_omodel(capi, O, model), U, I, k, P, Q, bu, bi, 3.53, tr, N)
    assert np.array_equal(counts, oc) and np.array_equal(items, oi)
    assert np.array_equal(scores.view(np.int64), os_.view(np.int64))
    assert stats["fast_users"] >= 0.9 * U and stats["sweep_error_over_bound"] < 1.0, stats


def test_topn_tensor_core_nonfinite_items_use_exact_kernel(O, capi):
    U, I, k, N = 600, 9000, 32, 10
    h, tr, P, Q, bu, bi = _setup(capi, O, capi.MODEL_PMF, U, I, k, seed=9, density=0.003, scale=0.1, topn_path=2)
    Q[17, 3] = np.inf; Q[4000, 0] = np.nan
    with h:
        h.set_factors(P, Q)
        items, scores, counts = h.topn(N)
        stats = h.topn_stats()
    oi, os_, oc = O.recommend_rank(O.PMF, U, I, k, P, Q, None, None, 3.53, tr, N)
    assert np.array_equal(counts, oc) and np.array_equal(items, oi)
    assert np.array_equal(scores.view(np.int64), os_.view(np.int64))
    assert stats["fallback_users"] == U and stats["fast_users"] == 0


# ---- ranking evaluators on the device: lists never leave the GPU between recommendRank and the measures ----
@pytest.mark.parametrize("U,I,k,N,path", [(700, 9000, 32, 10, 2), (300, 70000, 16, 5, 2), (120, 400, 8, 10, 1), (64, 3000, 8, 64, 1)])
def test_eval_ranking_matches_oracle(O, capi, U, I, k, N, path):
    rng = np.random.default_rng(U + I)
    P = rng.normal(0, 0.1, (U, k)); Q = rng.normal(0, 0.1, (I, k))
    tr = rng_csr(O, U, I, min(0.02, 40.0 / I), 3)
    # test rows: disjoint from train, some users without any test item, ratings 1..5
    rowptr, col, val = [0], [], []
    for u in range(U):
        n = 0 if u % 7 == 0 else int(rng.integers(1, 30))
        cand = np.setdiff1d(rng.choice(I, size=min(I, 3 * n + 5), replace=False), tr.col[tr.rowptr[u]:tr.rowptr[u + 1]])[:n]
        cand.sort()
        col += cand.tolist(); val += rng.integers(1, 6, cand.shape[0]).astype(float).tolist()
        rowptr.append(len(col))
    te = O.Csr(U, I, np.asarray(rowptr, np.int64), np.asarray(col, np.int32), np.asarray(val, np.float64))
    with capi.Handle(capi.MODEL_PMF, k, topn_path=path) as h:
        h.set_train_csr(U, I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q)
        got, (items, scores, counts) = h.eval_ranking(N, te.rowptr, te.col, te.val, want_lists=True)
        only = h.eval_ranking(N, te.rowptr, te.col, te.val)
    oi, os_, oc = O.recommend_rank(O.PMF, U, I, k, P, Q, None, None, 0.0, tr, N)
    assert np.array_equal(items, oi) and np.array_equal(counts, oc) and np.array_equal(scores.view(np.int64), os_.view(np.int64))
    exp = O.eval_ranking(te, tr, N, oi, oc)
    for name in O.RANKING_MEASURES:
        assert abs(got[name] - exp[name]) <= 1e-12 * max(1.0, abs(exp[name])), (name, got[name], exp[name])
        assert abs(got[name] - only[name]) <= 1e-12 * max(1.0, abs(exp[name]))
    assert exp["Precision"] > 0 or exp["AUC"] > 0

"""GPU: AoBPR on the device (csrc/sgd_aobpr.cuh + the aobpr_draw sampler of csrc/sgd.cuh; recommender/cf/ranking/AoBPRRecommender.java:60-200,
SURVEY.md 8f row N3) against the oracle's restatement (lro_aobpr_train, pinned by an independent pure-Python replay in
tests/test_oracle_aobpr.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ones(O, m):
    return O.Csr(m.U, m.I, m.rowptr, m.col, np.ones_like(m.val))


def test_aobpr_samples_follow_the_adaptive_sampler(O, capi, c1):
    """(u, i) is a train entry drawn uniformly over the ENTRIES (so users are drawn by degree, unlike BPR); j is unrated by u and comes
    from the head of a factor ranking: with lambda = 0.05 * numItems the rank is < lambda with probability 1 - 1/e ~ 0.63"""
    tr = _ones(O, c1["train"])
    k = 10
    rng = np.random.default_rng(5)
    P, Q = rng.normal(0, 0.1, (tr.U, k)), rng.normal(0, 0.1, (tr.I, k))
    with capi.Handle(capi.MODEL_AOBPR, k, seed=3) as h:
        h.set_param("aobpr.lambda", 0.05)
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q)
        s = h.bpr_peek_samples(1, 0, 40000)
    u, i, j = s[:, 0], s[:, 1], s[:, 2]
    assert (u >= 0).all()
    rows = [set(tr.col[tr.rowptr[x]:tr.rowptr[x + 1]].tolist()) for x in range(tr.U)]
    for t in range(0, 40000, 11):
        assert i[t] in rows[u[t]] and j[t] not in rows[u[t]]
    # users by degree: correlation of the draw counts with the degrees
    deg = np.diff(tr.rowptr).astype(np.float64)
    cnt = np.bincount(u, minlength=tr.U).astype(np.float64)
    assert np.corrcoef(deg, cnt)[0, 1] > 0.95
    # the negative sits near the top (p_uf > 0) or the bottom (p_uf <= 0) of SOME factor's ranking far more often than a uniform item would
    Pf = P.astype(np.float32).astype(np.float64); Qf = Q.astype(np.float32).astype(np.float64)
    order = np.argsort(-Qf, axis=0, kind="stable")            # [rank, f] -> item
    rank_of = np.empty_like(order)
    for f in range(k):
        rank_of[order[:, f], f] = np.arange(tr.I)
    lam = int(np.float32(0.05) * np.float32(tr.I))
    near = 0
    for t in range(0, 40000, 11):
        r = np.where(Pf[u[t]] > 0, rank_of[j[t]], tr.I - 1 - rank_of[j[t]])
        near += int((r < lam).any())
    frac = near / len(range(0, 40000, 11))
    uniform = 1.0 - (1.0 - lam / tr.I) ** k
    assert frac > 0.6 and frac > uniform + 0.15, (frac, uniform)


def test_aobpr_training_matches_the_oracle_in_loss_and_quality(O, capi, c1):
    tr, te = _ones(O, c1["train"]), c1["test"]
    k, lr, reg, lam, epochs = 10, 0.05, 0.01, 0.05, 8
    rng = np.random.default_rng(5)
    P0, Q0 = rng.normal(0, 0.1, (tr.U, k)), rng.normal(0, 0.1, (tr.I, k))
    oP, oQ = P0.copy(), Q0.copy()
    ol = np.zeros(epochs)
    O.lib().lro_seed(1)
    assert O.lib().lro_aobpr_train(tr.U, tr.I, tr.rowptr, tr.col, k, oP, oQ, lr, reg, reg, lam, epochs, ol.ctypes.data, None) == 0
    with capi.Handle(capi.MODEL_AOBPR, k, seed=1) as h:
        h.set_param("aobpr.lambda", lam)
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P0, Q0)
        gl = [h.sgd_epoch(lr, reg, reg, 0.0, it + 1) for it in range(epochs)]
        gP, gQ, _, _ = h.get_factors()
        users = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)
        items, scores, counts = h.topn(10, users=users)

    def precision(P, Q):
        it, _, cn = O.recommend_rank(O.BPR, tr.U, tr.I, k, P, Q, None, None, 0.0, tr, 10, users=users)
        hits = sum(np.intersect1d(it[r, :cn[r]], te.col[te.rowptr[x]:te.rowptr[x + 1]]).shape[0] for r, x in enumerate(users))
        return hits / (10.0 * users.shape[0]), it, cn
    gq, ei, ec = precision(gP, gQ)
    oq, _, _ = precision(oP, oQ)
    print("AoBPR loss_1 %.1f (oracle %.1f) loss_%d %.1f (oracle %.1f)  Precision@10 %.4f (oracle %.4f)" % (gl[0], ol[0], epochs, gl[-1], ol[-1], gq, oq))
    assert abs(gl[0] - ol[0]) < 0.03 * ol[0] and abs(gl[-1] - ol[-1]) < 0.05 * ol[-1]
    assert abs(gq - oq) < 0.03
    assert np.array_equal(items, ei) and np.array_equal(counts, ec)

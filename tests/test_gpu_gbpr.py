"""GPU: GBPR on the device (csrc/sgd_gbpr.cuh; recommender/cf/ranking/GBPRRecommender.java:82-172) against the
oracle's restatement (lro_gbpr_epoch, itself pinned by an independent pure-Python replay in tests/test_oracle_gbpr.py)."""
import numpy as np
import pytest

import parity_scale

pytestmark = pytest.mark.gpu


def _ones(O, m):
    return O.Csr(m.U, m.I, m.rowptr, m.col, np.ones_like(m.val))


def test_gbpr_samples_follow_the_reference_sampler(O, capi, c1):
    tr = _ones(O, c1["train"])
    with capi.Handle(capi.MODEL_GBPR, 10, seed=5) as h:
        h.set_param("gbpr.gsize", 3)
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        s = h.bpr_peek_samples(1, 0, 20000)
    u, i, j, grp = s[:, 0], s[:, 1], s[:, 2], s[:, 3:]
    assert (u >= 0).all()
    rows = [set(tr.col[tr.rowptr[x]:tr.rowptr[x + 1]].tolist()) for x in range(tr.U)]
    users_of = {}
    for x in range(tr.U):
        for it in rows[x]:
            users_of.setdefault(it, set()).add(x)
    for t in range(0, 20000, 7):
        assert i[t] in rows[u[t]] and j[t] not in rows[u[t]]
        g = [x for x in grp[t] if x >= 0]
        col = users_of[i[t]]
        assert len(set(g)) == len(g) and set(g) <= col
        if len(col) <= 3:
            assert set(g) == col
        else:
            assert len(g) == 3 and g[0] == u[t]
    # users uniform over the users with ratings: a coarse chi-square over 10 buckets of the user range
    cnt = np.bincount(u * 10 // tr.U, minlength=10).astype(np.float64)
    nz = np.array([sum(1 for x in range(b * tr.U // 10, (b + 1) * tr.U // 10) if len(rows[x]) > 0) for b in range(10)], np.float64)
    exp = nz / nz.sum() * u.shape[0]
    assert ((cnt - exp) ** 2 / exp).sum() < 40.0


def test_gbpr_epoch_loss_and_quality_match_the_oracle(O, capi, c1):
    """gbpr-test-like settings on the binarised C1 split: first-epoch loss (the factors are frozen inside an epoch, so it is a pure
    function of the sample distribution) within 2 %; after 30 epochs AUC / Precision@10 within the tolerance the two RNG streams allow;
    lists for the learned factors bit-identical to the oracle's (prediction = b_i + p_u.q_i)."""
    tr, te = _ones(O, c1["train"]), c1["test"]
    k, lr, reg, regb, rho, glen, epochs = 10, 0.05, 0.01, 0.01, 1.5, 2, 30
    rng = np.random.default_rng(3)
    P0, Q0 = rng.normal(0, 0.01, (tr.U, k)), rng.normal(0, 0.01, (tr.I, k))
    b0 = np.zeros(tr.I)
    oP, oQ, ob = P0.copy(), Q0.copy(), b0.copy()
    O.lib().lro_seed(11)
    ol = [O.lib().lro_gbpr_epoch(tr.U, tr.I, tr.rowptr, tr.col, k, oP, oQ, ob, lr, reg, reg, regb, rho, glen, None, None) for _ in range(epochs)]
    with capi.Handle(capi.MODEL_GBPR, k, seed=1) as h:
        h.set_param("gbpr.rho", rho)
        h.set_param("gbpr.gsize", glen)
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P0, Q0, None, b0)
        gl = [h.sgd_epoch(lr, reg, reg, regb, it + 1) for it in range(epochs)]
        gP, gQ, gbu, gbi = h.get_factors()
        users = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)[:200]
        items, scores, counts = h.topn(10, users=users)
    print("GBPR loss_1 %.1f (oracle %.1f)  loss_%d %.1f (oracle %.1f)" % (gl[0], ol[0], epochs, gl[-1], ol[-1]))
    assert abs(gl[0] - ol[0]) < 0.02 * ol[0]          # measured 0.7 %: the oracle's item biases move sequentially, the device's Hogwild
    assert abs(gl[-1] - ol[-1]) < 0.05 * ol[-1]
    assert np.all(gbu == 0.0)
    oi, os_, oc = O.recommend_rank(O.BIASEDMF, tr.U, tr.I, k, gP, gQ, gbu, gbi, 0.0, tr, 10, users=users)
    assert np.array_equal(items, oi) and np.array_equal(scores.view(np.int64), os_.view(np.int64)) and np.array_equal(counts, oc)

    def quality(P, Q, b):
        uu = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)
        it, _, cn = O.recommend_rank(O.BIASEDMF, tr.U, tr.I, k, P, Q, np.zeros(tr.U), b, 0.0, tr, 10, users=uu)
        hits = sum(np.intersect1d(it[r, :cn[r]], te.col[te.rowptr[x]:te.rowptr[x + 1]]).shape[0] for r, x in enumerate(uu))
        return hits / (10.0 * uu.shape[0])
    gq, oq = quality(gP, gQ, gbi), quality(oP, oQ, ob)
    print("GBPR Precision@10 %.4f (oracle %.4f)" % (gq, oq))
    assert abs(gq - oq) < 0.03

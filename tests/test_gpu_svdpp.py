"""GPU: SVD++ on the device (csrc/sgd_svdpp.cuh; recommender/cf/rating/SVDPlusPlusRecommender.java:62-123) against
the oracle's restatement (lro_svdpp_epoch, pinned by a hand-worked update in tests/test_oracle_pins.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

f32 = lambda a: np.asarray(a, np.float64).astype(np.float32).astype(np.float64)


def test_svdpp_users_with_private_items_match_the_oracle_epoch(O, capi):
    """every item is rated by exactly one user, so users share nothing: the device's user-major walk is the reference's loop and one
    epoch equals the oracle's up to fp32 rounding -- for p_u, b_u (sequential inside a row), q_i, b_i and the implicit factors y_j"""
    rng = np.random.default_rng(7)
    U, k = 600, 20
    deg = rng.integers(0, 40, U)
    deg[:5] = [0, 1, 2, 33, 39]
    I = int(deg.sum()) + 50
    perm = rng.permutation(I)[:deg.sum()]
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = np.concatenate([np.sort(perm[rowptr[u]:rowptr[u + 1]]) for u in range(U)]).astype(np.int32)
    val = rng.integers(1, 11, col.shape[0]).astype(np.float64) / 2.0
    tr = O.Csr(U, I, rowptr, col, val)
    P, Q, Y = f32(rng.normal(0, 0.1, (U, k))), f32(rng.normal(0, 0.1, (I, k))), f32(rng.normal(0, 0.1, (I, k)))
    bu, bi = f32(rng.normal(0, 0.1, U)), f32(rng.normal(0, 0.1, I))
    mu, lr, ru, ri, rb, rimp = 3.0, 0.01, 0.02, 0.03, 0.04, 0.015
    with capi.Handle(capi.MODEL_SVDPP, k) as h:
        h.set_param("svdpp.reg_imp", rimp)
        h.set_train_csr(U, I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q, bu, bi, mu)
        h.set_matrix("svdpp.y", Y)
        loss = h.sgd_epoch(lr, ru, ri, rb, 1)
        gP, gQ, gbu, gbi = h.get_factors()
        gY = h.get_matrix("svdpp.y", (I, k))
        with pytest.raises(capi.LibrecException):
            h.topn(5)
    oP, oQ, oY, obu, obi = P.copy(), Q.copy(), Y.copy(), bu.copy(), bi.copy()
    oloss = O.lib().lro_svdpp_epoch(U, tr.rowptr, tr.col, tr.val, k, oP, oQ, oY, obu, obi, mu, lr, ru, ri, rb, rimp)
    for got, want in ((gP, oP), (gQ, oQ), (gY, oY), (gbu, obu), (gbi, obi)):
        assert np.allclose(got, want, rtol=0, atol=3e-6), float(np.abs(got - want).max())
    assert abs(loss - oloss) <= 2e-5 * abs(oloss)
    untouched = np.setdiff1d(np.arange(I), col)
    assert np.array_equal(gQ[untouched], Q[untouched]) and np.array_equal(gY[untouched], Y[untouched])


def test_svdpp_c1_rmse_within_1e3_of_the_oracle(O, capi, c1):
    """svdpp-test.properties (k=20, lr 0.002, reg 0.01, 100 iterations) on the seeded C1 split; predictions of both factor sets by the
    oracle's restatement of SVD++'s predict() (the implicit-feedback sum is part of it)"""
    tr, te = c1["train"], c1["test"]
    k, lr, reg, iters = 20, 0.002, 0.01, 100
    rng = np.random.default_rng(5)
    sd = 0.001
    P, Q, Y = rng.normal(0, sd, (tr.U, k)), rng.normal(0, sd, (tr.I, k)), rng.normal(0, sd, (tr.I, k))
    bu, bi = rng.normal(0, sd, tr.U), rng.normal(0, sd, tr.I)
    mu = c1["pins"]["global_mean"]
    oP, oQ, oY, obu, obi = P.copy(), Q.copy(), Y.copy(), bu.copy(), bi.copy()
    ol = [O.lib().lro_svdpp_epoch(tr.U, tr.rowptr, tr.col, tr.val, k, oP, oQ, oY, obu, obi, mu, lr, reg, reg, reg, reg) for _ in range(iters)]
    with capi.Handle(capi.MODEL_SVDPP, k) as h:
        h.set_param("svdpp.reg_imp", reg)
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q, bu, bi, mu)
        h.set_matrix("svdpp.y", Y)
        gl = [h.sgd_epoch(lr, reg, reg, reg, it + 1) for it in range(iters)]
        gP, gQ, gbu, gbi = h.get_factors()
        gY = h.get_matrix("svdpp.y", (tr.I, k))
    rows = te.rows().astype(np.int32)

    def rmse_mae(P_, Q_, Y_, bu_, bi_):
        out = np.zeros(te.nnz)
        O.lib().lro_svdpp_predict_pairs(k, P_, Q_, Y_, bu_, bi_, mu, tr.rowptr, tr.col, rows, te.col, te.nnz, out)
        d = te.val - np.clip(out, 1.0, 5.0)
        return float(np.sqrt(np.mean(d * d))), float(np.mean(np.abs(d)))
    g, o = rmse_mae(gP, gQ, gY, gbu, gbi), rmse_mae(oP, oQ, oY, obu, obi)
    print("SVD++ C1: rmse %.6f (oracle %.6f) mae %.6f (oracle %.6f) loss_100 %.2f (oracle %.2f)" % (g[0], o[0], g[1], o[1], gl[-1], ol[-1]))
    assert abs(g[0] - o[0]) < 1e-3 and abs(g[1] - o[1]) < 1e-3
    assert abs(gl[-1] - ol[-1]) < 0.01 * ol[-1] and all(b < a for a, b in zip(gl, gl[1:]))

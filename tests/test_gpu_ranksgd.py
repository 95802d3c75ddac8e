"""GPU: RankSGD (SURVEY.md 8f, row N3 -- recommender/cf/ranking/RankSGDRecommender.java) through the C ABI against the
oracle restatement: the sampler's contract, the epoch arithmetic on the kernel's own samples, and the C1 run of
ranksgd-test.properties."""
import numpy as np
import pytest

from conftest import rng_csr

pytestmark = pytest.mark.gpu


def _skewed_csr(O, U, I, seed):
    """item popularity ~ 1/rank so that popularity sampling is visible; ascending columns per row"""
    rng = np.random.default_rng(seed)
    p = 0.6 / np.arange(1, I + 1) ** 0.7
    mask = rng.random((U, I)) < p[None, :]
    mask[:, I - 3:] = False                                   # three items nobody rated: never drawn (:52-53)
    rows, cols = np.nonzero(mask)
    rowptr = np.zeros(U + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    val = rng.integers(1, 6, rows.shape[0]).astype(np.float64)
    return O.Csr(U, I, np.cumsum(rowptr), cols.astype(np.int32), val)


def test_ranksgd_samples_cover_the_train_entries_and_follow_popularity(O, capi):
    tr = _skewed_csr(O, 600, 200, 3)
    with capi.Handle(capi.MODEL_RANKSGD, 16, seed=5) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        s1 = h.bpr_peek_samples(1, 0, tr.nnz)
        s1b = h.bpr_peek_samples(1, 100, 500)
        s2 = h.bpr_peek_samples(2, 0, tr.nnz)
        with pytest.raises(capi.LibrecException):
            h.bpr_peek_samples(1, tr.nnz - 10, 11)
    u, i, j = s1[:, 0].astype(np.int64), s1[:, 1].astype(np.int64), s1[:, 2].astype(np.int64)
    # one sample per train entry (:66): the (u, i) pairs are a permutation of the CSR entries
    assert np.array_equal(np.sort(u * tr.I + i), np.sort(tr.rows().astype(np.int64) * tr.I + tr.col))
    assert np.array_equal(s1[100:600], s1b)
    assert np.array_equal(s1[:, :2], s2[:, :2]) and not np.array_equal(s1[:, 2], s2[:, 2])   # same entries, fresh negatives
    # negatives: unrated by the user (:86-88), never an item without ratings (:52-53)
    assert j.min() >= 0 and j.max() < tr.I - 3
    member = set((tr.rows().astype(np.int64) * tr.I + tr.col).tolist())
    assert not any(int(k) in member for k in (u * tr.I + j))
    # ... with probability users(j) / numRates (:50), up to the per-user rejection: compare with the exact expectation
    pop = np.bincount(tr.col, minlength=tr.I).astype(np.float64)
    rated = np.zeros((tr.U, tr.I), bool); rated[tr.rows(), tr.col] = True
    w = np.where(rated, 0.0, pop[None, :])
    w /= w.sum(1, keepdims=True)
    expect = (w * np.diff(tr.rowptr)[:, None]).sum(0)                     # expected draws per item
    drawn = np.bincount(j, minlength=tr.I).astype(np.float64)
    big = expect > 20
    chi2 = ((drawn[big] - expect[big]) ** 2 / expect[big]).sum()
    assert chi2 < 1.5 * big.sum(), (chi2, big.sum())


def test_ranksgd_epoch_tracks_oracle_on_its_own_samples(O, capi):
    """feed the kernel's own (u, i, j) draws through the reference arithmetic: same loss, close factors"""
    tr = rng_csr(O, 3000, 2000, 0.01, 3)
    k = 32
    rng = np.random.default_rng(0)
    P = rng.normal(0, 0.1, (tr.U, k)).astype(np.float32).astype(np.float64)
    Q = rng.normal(0, 0.1, (tr.I, k)).astype(np.float32).astype(np.float64)
    with capi.Handle(capi.MODEL_RANKSGD, k, seed=9) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q)
        trip = h.bpr_peek_samples(1, 0, tr.nnz)
        loss = h.sgd_epoch(0.002, 0.0, 0.0, 0.0, 1)
        gP, gQ, _, _ = h.get_factors()
    assert trip[:, 2].min() >= 0
    oP, oQ = P.copy(), Q.copy()
    trip = np.ascontiguousarray(trip.reshape(-1))
    oloss = O.lib().lro_ranksgd_epoch(tr.U, tr.I, tr.rowptr, tr.col, tr.val, k, oP, oQ, 0.002, tr.nnz, trip.ctypes.data, None)
    assert abs(loss - oloss) < 2e-3 * oloss, (loss, oloss)
    step_p, step_q = np.abs(oP - P).max(), np.abs(oQ - Q).max()
    assert step_p > 1e-3 and step_q > 1e-3
    assert np.abs(gP - oP).max() < 0.1 * step_p + 1e-6, (np.abs(gP - oP).max(), step_p)
    assert np.abs(gQ - oQ).max() < 0.1 * step_q + 1e-6, (np.abs(gQ - oQ).max(), step_q)


def test_ranksgd_c1_learns_like_the_oracle(O, capi, c1):
    """ranksgd-test.properties (lr 0.01, 30 iterations, 10 factors) on the C1 split.  The oracle's sequential run reaches
    loss 548.6 k -> 286.3 k and Precision@10 0.176 in CSR order (0.209-0.214 when the same arithmetic visits the entries in
    a shuffled order); the RNG streams and the visiting order differ, so the comparison is at the level of the loss curve
    and the ranking quality.  r01 on a B200: loss_30 286.3 k, Precision@10 0.2115."""
    tr, te = c1["train"], c1["test"]
    O.lib().lro_rng_set_state(*c1["rng_state"])
    P, Q, _, _ = O.mf_setup(tr.U, tr.I, 10, False)
    oP, oQ = P.copy(), Q.copy()
    olosses = O.train(O.RANKSGD, tr, 10, oP, oQ, None, None, 0.0, 0.01, 0.01, 0.01, 0.01, 0.0, 30)
    olosses = olosses[1] if isinstance(olosses, tuple) else olosses
    with capi.Handle(capi.MODEL_RANKSGD, 10, seed=1) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q)
        losses = [h.sgd_epoch(0.01, 0.01, 0.01, 0.0, it + 1) for it in range(30)]
        users = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)
        items, scores, counts = h.topn(10, users=users)
        gP, gQ, _, _ = h.get_factors()

    def precision(lists, cnt):
        hits = sum(np.intersect1d(lists[r, :cnt[r]], te.col[te.rowptr[u]:te.rowptr[u + 1]]).shape[0] for r, u in enumerate(users))
        return hits / (10.0 * users.shape[0])
    oi, _, oc = O.recommend_rank(O.BPR, tr.U, tr.I, 10, oP, oQ, None, None, 0.0, tr, 10, users=users)
    p_gpu, p_ora = precision(items, counts), precision(oi, oc)
    print("RankSGD C1: loss_1 %.1f (oracle %.1f) loss_30 %.1f (oracle %.1f)  P@10 %.4f (oracle %.4f)" % (
        losses[0], olosses[0], losses[-1], olosses[29], p_gpu, p_ora))
    assert abs(losses[0] - olosses[0]) < 0.01 * olosses[0]              # first epoch: factors still ~0, error = -r
    assert losses[-1] < 0.75 * losses[0] and abs(losses[-1] - olosses[29]) < 0.05 * olosses[29]
    assert p_gpu > 0.10 and p_gpu > 0.7 * p_ora
    # the lists the device returns are the reference's lists for the factors it learned
    ei, es, ec = O.recommend_rank(O.BPR, tr.U, tr.I, 10, gP, gQ, None, None, 0.0, tr, 10, users=users)
    assert np.array_equal(items, ei) and np.array_equal(counts, ec) and np.array_equal(scores.view(np.int64), es.view(np.int64))

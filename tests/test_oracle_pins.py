"""CPU: pins of the oracle itself (JDK known answers, reference fixtures, hand-worked updates)."""
import ctypes as C
import math
import os

import numpy as np

from conftest import GOLDEN, rng_csr


# ---- java.util.Random (JDK 8) published known answers -------------------------------------------
def test_java_random_known_answers(O):
    L = O.lib()
    L.lro_seed(42); assert L.lro_next_int() == -1170105035
    L.lro_seed(0); assert L.lro_next_int() == -1155484576
    L.lro_seed(0); assert [L.lro_uniform_int(100) for _ in range(10)] == [60, 48, 29, 47, 15, 53, 91, 61, 19, 54]
    L.lro_seed(0); assert L.lro_uniform() == 0.730967787376657
    L.lro_seed(42); assert L.lro_uniform() == 0.7275636800328681
    L.lro_seed(0); assert L.lro_next_gaussian() == 0.8025330637390305
    L.lro_seed(42); assert L.lro_next_gaussian() == 1.1419053154730547


def test_next_int_power_of_two_and_rejection(O):
    L = O.lib()
    # independent pure-Python java.util.Random
    class JR:
        def __init__(s, seed): s.s = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)
        def next(s, bits):
            s.s = (s.s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
            v = s.s >> (48 - bits)
            return v - (1 << 32) if v >= (1 << 31) else v
        def next_int(s, bound):
            r = s.next(31); m = bound - 1
            if bound & m == 0: return (bound * r) >> 31
            u = r
            while True:
                r = u % bound
                t = (u - r + m) & 0xFFFFFFFF
                if t < (1 << 31): return r
                u = s.next(31)
    for seed in (1, 7, 123456789):
        for bound in (1, 2, 16, 943, 1682, 1 << 20, (1 << 30) + 12345, 2147483647):
            L.lro_seed(seed); j = JR(seed)
            assert [L.lro_uniform_int(bound) for _ in range(200)] == [j.next_int(bound) for _ in range(200)]


def test_fdlibm_log_is_within_one_ulp(O):
    L = O.lib()
    xs = np.random.default_rng(0).random(20000) * 0.999 + 1e-9
    for x in xs[:5000]:
        a, b = L.lro_strictmath_log(float(x)), math.log(float(x))
        assert abs(a - b) <= 2 * np.spacing(abs(b))
    assert L.lro_strictmath_log(1.0) == 0.0


def test_float_promotion(O):
    L = O.lib()
    # Float.valueOf then widening
    assert L.lro_float_promote(b"0.002") == 0.0020000000949949026
    assert L.lro_float_promote(b"0.01") == 0.009999999776482582
    assert L.lro_float_promote(b"0.08") == 0.07999999821186066


# ---- loader / splitter against the reference's own test expectations -----------------------------
def test_loader_matrix4by4_has_13_entries(O):
    # data/model/TextDataModelTestCase.java:66  assertEquals(getDataSize, 13)
    m = O.load_text(os.path.join(GOLDEN, "matrix4by4.txt"))
    assert (m.U, m.I, m.nnz) == (4, 4, 13)
    # first-seen ids: user "1" -> 0 ...; rows sorted by column
    assert m.rowptr.tolist() == [0, 4, 7, 10, 13]
    assert m.col.tolist() == [0, 1, 2, 3, 1, 2, 3, 0, 1, 3, 0, 1, 2]
    assert m.val.tolist() == [float(v) for v in range(1, 14)]


def test_loader_reference_fixtures_uirt_csv_and_directory(O):
    """the other expectations of the reference's loader test, on its own fixture files (copied under tests/golden/):
    data/model/TextDataModelTestCase.java:84 (UIRT, 13), :101 (directory with sub-directories, 26), :118 (CSV, 13)"""
    g = os.path.join(GOLDEN, "datamodeltest")
    m = O.load_text(os.path.join(g, "matrix4by4-date.txt"), column_format="UIRT")
    assert (m.U, m.I, m.nnz) == (4, 4, 13)
    c = O.load_text(os.path.join(g, "testCSV.txt"))                       # ',', ' ' and tab mixed in one file
    assert (c.U, c.I, c.nnz) == (4, 4, 13)
    d = O.load_text(os.path.join(g, "test-convert-dir"))                  # 3 files, one of them twice -> duplicates collapse
    assert d.nnz == 26
    two = O.load_text(os.path.join(g, "testCSV.txt") + ":" + os.path.join(g, "test-convert-dir", "subdir1"))
    assert two.nnz == 26                                                  # data.input.path is ':'-separated (TextDataModel.java:60)


def test_loader_duplicate_keeps_earliest_and_blank_line_stops(O, tmp_path):
    p = tmp_path / "r.txt"
    p.write_text("a x 1\nb y 2\na x 5\nb,x;3\n\nc z 9\n")
    m = O.load_text(str(p))
    assert (m.U, m.I, m.nnz) == (2, 2, 3)       # "c z 9" is after the blank line (TextDataConvertor.java:176-178)
    assert m.val.tolist() == [1.0, 3.0, 2.0]     # (a,x)=1 earliest wins (DataFrame.java:244-255); row b: x=3, y=2
    b = O.load_text(str(p), 0.0)                 # binarize threshold 0.0 (bpr-test.properties)
    assert b.val.tolist() == [1.0, 1.0, 1.0]


def test_splitter_ratio_bound(O, c1):
    # data/splitter/RatioDataSplitterTestCase.java:74  |ratio - 0.8| <= 0.01
    tr, te, full = c1["train"], c1["test"], c1["full"]
    assert tr.nnz + te.nnz == full.nnz == 100000
    assert abs(tr.nnz / full.nnz - 0.8) <= 0.01
    assert (tr.nnz, te.nnz) == (c1["pins"]["train_nnz"], c1["pins"]["test_nnz"])
    # replay: java.util.Random(1), one nextDouble per entry in CSR order
    L = O.lib()
    L.lro_seed(1)
    flags = np.zeros(full.nnz, np.uint8)
    L.lro_split_ratio(full.nnz, full.val, 0.8, flags)
    assert int((flags == 1).sum()) == tr.nnz


def test_other_splitters_meet_the_reference_test_expectations(O, c1):
    """sizes asserted by the reference's splitter tests on its 4x4 fixtures: GivenNDataSplitterTestCase.java:70-71,90-91
    (N=1: 4 train / 9 test), LOOCVDataSplitterTestCase.java:68-69,86-87 (9 / 4), KCVDataSplitterTestCase.java:68-69
    (matrix4by4A.txt, 6 folds: 10 / 2 each), RatioDataSplitterTestCase.java:93,112 (user / item ratio within 0.01 of 0.8)"""
    m = O.load_text(os.path.join(GOLDEN, "matrix4by4.txt"))
    for seed in range(1, 6):
        for by in ("user", "item"):
            O.lib().lro_seed(seed)
            tr, te = O.split(m, "givenn", by, n_given=1)
            assert (tr.nnz, te.nnz) == (4, 9)
            if by == "user":
                assert np.diff(tr.rowptr).tolist() == [1, 1, 1, 1]           # exactly N given entries per user
            O.lib().lro_seed(seed)
            tr, te = O.split(m, "loocv", by)
            assert (tr.nnz, te.nnz) == (9, 4)
    a = O.load_text(os.path.join(GOLDEN, "datamodeltest", "matrix4by4A.txt"))
    O.lib().lro_seed(1)
    folds = O.split(a, "kcv", k_fold=6)
    assert [(x.nnz, y.nnz) for x, y in folds] == [(10, 2)] * 6
    full = c1["full"]
    for by in ("user", "item"):
        O.lib().lro_seed(1)
        tr, te = O.split(full, "ratio", by, ratio=0.8)
        assert abs(tr.nnz / float(full.nnz) - 0.8) <= 0.01 and tr.nnz + te.nnz == full.nnz
    O.lib().lro_seed(1)
    tr_u, _ = O.split(full, "ratio", "user", ratio=0.8)
    assert np.array_equal(tr_u.col, c1["train"].col)                           # getRatioByUser == getRatioByRating draw for draw


def test_date_splitters_meet_the_reference_test_expectations(O):
    """LOOCVDataSplitterTestCase.java:103-104,120-121 (9 / 4), GivenNDataSplitterTestCase.java:109-110,128-129 (4 / 9) on
    matrix4by4-date.txt; RatioDataSplitterTestCase.java:153,171,189 on ratings-date.txt with that test's tolerances"""
    g = os.path.join(GOLDEN, "datamodeltest")
    m = O.load_text(os.path.join(g, "matrix4by4-date.txt"), column_format="UIRT")
    assert m.date.tolist() == list(range(1, 14))
    for by in ("userdate", "itemdate"):
        tr, te = O.split(m, "loocv", by)
        assert (tr.nnz, te.nnz) == (9, 4)
        tr, te = O.split(m, "givenn", by, n_given=1)
        assert (tr.nnz, te.nnz) == (4, 9)
    r = O.load_text(os.path.join(g, "ratings-date.txt"), column_format="UIRT")
    assert (r.U, r.I, r.nnz) == (1508, 2071, 35492)                          # 5 of the 35 497 lines repeat a (user, item) pair
    for by, tol in (("ratingdate", 0.01), ("userdate", 0.02), ("itemdate", 0.04)):
        tr, te = O.split(r, "ratio", by, ratio=0.8)
        assert abs(tr.nnz / float(r.nnz) - 0.8) <= tol and tr.nnz + te.nnz == r.nnz


def test_matrix_setup(O, c1):
    mu, mn, mx = O.matrix_setup(c1["train"])
    assert mu == c1["pins"]["global_mean"] and (mn, mx) == (1.0, 5.0)
    assert abs(mu - c1["train"].val.mean()) < 1e-12
    one = O.Csr(1, 2, [0, 2], [0, 1], [3.0, 3.0])
    assert O.matrix_setup(one)[1:] == (0.0, 3.0)     # minRate = 0 when min == max (MatrixRecommender.java:105-107)


# ---- hand-worked single updates -----------------------------------------------------------------
def _f32(x):
    return float(np.float32(x))


def test_biasedmf_single_update_by_hand(O):
    # BiasedMFRecommender.java:77-98 worked by hand for one rating, k = 2
    L = O.lib()
    tr = O.Csr(1, 1, [0, 1], [0], [4.0])
    P = np.array([[0.1, -0.2]]); Q = np.array([[0.3, 0.5]]); bu = np.array([0.05]); bi = np.array([-0.02]); mu = 3.0
    lr, ru, ri, rb = _f32(0.01), _f32(0.02), _f32(0.03), 0.04
    pred = (0.0 + 0.3 * 0.1) + 0.5 * -0.2
    pred = pred + 0.05 + -0.02 + mu
    e = 4.0 - pred
    loss = e * e
    loss += rb * 0.05 * 0.05
    nbu = 0.05 + lr * (e - rb * 0.05)
    loss += rb * -0.02 * -0.02
    nbi = -0.02 + lr * (e - rb * -0.02)
    nP, nQ = P.copy(), Q.copy()
    for f in range(2):
        uf, itf = P[0, f], Q[0, f]
        nP[0, f] = uf + lr * (e * itf - ru * uf)
        nQ[0, f] = itf + lr * (e * uf - ri * itf)
        loss += ru * uf * uf + ri * itf * itf
    loss *= 0.5
    got = L.lro_biasedmf_epoch(1, tr.rowptr, tr.col, tr.val, 2, P, Q, bu, bi, mu, 0.01, 0.02, 0.03, rb, None, None)
    assert got == loss
    assert P.tolist() == nP.tolist() and Q.tolist() == nQ.tolist()
    assert bu[0] == nbu and bi[0] == nbi


def test_pmf_single_update_by_hand(O):
    L = O.lib()
    tr = O.Csr(1, 1, [0, 1], [0], [2.0])
    P = np.array([[0.5, 0.25]]); Q = np.array([[-1.0, 2.0]])
    lr, ru, ri = _f32(0.01), _f32(0.08), _f32(0.08)
    e = 2.0 - ((0.0 + -1.0 * 0.5) + 2.0 * 0.25)
    loss = e * e
    nP, nQ = P.copy(), Q.copy()
    for f in range(2):
        uf, itf = P[0, f], Q[0, f]
        nP[0, f] = uf + lr * (e * itf - ru * uf)
        nQ[0, f] = itf + lr * (e * uf - ri * itf)
        loss += ru * uf * uf + ri * itf * itf
    got = L.lro_pmf_epoch(1, tr.rowptr, tr.col, tr.val, 2, P, Q, 0.01, 0.08, 0.08, None, None)
    assert got == 0.5 * loss and P.tolist() == nP.tolist() and Q.tolist() == nQ.tolist()


def test_bpr_single_update_by_hand(O):
    L = O.lib()
    tr = O.Csr(1, 2, [0, 1], [0], [1.0])
    P = np.array([[0.2, -0.1]]); Q = np.array([[0.4, 0.3], [-0.5, 0.6]])
    trip = np.array([0, 0, 1], np.int32)
    lr, ru, ri = _f32(0.01), _f32(0.01), _f32(0.01)
    pos = (0.0 + 0.4 * 0.2) + 0.3 * -0.1
    neg = (0.0 + -0.5 * 0.2) + 0.6 * -0.1
    x = pos - neg
    loss = -math.log(1.0 / (1.0 + math.exp(-x)))
    d = 1.0 / (1.0 + math.exp(x))
    nP, nQ = P.copy(), Q.copy()
    for f in range(2):
        uf, pf, nf = P[0, f], Q[0, f], Q[1, f]
        nP[0, f] = uf + lr * (d * (pf - nf) - ru * uf)
        nQ[0, f] = pf + lr * (d * uf - ri * pf)
        nQ[1, f] = nf + lr * (d * (-uf) - ri * nf)
        loss += ru * uf * uf + ri * pf * pf + ri * nf * nf
    got = L.lro_bpr_epoch(1, 2, tr.rowptr, tr.col, 2, P, Q, 0.01, 0.01, 0.01, 1, trip.ctypes.data, None)
    assert abs(got - loss) < 1e-15
    assert np.allclose(P, nP, rtol=0, atol=1e-17) and np.allclose(Q, nQ, rtol=0, atol=1e-17)


def test_bpr_sampler_respects_rows(O):
    tr = rng_csr(O, 30, 17, 0.3, 5, values=(1.0,))
    tr.rowptr[:] = tr.rowptr  # keep
    P = np.zeros((30, 4)); Q = np.zeros((17, 4))
    out = np.zeros(3 * 2000, np.int32)
    O.lib().lro_seed(3)
    O.lib().lro_bpr_epoch(30, 17, tr.rowptr, tr.col, 4, P, Q, 0.01, 0.01, 0.01, 2000, None, out.ctypes.data)
    t = out.reshape(-1, 3)
    for u, i, j in t[:500]:
        row = tr.col[tr.rowptr[u]:tr.rowptr[u + 1]]
        assert i in row and j not in row


# ---- RankSGD: recommender/cf/ranking/RankSGDRecommender.java ---------------------------
def test_ranksgd_single_update_by_hand(O):
    """:86-103 worked by hand: error = (pos - neg) - (r - 0); no regularisation; the OLD user factor feeds both item updates"""
    L = O.lib()
    tr = O.Csr(1, 2, [0, 1], [0], [4.0])
    P = np.array([[0.2, -0.1]]); Q = np.array([[0.4, 0.3], [-0.5, 0.6]])
    trip = np.array([0, 0, 1], np.int32)
    lr = _f32(0.01)
    pos = (0.0 + 0.2 * 0.4) + -0.1 * 0.3
    neg = (0.0 + 0.2 * -0.5) + -0.1 * 0.6
    err = (pos - neg) - (4.0 - 0.0)
    sgd = lr * err
    nP, nQ = P.copy(), Q.copy()
    for f in range(2):
        uf, pf, nf = P[0, f], Q[0, f], Q[1, f]
        nP[0, f] = uf + -sgd * (pf - nf)
        nQ[0, f] = pf + -sgd * uf
        nQ[1, f] = nf + sgd * uf
    got = L.lro_ranksgd_epoch(1, 2, tr.rowptr, tr.col, tr.val, 2, P, Q, 0.01, 1, trip.ctypes.data, None)
    assert got == 0.5 * err * err
    assert P.tolist() == nP.tolist() and Q.tolist() == nQ.tolist()


def test_ranksgd_item_probs_ascending_with_hashmap_tie_order(O):
    """:47-57: prob = users(j) / numRates, zero-probability items dropped, ascending by prob, ties in HashMap order
    (= ascending item id while ids stay below the table size)"""
    #            item: 0  1  2  3  4     users: 2, 0, 1, 2, 1
    tr = O.Csr(3, 5, [0, 3, 5, 6], [0, 2, 3, 0, 3, 4], np.ones(6))
    items = np.zeros(5, np.int32); probs = np.zeros(5)
    m = O.lib().lro_ranksgd_item_probs(3, 5, tr.rowptr, tr.col, items, probs)
    assert m == 4
    assert items[:m].tolist() == [2, 4, 0, 3]
    assert probs[:m].tolist() == [1 / 6, 1 / 6, 2 / 6, 2 / 6]


def test_ranksgd_reference_sampler_draws_unrated_items_by_popularity(O):
    tr = rng_csr(O, 200, 40, 0.2, 7)
    k = 4
    P = np.zeros((tr.U, k)); Q = np.zeros((tr.I, k))
    out = np.zeros(3 * tr.nnz, np.int32)
    O.lib().lro_seed(11)
    loss = O.lib().lro_ranksgd_epoch(tr.U, tr.I, tr.rowptr, tr.col, tr.val, k, P, Q, 0.0, 0, None, out.ctypes.data)
    t = out.reshape(-1, 3)
    assert np.array_equal(t[:, 0], tr.rows()) and np.array_equal(t[:, 1], tr.col)        # every train entry, CSR order
    for u, i, j in t[:800]:
        assert j not in tr.col[tr.rowptr[u]:tr.rowptr[u + 1]]
    assert loss == 0.5 * float(np.sum(tr.val ** 2))                                       # zero factors: error = -r
    # negatives follow item popularity (before the per-user rejection): the most popular quarter of the catalogue is
    # drawn more often than the least popular quarter
    pop = np.bincount(tr.col, minlength=tr.I)
    order = np.argsort(pop)
    drawn = np.bincount(t[:, 2], minlength=tr.I)
    assert drawn[order[-10:]].sum() > 1.3 * drawn[order[:10]].sum()
    assert drawn[pop == 0].sum() == 0


# ---- SVD++ (oracle groundwork): recommender/cf/rating/SVDPlusPlusRecommender.java ------------
def test_svdpp_epoch_by_hand(O):
    """one user with two ratings, k = 2, worked by hand from :62-123: fv is fixed for the row, the second rating sees the
    updated p_u / b_u, the item gradient uses (p + fv), steps accumulate e * q_old * scale, Y moves after the row"""
    L = O.lib()
    tr = O.Csr(1, 3, [0, 2], [0, 2], [4.0, 2.0])
    P = np.array([[0.1, -0.2]]); Q = np.array([[0.3, 0.1], [0.9, 0.9], [-0.2, 0.4]])
    Y = np.array([[0.05, 0.02], [0.7, 0.7], [-0.01, 0.03]])
    bu = np.array([0.1]); bi = np.array([0.2, 0.5, -0.1]); mu = 3.0
    lr, ru, ri = _f32(0.01), _f32(0.02), _f32(0.03)
    rb, rimp = 0.04, 0.015
    p, q, y, b_u, b_i = P.copy(), Q.copy(), Y.copy(), bu.copy(), bi.copy()
    n = 2
    scale = math.pow(n, -0.5)
    fv = [(0.0 + y[0, f]) + y[2, f] for f in range(2)]          # value = row(j).get(f) + value, entries in item order
    fv = [v * scale for v in fv]
    steps = [0.0, 0.0]
    loss = 0.0
    for i, r in ((0, 4.0), (2, 2.0)):
        pred = b_u[0] + b_i[i] + mu
        for f in range(2):
            pred += (fv[f] + p[0, f]) * q[i, f]
        e = r - pred
        loss += e * e
        ub = b_u[0]; b_u[0] += lr * (e - rb * ub); loss += rb * ub * ub
        ib = b_i[i]; b_i[i] += lr * (e - rb * ib); loss += rb * ib * ib
        for f in range(2):
            uf, qf = p[0, f], q[i, f]
            p[0, f] += lr * (e * qf - ru * uf)
            q[i, f] += lr * (e * (uf + fv[f]) - ri * qf)
            loss += ru * uf * uf + ri * qf * qf
            steps[f] += e * qf * scale
    for j in (0, 2):
        for f in range(2):
            fac = y[j, f]
            y[j, f] += lr * (steps[f] - rimp * fac * n)
            loss += rimp * fac * fac * n
    got = L.lro_svdpp_epoch(1, tr.rowptr, tr.col, tr.val, 2, P, Q, Y, bu, bi, mu, 0.01, 0.02, 0.03, rb, rimp)
    assert got == 0.5 * loss
    assert P.tolist() == p.tolist() and Q.tolist() == q.tolist() and Y.tolist() == y.tolist()
    assert bu.tolist() == b_u.tolist() and bi.tolist() == b_i.tolist()
    assert Q[1].tolist() == [0.9, 0.9] and Y[1].tolist() == [0.7, 0.7]      # the unrated item is untouched
    # predict (:138-150) divides by Math.sqrt(n)
    out = np.zeros(1)
    L.lro_svdpp_predict_pairs(2, P, Q, Y, bu, bi, mu, tr.rowptr, tr.col, np.array([0], np.int32), np.array([1], np.int32), 1, out)
    fvp = [((0.0 + Y[0, f]) + Y[2, f]) / math.sqrt(2.0) for f in range(2)]
    exp = bu[0] + bi[1] + mu
    for f in range(2):
        exp += (fvp[f] + P[0, f]) * Q[1, f]
    assert out[0] == exp


def test_svdpp_learns_on_c1(O, c1):
    """svdpp-test.properties-like run on the C1 split (lr 0.01, reg 0.1 / 0.015, 10 factors, 13 iterations): the restated loop
    converges and beats the global-mean predictor clearly"""
    tr, te = c1["train"], c1["test"]
    k = 10
    rng = np.random.default_rng(3)
    P = rng.normal(0, 0.001, (tr.U, k)); Q = rng.normal(0, 0.001, (tr.I, k)); Y = rng.normal(0, 0.001, (tr.I, k))
    bu = rng.normal(0, 0.001, tr.U); bi = rng.normal(0, 0.001, tr.I)
    mu = c1["pins"]["global_mean"]
    losses = [O.lib().lro_svdpp_epoch(tr.U, tr.rowptr, tr.col, tr.val, k, P, Q, Y, bu, bi, mu, 0.01, 0.1, 0.1, 0.1, 0.015) for _ in range(13)]
    assert all(b < a for a, b in zip(losses, losses[1:]))
    rows = te.rows().astype(np.int32)
    out = np.zeros(te.nnz)
    O.lib().lro_svdpp_predict_pairs(k, P, Q, Y, bu, bi, mu, tr.rowptr, tr.col, rows, te.col, te.nnz, out)
    rmse = float(np.sqrt(np.mean((te.val - np.clip(out, 1.0, 5.0)) ** 2)))
    base = float(np.sqrt(np.mean((te.val - mu) ** 2)))
    assert rmse < 0.97 and rmse < base - 0.15, (rmse, base)


# ---- learning-rate schedule / convergence (host logic the shim keeps in Java) -----------------------
def test_update_lrate_and_is_converged(O):
    L = O.lib()
    last = C.c_double(10.0)
    lr = L.lro_update_lrate(0.01, 1000.0, 2, 1, 1.0, 9.0, C.byref(last))      # bold driver, loss went down
    assert lr == _f32(np.float32(0.01) * np.float32(1.05)) and last.value == 9.0
    lr = L.lro_update_lrate(0.01, 1000.0, 2, 1, 1.0, 11.0, C.byref(last))     # loss went up
    assert lr == _f32(np.float32(0.01) * np.float32(0.5))
    lr = L.lro_update_lrate(0.01, 1000.0, 1, 1, 0.9, 5.0, C.byref(last))      # iter 1 falls through to decay
    assert lr == _f32(np.float32(0.01) * np.float32(0.9))
    lr = L.lro_update_lrate(0.5, 0.01, 3, 0, 1.0, 5.0, C.byref(last))         # clamp to max
    assert lr == _f32(0.01)
    d = C.c_float()
    assert L.lro_is_converged(1.0, 1.0 - 1e-6, C.byref(d)) == 1
    assert L.lro_is_converged(1.0, 0.5, C.byref(d)) == 0 and d.value == 0.5
    assert L.lro_is_converged(1.0, float("nan"), None) == -1
    assert L.lro_is_converged(1.0, float("inf"), None) == -1


# ---- java.util.PriorityQueue + stable sort ------------------------------------------------------
def _py_topk(values, k):
    """independent replay of Lists.sortKeyValueListTopK(inverse=true) incl. JDK-8 siftUp/siftDown"""
    import struct

    def cmp(a, b):
        if a < b: return -1
        if a > b: return 1
        x = struct.unpack("<q", struct.pack("<d", a))[0]; y = struct.unpack("<q", struct.pack("<d", b))[0]
        return (x > y) - (x < y)
    n = len(values); kk = min(k, n)
    if kk == 0: return [], []
    q = []

    def sift_up(pos, x):
        while pos > 0:
            par = (pos - 1) >> 1
            if cmp(x[1], q[par][1]) >= 0: break
            q[pos] = q[par]; pos = par
        q[pos] = x

    def sift_down(pos, x, size):
        half = size >> 1
        while pos < half:
            ch = 2 * pos + 1; r = ch + 1
            if r < size and cmp(q[ch][1], q[r][1]) > 0: ch = r
            if cmp(x[1], q[ch][1]) <= 0: break
            q[pos] = q[ch]; pos = ch
        q[pos] = x
    for i in range(kk):
        q.append(None); sift_up(i, (i, values[i]))
    for i in range(kk, n):
        if cmp(values[i], q[0][1]) > 0:
            last = q.pop()
            if q: sift_down(0, last, len(q))
            q.append(None); sift_up(len(q) - 1, (i, values[i]))
    heap_keys = [e[0] for e in q]
    import functools
    out = sorted(q, key=functools.cmp_to_key(lambda a, b: -cmp(a[1], b[1])))   # sorted() is stable
    return heap_keys, [e[0] for e in out]


def test_priority_queue_heap_order_with_ties(O):
    L = O.lib()
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(1, 60)); k = int(rng.integers(1, 14))
        vals = rng.integers(0, 6, n).astype(np.float64) / 2.0          # many ties
        if trial % 5 == 0: vals[rng.integers(0, n)] = -0.0
        keys = np.zeros(k, np.int32)
        m = L.lro_heap_trace(vals, n, k, keys)
        hk, _ = _py_topk(vals.tolist(), k)
        assert keys[:m].tolist() == hk


def test_recommend_rank_matches_replay_and_naive(O):
    rng = np.random.default_rng(2)
    U, I, k, N = 23, 57, 5, 10
    P = rng.normal(size=(U, k)); Q = rng.normal(size=(I, k))
    Q[10] = Q[3]; Q[40] = Q[3]                     # exact ties
    bu = rng.normal(size=U); bi = rng.normal(size=I); bi[10] = bi[3]; bi[40] = bi[3]
    tr = rng_csr(O, U, I, 0.2, 9)
    items, scores, counts = O.recommend_rank(O.BIASEDMF, U, I, k, P, Q, bu, bi, 3.5, tr, N, nthreads=2)
    for u in range(U):
        row = set(tr.col[tr.rowptr[u]:tr.rowptr[u + 1]].tolist())
        cand = [i for i in range(I) if i not in row]
        vals = []
        for i in cand:
            d = 0.0
            for f in range(k): d += Q[i, f] * P[u, f]
            vals.append(d + bu[u] + bi[i] + 3.5)
        _, order = _py_topk(vals, N)
        assert items[u, :counts[u]].tolist() == [cand[t] for t in order]
        assert scores[u, :counts[u]].tolist() == [vals[t] for t in order]
        assert counts[u] == min(N, len(cand))
        # value multiset equals a naive descending sort
        assert sorted(vals, reverse=True)[:N] == scores[u, :counts[u]].tolist()


def test_recommend_rank_edge_cases(O):
    U, I, k = 3, 4, 2
    P = np.ones((U, k)); Q = np.arange(I * k, dtype=np.float64).reshape(I, k)
    Q[2, 0] = np.nan
    full = O.Csr(U, I, [0, 4, 4, 6], [0, 1, 2, 3, 0, 3], np.ones(6))
    items, scores, counts = O.recommend_rank(O.PMF, U, I, k, P, Q, None, None, 0.0, full, 10)
    assert counts.tolist() == [0, 3, 1]             # user 0 trained on everything; NaN item dropped
    assert items[1, :3].tolist() == [3, 1, 0] and items[2, 0] == 1 and items[0, 0] == -1


def test_c1_regression_pins(O, c1):
    """full config C1 through the oracle == the numbers committed with the fixture"""
    pins = c1["pins"]; tr, te = c1["train"], c1["test"]
    O.lib().lro_rng_set_state(*c1["rng_state"])
    P, Q, bu, bi = O.mf_setup(tr.U, tr.I, 20, True)
    assert P[0, 0] == pins["biasedmf_P00"] and bi[-1] == pins["biasedmf_bi_last"]
    done, losses = O.train(O.BIASEDMF, tr, 20, P, Q, bu, bi, pins["global_mean"], 0.002, 0.01, 0.01, 0.01, 0.01, 100)
    rmse, mae = O.eval_rating(O.BIASEDMF, te, 20, P, Q, bu, bi, pins["global_mean"], 1.0, 5.0)
    assert done == 100
    assert abs(losses[0] - pins["biasedmf"]["loss_1"]) < 1e-6 and abs(losses[-1] - pins["biasedmf"]["loss_100"]) < 1e-6
    assert abs(rmse - pins["biasedmf"]["rmse"]) < 1e-12 and abs(mae - pins["biasedmf"]["mae"]) < 1e-12
    assert 0.90 < rmse < 0.96 and 0.71 < mae < 0.76     # sanity: the range LibRec documents for BiasedMF on ml-100k


def test_c1_ranksgd_regression_pins(O, c1):
    """ranksgd-test.properties on C1 through the oracle (reference order, java.util.Random negatives) == the committed pins"""
    pins = c1["pins"]["ranksgd"]; tr, te = c1["train"], c1["test"]
    O.lib().lro_rng_set_state(*c1["rng_state"])
    P, Q, _, _ = O.mf_setup(tr.U, tr.I, 10, False)
    done, losses = O.train(O.RANKSGD, tr, 10, P, Q, None, None, 0.0, 0.01, 0.01, 0.01, 0.01, 0.0, 30)
    assert done == pins["iters"] == 30
    assert abs(losses[0] - pins["loss_1"]) < 1e-6 and abs(losses[-1] - pins["loss_30"]) < 1e-6
    users = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)
    items, _, counts = O.recommend_rank(O.BPR, tr.U, tr.I, 10, P, Q, None, None, 0.0, tr, 10, users=users)
    hits = sum(np.intersect1d(items[r, :counts[r]], te.col[te.rowptr[u]:te.rowptr[u + 1]]).shape[0] for r, u in enumerate(users))
    assert abs(hits / (10.0 * users.shape[0]) - pins["precision_at_10"]) < 1e-12

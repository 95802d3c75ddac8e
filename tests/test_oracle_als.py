"""CPU: the oracle's WRMF / eALS restatements against independent pure-Python replays written from
recommender/cf/ranking/WRMFRecommender.java:74-166, EALSRecommender.java:114-214 and math/structure/DenseMatrix.java:362-437
(Python floats are IEEE doubles without FMA contraction, like the JVM's).  Both models are deterministic, so the bar is BIT equality."""
import math

import numpy as np

from conftest import rng_csr


def java_inverse(a):
    """DenseMatrix.inverse(): Gauss-Jordan, partial pivoting, lists of lists"""
    n = len(a)
    inv = [[1.0 if r == c else 0.0 for c in range(n)] for r in range(n)]
    if n == 1:
        inv[0][0] = 1.0 / a[0][0]
        return inv
    m = [row[:] for row in a]
    for r in range(n):
        mag, pivot = 0.0, -1
        for j in range(r, n):
            mag2 = abs(m[j][r])
            if mag2 > mag:
                mag, pivot = mag2, j
        if pivot == -1 or mag == 0:
            return inv
        if pivot != r:
            for c in range(r, n):
                m[r][c], m[pivot][c] = m[pivot][c], m[r][c]
            for c in range(n):
                inv[r][c], inv[pivot][c] = inv[pivot][c], inv[r][c]
        mag = m[r][r]
        for c in range(r, n):
            m[r][c] = m[r][c] / mag
        for c in range(n):
            inv[r][c] = inv[r][c] / mag
        for r2 in range(n):
            if r2 == r:
                continue
            mag2 = m[r2][r]
            for c in range(r, n):
                m[r2][c] = m[r2][c] - mag2 * m[r][c]
            for c in range(n):
                inv[r2][c] = inv[r2][c] - mag2 * inv[r][c]
    return inv


def gram(M, k):
    out = [[0.0] * k for _ in range(k)]
    for r in range(k):
        for c in range(k):
            v = 0.0
            for row in M:
                v += row[c] * row[r]
            out[r][c] = v
    return out


def wrmf_replay(rows, cols, k, X, Y, reg_u, reg_i):
    """rows[u] = [(item, weight)], cols[i] = [(user, weight)]; X, Y lists of lists, updated in place"""
    def side(entries, F, OUT, reg):
        G = gram(F, k)
        for r, ent in enumerate(entries):
            b = [0.0] * k
            for idx, w in ent:
                weight = w + 1.0
                for f in range(k):
                    b[f] += F[idx][f] * weight
            A = [[G[x][y] + reg for y in range(k)] for x in range(k)]
            for idx, w in ent:
                for x in range(k):
                    temp = F[idx][x] * w
                    for y in range(k):
                        A[x][y] += temp * F[idx][y]
            W = java_inverse(A)
            new = []
            for x in range(k):
                v = 0.0
                for y in range(k):
                    v += b[y] * W[x][y]
                new.append(v)
            OUT[r] = new
    side(rows, Y, X, reg_u)
    side(cols, X, Y, reg_i)


def eals_replay(rows, cols, k, P, Q, conf, reg_u, reg_i):
    U, I = len(rows), len(cols)
    Sq = [[0.0] * k for _ in range(k)]
    for f1 in range(k):
        for f2 in range(f1 + 1):
            v = 0.0
            for i in range(I):
                v += conf[i] * Q[i][f1] * Q[i][f2]
            Sq[f1][f2] = v
            Sq[f2][f1] = v
    ipred = [0.0] * I
    for u in range(U):
        for i, _ in rows[u]:
            d = 0.0
            for f in range(k):
                d += Q[i][f] * P[u][f]
            ipred[i] = d
        for f in range(k):
            numer, denom = 0.0, reg_u + Sq[f][f]
            for f2 in range(k):
                if f2 != f:
                    numer -= P[u][f2] * Sq[f][f2]
            for i, w in rows[u]:
                ipred[i] -= P[u][f] * Q[i][f]
                numer += (w - (w - conf[i]) * ipred[i]) * Q[i][f]
                denom += (w - conf[i]) * Q[i][f] * Q[i][f]
            P[u][f] = numer / denom
            for i, _ in rows[u]:
                ipred[i] += P[u][f] * Q[i][f]
    Sp = gram(P, k)
    upred = [0.0] * U
    for i in range(I):
        for u, _ in cols[i]:
            d = 0.0
            for f in range(k):
                d += Q[i][f] * P[u][f]
            upred[u] = d
        for f in range(k):
            numer, denom = 0.0, conf[i] * Sp[f][f] + reg_i
            for f2 in range(k):
                if f2 != f:
                    numer -= Q[i][f2] * Sp[f2][f]
            numer *= conf[i]
            for u, w in cols[i]:
                upred[u] -= P[u][f] * Q[i][f]
                numer += (w - (w - conf[i]) * upred[u]) * P[u][f]
                denom += (w - conf[i]) * P[u][f] * P[u][f]
            Q[i][f] = numer / denom
            for u, _ in cols[i]:
                upred[u] += P[u][f] * Q[i][f]


def _lists(tr, val):
    rows = [[(int(tr.col[e]), float(val[e])) for e in range(tr.rowptr[u], tr.rowptr[u + 1])] for u in range(tr.U)]
    cols = [[] for _ in range(tr.I)]
    for u in range(tr.U):
        for i, w in rows[u]:
            cols[i].append((u, w))
    return rows, cols


def test_dense_inverse_is_the_reference_gauss_jordan(O):
    rng = np.random.default_rng(5)
    for n in (1, 2, 5, 12):
        a = rng.standard_normal((n, n)) + np.eye(n) * 0.1
        a[0, 0] = 1e-9 if n > 1 else a[0, 0]           # forces a row swap
        inv = np.zeros((n, n))
        O.lib().lro_dense_inverse_export(np.ascontiguousarray(a), n, inv)
        assert np.array_equal(inv, np.array(java_inverse(a.tolist())))
        assert np.allclose(inv @ a, np.eye(n), atol=1e-8)
    # "no pivot": the half-built inverse is returned as it stands (DenseMatrix.java:393-394)
    a = np.array([[2.0, 4.0, 1.0], [1.0, 2.0, 0.5], [0.0, 0.0, 0.0]])
    inv = np.zeros((3, 3))
    O.lib().lro_dense_inverse_export(a, 3, inv)
    assert np.array_equal(inv, np.array(java_inverse(a.tolist())))


def test_wrmf_epochs_bit_identical_to_the_python_replay(O):
    tr = rng_csr(O, 23, 17, 0.3, 11)
    k = 5
    val = np.array([O.lib().lro_wrmf_weight(float(v), 4.0) for v in tr.val])
    assert val[0] == math.log(1.0 + 10000.0 * tr.val[0]) or abs(val[0] - math.log(1.0 + 10000.0 * tr.val[0])) < 4e-15
    rng = np.random.default_rng(2)
    P, Q = rng.normal(0, 0.1, (tr.U, k)), rng.normal(0, 0.1, (tr.I, k))
    rows, cols = _lists(tr, val)
    X, Y = P.tolist(), Q.tolist()
    for _ in range(3):
        O.lib().lro_wrmf_epoch(tr.U, tr.I, tr.rowptr, tr.col, val, k, P, Q, 0.01, 0.02)
        wrmf_replay(rows, cols, k, X, Y, float(np.float32(0.01)), float(np.float32(0.02)))
    assert np.array_equal(P, np.array(X)) and np.array_equal(Q, np.array(Y))
    # sanity of the model itself: observed entries score above the unobserved ones on average
    S = P @ Q.T
    mask = np.zeros((tr.U, tr.I), bool)
    for u in range(tr.U):
        mask[u, tr.col[tr.rowptr[u]:tr.rowptr[u + 1]]] = True
    assert S[mask].mean() > S[~mask].mean() + 0.2


def test_eals_epochs_bit_identical_to_the_python_replay(O):
    tr = rng_csr(O, 19, 13, 0.35, 4)
    k = 4
    for judge in (0, 1, 2):
        conf = np.zeros(tr.I)
        O.lib().lro_eals_confidences(tr.U, tr.I, tr.rowptr, tr.col, 0.4, 128.0, judge, conf)
        nnz = int(tr.rowptr[-1])
        if judge == 1:
            assert (conf == 1.0).all()
        else:
            cnt = np.bincount(tr.col, minlength=tr.I)
            alpha = [math.pow(int(c) * 1.0 / nnz, float(np.float32(0.4))) for c in cnt]
            s = 0.0
            for a in alpha:
                s += a
            assert np.allclose(conf, [128.0 * a / s for a in alpha], rtol=1e-15)
        val = np.array([O.lib().lro_eals_weight(float(v), 4.0, judge) for v in tr.val])
        assert val[0] == (1.0 + 4.0 * tr.val[0] if judge else 1.0)
        rng = np.random.default_rng(8)
        P, Q = np.zeros((tr.U, k)), rng.normal(0, 0.1, (tr.I, k))        # EALSRecommender.java:125: userFactors starts at zero
        rows, cols = _lists(tr, val)
        X, Y = P.tolist(), Q.tolist()
        for _ in range(3):
            O.lib().lro_eals_epoch(tr.U, tr.I, tr.rowptr, tr.col, val, k, P, Q, conf, 0.01, 0.02)
            eals_replay(rows, cols, k, X, Y, conf.tolist(), float(np.float32(0.01)), float(np.float32(0.02)))
        assert np.array_equal(P, np.array(X)) and np.array_equal(Q, np.array(Y))
        assert np.isfinite(P).all() and np.abs(P).max() > 0


def test_wrmf_hand_worked_k1(O):
    """k = 1: the system is scalar.  x_u = sum_i (w_ui + 1) y_i / (sum_all y^2 + reg + sum_i w_ui y_i^2), written out by hand
    (WRMFRecommender.java:103-123 with a 1 x 1 inverse, DenseMatrix.java:373-376)"""
    rowptr = np.array([0, 2, 3], np.int64)
    col = np.array([0, 1, 1], np.int32)
    w = np.array([2.0, 3.0, 0.5])
    P = np.array([[0.7], [-0.2]])
    Q = np.array([[0.5], [-1.5]])
    O.lib().lro_wrmf_epoch(2, 2, rowptr, col, w, 1, P, Q, 0.25, 0.125)
    yty = 0.0 + 0.5 * 0.5 + (-1.5) * (-1.5)
    x0 = ((0.0 + 0.5 * 3.0) + (-1.5) * 4.0) * (1.0 / (((yty + 0.25) + (0.5 * 2.0) * 0.5) + (-1.5 * 3.0) * -1.5))
    x1 = (0.0 + (-1.5) * 1.5) * (1.0 / ((yty + 0.25) + (-1.5 * 0.5) * -1.5))
    assert P[0, 0] == x0 and P[1, 0] == x1
    xtx = 0.0 + x0 * x0 + x1 * x1
    y0 = (0.0 + x0 * 3.0) * (1.0 / ((xtx + 0.125) + (x0 * 2.0) * x0))
    y1 = ((0.0 + x0 * 4.0) + x1 * 1.5) * (1.0 / (((xtx + 0.125) + (x0 * 3.0) * x0) + (x1 * 0.5) * x1))
    assert Q[0, 0] == y0 and Q[1, 0] == y1


def test_eals_hand_worked_k1(O):
    """k = 1, one user with two items, zero user factor (EALSRecommender.java:125): every sum written out by hand (:128-209)"""
    rowptr = np.array([0, 2], np.int64)
    col = np.array([0, 1], np.int32)
    w = np.array([3.0, 5.0])
    conf = np.array([0.5, 0.25])
    P = np.zeros((1, 1))
    q0, q1 = 0.75, -0.5
    Q = np.array([[q0], [q1]])
    O.lib().lro_eals_epoch(1, 2, rowptr, col, w, 1, P, Q, conf, 0.25, 0.125)
    sq = (0.0 + 0.5 * q0 * q0) + 0.25 * q1 * q1
    pred0 = 0.0 - 0.0 * q0
    pred1 = 0.0 - 0.0 * q1
    numer = (0.0 + (3.0 - (3.0 - 0.5) * pred0) * q0) + (5.0 - (5.0 - 0.25) * pred1) * q1
    denom = ((0.25 + sq) + (3.0 - 0.5) * q0 * q0) + (5.0 - 0.25) * q1 * q1
    p = numer / denom
    assert P[0, 0] == p
    sp = 0.0 + p * p
    out = []
    for q, c, wt in ((q0, 0.5, 3.0), (q1, 0.25, 5.0)):
        up = (0.0 + q * p) - p * q
        n = 0.0 * c + (wt - (wt - c) * up) * p
        d = (c * sp + 0.125) + (wt - c) * p * p
        out.append(n / d)
    assert Q[0, 0] == out[0] and Q[1, 0] == out[1]

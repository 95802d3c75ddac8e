"""CPU: the oracle's AoBPR restatement (SURVEY 8f N3 groundwork) against an independent pure-Python replay of
recommender/cf/ranking/AoBPRRecommender.java with its own java.util.Random -- same draws, same rank tables, same factors."""
import math

import numpy as np

from conftest import rng_csr


class JavaRandom:
    """java.util.Random (JDK 8): 48-bit LCG, nextInt(bound), nextDouble"""

    def __init__(self, seed):
        self.s = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def _next(self, bits):
        self.s = (self.s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = self.s >> (48 - bits)
        return v - (1 << bits) if v >= (1 << (bits - 1)) and bits == 32 else v

    def next_int(self, bound):
        r = self._next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if u - r + m < (1 << 31):
                return r
            u = self._next(31)

    def next_double(self):
        return ((self._next(26) << 27) + self._next(27)) * (1.0 / (1 << 53))


def _discrete(rng, a):
    total = 0.0
    for x in a:
        total = total + x
    assert abs(total - 1.0) <= 1e-6
    while True:
        r = rng.next_double()
        acc = 0.0
        for i, x in enumerate(a):
            acc = acc + x
            if acc > r:
                return i


def aobpr_replay(seed, U, I, rowptr, col, k, P, Q, lr, reg_u, reg_i, dist_param, iters):
    rng = JavaRandom(seed)
    lr, reg_u, reg_i = float(np.float32(lr)), float(np.float32(reg_u)), float(np.float32(reg_i))
    nnz = int(rowptr[-1])
    lam = int(np.float32(dist_param) * np.float32(I))
    loop = int(I * math.log(I))
    pro = [math.exp(-((i + 1) // lam)) for i in range(I)]
    tot = 0.0
    for x in pro:
        tot += x
    pro = [x / tot for x in pro]
    user_of = np.repeat(np.arange(U), np.diff(rowptr))
    rows = [set(col[rowptr[u]:rowptr[u + 1]].tolist()) for u in range(U)]
    ranking, var = None, None
    count = 0
    losses, trips = [], []
    for it in range(iters):
        loss = 0.0
        for _ in range(nnz):
            if count % loop == 0:
                ranking, var = [], []
                for f in range(k):
                    order = sorted(range(I), key=lambda i: -Q[i, f])          # stable: ties keep ascending item id
                    vals = [Q[i, f] for i in order]
                    m = 0.0
                    for v in vals:
                        m += v
                    m = m / I
                    s2 = 0.0
                    for v in vals:
                        s2 += (v - m) * (v - m)
                    ranking.append(order); var.append(s2 / I)
                count = 0
            count += 1
            while True:
                d = rng.next_int(nnz)
                u = int(user_of[d])
                if len(rows[u]) == 0 or len(rows[u]) == I:
                    continue
                i = int(col[d])
                while True:
                    r = _discrete(rng, pro)
                    pfc = [abs(P[u, f]) * var[f] for f in range(k)]
                    sfc = 0.0
                    for x in pfc:
                        sfc += x
                    pfc = [x / sfc for x in pfc]
                    f = _discrete(rng, pfc)
                    j = ranking[f][r] if P[u, f] > 0 else ranking[f][I - r - 1]
                    if j not in rows[u]:
                        break
                break
            if it == 0:
                trips.append((u, i, j))
            xui = 0.0
            for f in range(k):
                xui += P[u, f] * Q[i, f]
            xuj = 0.0
            for f in range(k):
                xuj += P[u, f] * Q[j, f]
            diff = xui - xuj
            loss += -math.log(1.0 / (1.0 + math.exp(-diff)))
            deri = 1.0 / (1.0 + math.exp(diff))
            for f in range(k):
                uf, pf, nf = P[u, f], Q[i, f], Q[j, f]
                P[u, f] += lr * (deri * (pf - nf) - reg_u * uf)
                Q[i, f] += lr * (deri * uf - reg_i * pf)
                Q[j, f] += lr * (deri * (-uf) - reg_i * nf)
                loss += reg_u * uf * uf + reg_i * pf * pf + reg_i * nf * nf
        losses.append(loss)
    return losses, trips


def test_aobpr_matches_python_replay(O):
    tr = rng_csr(O, 25, 18, 0.25, 4, values=(1.0,))
    k = 3
    rng = np.random.default_rng(2)
    P0 = rng.normal(0, 0.1, (tr.U, k)); Q0 = rng.normal(0, 0.1, (tr.I, k))
    P, Q = P0.copy(), Q0.copy()
    losses = np.zeros(3)
    trip = np.zeros(3 * tr.nnz, np.int32)
    O.lib().lro_seed(7)
    rc = O.lib().lro_aobpr_train(tr.U, tr.I, tr.rowptr, tr.col, k, P, Q, 0.05, 0.01, 0.02, 0.2, 3, losses.ctypes.data, trip.ctypes.data)
    assert rc == 0
    eP, eQ = P0.copy(), Q0.copy()
    elosses, etrips = aobpr_replay(7, tr.U, tr.I, tr.rowptr, tr.col, k, eP, eQ, 0.05, 0.01, 0.02, 0.2, 3)
    assert trip.reshape(-1, 3).tolist() == [list(t) for t in etrips]
    assert np.array_equal(P, eP) and np.array_equal(Q, eQ)
    assert all(abs(a - b) <= 1e-12 * abs(b) for a, b in zip(losses.tolist(), elosses))
    # the negatives are unrated items, and low ranks dominate: exp(-((i+1)/lambda)) with integer division is a step function
    t = trip.reshape(-1, 3)
    for u, i, j in t:
        row = tr.col[tr.rowptr[u]:tr.rowptr[u + 1]]
        assert i in row and j not in row


def test_aobpr_learns_a_ranking(O, c1):
    tr, te = c1["train"], c1["test"]
    ones = O.Csr(tr.U, tr.I, tr.rowptr, tr.col, np.ones(tr.nnz))
    k = 10
    rng = np.random.default_rng(5)
    P = rng.normal(0, 0.1, (tr.U, k)); Q = rng.normal(0, 0.1, (tr.I, k))
    losses = np.zeros(4)
    O.lib().lro_seed(1)
    assert O.lib().lro_aobpr_train(tr.U, tr.I, ones.rowptr, ones.col, k, P, Q, 0.05, 0.01, 0.01, 0.05, 4, losses.ctypes.data, None) == 0
    assert losses[-1] < losses[0]
    users = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)
    items, _, counts = O.recommend_rank(O.BPR, tr.U, tr.I, k, P, Q, None, None, 0.0, ones, 10, users=users)
    hits = sum(np.intersect1d(items[r, :counts[r]], te.col[te.rowptr[u]:te.rowptr[u + 1]]).shape[0] for r, u in enumerate(users))
    assert hits / (10.0 * users.shape[0]) > 0.05          # chance ~ 0.012

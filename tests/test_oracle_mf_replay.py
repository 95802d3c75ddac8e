"""CPU: the oracle's BiasedMF / PMF / BPR TRAINING restatements and the seeded ratio split against an independent pure-Python replay
of the Java loops (multi-epoch independent pins existed for AoBPR / GBPR / the ranking evaluators only).

The replay is written from the Java sources, not from oracle/lrk_oracle.cpp: plain Python floats (IEEE double, the JVM's arithmetic),
the float -> double promotion of learnRate / regUser / regItem (MatrixFactorizationRecommender.java:16,54,59) through numpy.float32,
its own java.util.Random (JDK 8: 48-bit LCG, nextInt(bound) with the rejection rule, nextDouble).  Results must be EQUAL, not close:
  BiasedMFRecommender.java:67-107, PMFSimilarityRecommender.java:59-90 (vanilla PMF), BPRRecommender.java:45-99,
  RatioDataSplitter.java:136-156 (one nextDouble per entry in CSR order, < ratio -> train)."""
import math

import numpy as np

from conftest import rng_csr
from test_oracle_aobpr import JavaRandom


def _f(x):
    return float(np.float32(x))


def biasedmf_replay(tr, k, P, Q, bu, bi, mu, lr, reg_u, reg_i, reg_b, iters):
    lr, reg_u, reg_i = _f(lr), _f(reg_u), _f(reg_i)           # float fields promoted to double; regBias is a double (:36)
    P, Q, bu, bi = [r[:] for r in P.tolist()], [r[:] for r in Q.tolist()], bu.tolist(), bi.tolist()
    losses = []
    for _ in range(iters):
        loss = 0.0
        for u in range(tr.U):
            for e in range(int(tr.rowptr[u]), int(tr.rowptr[u + 1])):
                i, r = int(tr.col[e]), float(tr.val[e])
                dot = 0.0
                for f in range(k):                                # DenseVector.dot: left to right from 0.0
                    dot += Q[i][f] * P[u][f]
                err = r - (dot + bu[u] + bi[i] + mu)              # :118-120 evaluated once, before the bias updates (:77-78)
                loss += err * err
                ub = bu[u]
                bu[u] += lr * (err - reg_b * ub)
                loss += reg_b * ub * ub
                ib = bi[i]
                bi[i] += lr * (err - reg_b * ib)
                loss += reg_b * ib * ib
                for f in range(k):
                    pf, qf = P[u][f], Q[i][f]
                    P[u][f] += lr * (err * qf - reg_u * pf)
                    Q[i][f] += lr * (err * pf - reg_i * qf)
                    loss += reg_u * pf * pf + reg_i * qf * qf
        losses.append(loss * 0.5)
    return np.array(P), np.array(Q), np.array(bu), np.array(bi), losses


def pmf_replay(tr, k, P, Q, lr, reg_u, reg_i, iters):
    lr, reg_u, reg_i = _f(lr), _f(reg_u), _f(reg_i)
    P, Q = [r[:] for r in P.tolist()], [r[:] for r in Q.tolist()]
    losses = []
    for _ in range(iters):
        loss = 0.0
        for u in range(tr.U):
            for e in range(int(tr.rowptr[u]), int(tr.rowptr[u + 1])):
                i, r = int(tr.col[e]), float(tr.val[e])
                dot = 0.0
                for f in range(k):
                    dot += Q[i][f] * P[u][f]
                err = r - dot
                loss += err * err
                for f in range(k):
                    pf, qf = P[u][f], Q[i][f]
                    P[u][f] += lr * (err * qf - reg_u * pf)
                    Q[i][f] += lr * (err * pf - reg_i * qf)
                    loss += reg_u * pf * pf + reg_i * qf * qf
        losses.append(loss * 0.5)
    return np.array(P), np.array(Q), losses


def bpr_replay(seed, tr, k, P, Q, lr, reg_u, reg_i, iters):
    rng = JavaRandom(seed)
    lr, reg_u, reg_i = _f(lr), _f(reg_u), _f(reg_i)
    P, Q = [r[:] for r in P.tolist()], [r[:] for r in Q.tolist()]
    rows = [tr.col[tr.rowptr[u]:tr.rowptr[u + 1]].tolist() for u in range(tr.U)]
    sets = [set(r) for r in rows]
    nnz = int(tr.rowptr[-1])
    losses = []
    for _ in range(iters):
        loss = 0.0
        for _s in range(nnz):                                     # :48 one sample per train entry
            while True:                                           # :54-67
                u = rng.next_int(tr.U)
                if len(rows[u]) == 0 or len(rows[u]) == tr.I:
                    continue
                pi = rows[u][rng.next_int(len(rows[u]))]
                while True:
                    nj = rng.next_int(tr.I)
                    if nj not in sets[u]:
                        break
                break
            dp = 0.0
            for f in range(k):
                dp += Q[pi][f] * P[u][f]
            dn = 0.0
            for f in range(k):
                dn += Q[nj][f] * P[u][f]
            diff = dp - dn
            loss += -math.log(1.0 / (1.0 + math.exp(-diff)))     # Maths.logistic :127-129
            deri = 1.0 / (1.0 + math.exp(diff))
            for f in range(k):
                uf, pf, nf = P[u][f], Q[pi][f], Q[nj][f]
                P[u][f] += lr * (deri * (pf - nf) - reg_u * uf)
                Q[pi][f] += lr * (deri * uf - reg_i * pf)
                Q[nj][f] += lr * (deri * (-uf) - reg_i * nf)
                loss += reg_u * uf * uf + reg_i * pf * pf + reg_i * nf * nf
        losses.append(loss)                                       # no 0.5
    return np.array(P), np.array(Q), losses


def _case(O, seed, U=40, I=30, k=5):
    tr = rng_csr(O, U, I, 0.25, seed, values=(0.5, 1.0, 2.5, 3.0, 4.0, 5.0))
    rng = np.random.default_rng(seed)
    return tr, rng.normal(0, 0.1, (U, k)), rng.normal(0, 0.1, (I, k)), rng.normal(0, 0.1, U), rng.normal(0, 0.1, I)


def test_biasedmf_three_epochs_equal_the_python_replay(O):
    tr, P, Q, bu, bi = _case(O, 1)
    oP, oQ, obu, obi = P.copy(), Q.copy(), bu.copy(), bi.copy()
    ol = [O.lib().lro_biasedmf_epoch(tr.U, tr.rowptr, tr.col, tr.val, 5, oP, oQ, obu, obi, 3.25, 0.002, 0.01, 0.03, 0.02, None, None) for _ in range(3)]
    rP, rQ, rbu, rbi, rl = biasedmf_replay(tr, 5, P, Q, bu, bi, 3.25, 0.002, 0.01, 0.03, 0.02, 3)
    assert np.array_equal(oP, rP) and np.array_equal(oQ, rQ) and np.array_equal(obu, rbu) and np.array_equal(obi, rbi)
    assert ol == rl


def test_pmf_three_epochs_equal_the_python_replay(O):
    tr, P, Q, _, _ = _case(O, 2, k=6)
    oP, oQ = P.copy(), Q.copy()
    ol = [O.lib().lro_pmf_epoch(tr.U, tr.rowptr, tr.col, tr.val, 6, oP, oQ, 0.01, 0.08, 0.08, None, None) for _ in range(3)]
    rP, rQ, rl = pmf_replay(tr, 6, P, Q, 0.01, 0.08, 0.08, 3)
    assert np.array_equal(oP, rP) and np.array_equal(oQ, rQ) and ol == rl


def test_bpr_two_epochs_equal_the_python_replay_with_its_own_java_random(O):
    tr, P, Q, _, _ = _case(O, 3, U=25, I=20, k=4)
    ones = O.Csr(tr.U, tr.I, tr.rowptr, tr.col, np.ones_like(tr.val))
    oP, oQ = P.copy(), Q.copy()
    O.lib().lro_seed(77)
    ol = [O.lib().lro_bpr_epoch(ones.U, ones.I, ones.rowptr, ones.col, 4, oP, oQ, 0.05, 0.01, 0.02, ones.nnz, None, None) for _ in range(2)]
    rP, rQ, rl = bpr_replay(77, ones, 4, P, Q, 0.05, 0.01, 0.02, 2)
    assert np.array_equal(oP, rP) and np.array_equal(oQ, rQ)
    # the loss goes through libm's log / exp on both sides (Java: StrictMath-compatible Math.log / Math.exp intrinsics)
    assert np.allclose(ol, rl, rtol=1e-14, atol=0)


def test_ratio_split_equals_the_python_replay(O, c1):
    """RatioDataSplitter.getRatioByRating (:136-156): one Randoms.uniform() per entry of the preference matrix in CSR order"""
    full = c1["full"]
    O.lib().lro_seed(1)
    flags = np.zeros(full.nnz, np.uint8)
    O.lib().lro_split_ratio(full.nnz, full.val, 0.8, flags)
    rng = JavaRandom(1)
    want = np.array([1 if rng.next_double() < 0.8 else 0 for _ in range(full.nnz)], np.uint8)
    assert np.array_equal(flags, want)
    # ... and this is the split the committed C1 fixture holds (tests/golden/ml100k_seed1_split.npz)
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "ml100k_seed1_split.npz"))
    assert np.array_equal(z["flags"].astype(np.uint8), want)

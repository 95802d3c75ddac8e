"""CPU: the oracle's ranking evaluators against an independent pure-Python replay of the Java code,
including a literal java.util.HashMap (JDK 8: table doubling at 0.75 load, order-preserving bucket split) for the
iteration order AUCEvaluator.java:91-97 depends on."""
import math

import numpy as np


class JavaHashSet:
    """java.util.HashSet<Integer> of JDK 8 without tree bins: insertion + iteration order"""

    def __init__(self):
        self.cap, self.table, self.size = 0, [], 0

    @staticmethod
    def _hash(key):
        h = key & 0xFFFFFFFF
        return h ^ (h >> 16)

    def _resize(self):
        old = self.table
        self.cap = 16 if self.cap == 0 else self.cap * 2
        self.table = [[] for _ in range(self.cap)]
        for bucket in old:                       # split keeps the relative order of a chain
            for k in bucket:
                self.table[self._hash(k) & (self.cap - 1)].append(k)

    def add(self, key):
        if self.cap == 0:
            self._resize()
        b = self.table[self._hash(key) & (self.cap - 1)]
        if key in b:
            return
        b.append(key)
        self.size += 1
        if self.size > 0.75 * self.cap:
            self._resize()

    def __iter__(self):
        for bucket in self.table:
            yield from bucket

    def __contains__(self, key):
        return self.cap > 0 and key in self.table[self._hash(key) & (self.cap - 1)]


def java_eval(topn, rec_items, rec_counts, t_rowptr, t_col, t_val, num_dropped):
    U = len(rec_counts)
    acc = dict(AUC=0.0, AP=0.0, NDCG=0.0, Precision=0.0, Recall=0.0, RR=0.0)
    nz = nz_ap = 0
    for u in range(U):
        test = [(int(t_col[e]), float(t_val[e])) for e in range(t_rowptr[u], t_rowptr[u + 1])]
        if not test:
            continue
        nz += 1
        tset = JavaHashSet()
        for k, _ in test:
            tset.add(k)
        rec = [int(x) for x in rec_items[u][:min(topn, rec_counts[u])]]
        topk = len(rec)
        hits, tmp, rr_done, dcg, has = 0, 0.0, False, 0.0, False
        gt = [[k, v] for k, v in test]
        for i, key in enumerate(rec):
            if key in tset:
                hits += 1
                tmp += 1.0 * hits / (i + 1)
                if not rr_done:
                    acc["RR"] += 1.0 / (i + 1.0); rr_done = True
                val = 0.0
                for ent in gt:                     # getValueByKey
                    if ent[0] == key:
                        ent[0] = -1; val = ent[1]; break
                has = True
                dcg += val / (math.log(i + 2) / math.log(2))
        acc["Precision"] += hits / (topn + 0.0)
        acc["Recall"] += hits / (len(test) + 0.0)
        if topk != 0:
            acc["AP"] += tmp / (len(test) if len(test) < topk else topk); nz_ap += 1
        if has and dcg != 0:
            vals = sorted([ent[1] for ent in gt if ent[0] == -1], reverse=True)
            idcg = sum(v / (math.log(i + 2) / math.log(2)) for i, v in enumerate(vals))
            if idcg != 0:
                acc["NDCG"] += dcg / idcg
        nd = int(num_dropped[u]) - topk
        rset = JavaHashSet()
        for k in rec:
            rset.add(k)
        rel = sum(1 for k in rset if k in tset)
        miss = sum(1 for k in rset if k not in tset)
        pairs = (nd + topk - rel) * rel
        if pairs == 0:
            acc["AUC"] += 0.5
            continue
        correct, h = 0, 0
        for k in tset:
            if k not in rset:
                correct += h
            else:
                h += 1
        correct += h * (nd - miss)
        acc["AUC"] += (correct + 0.0) / pairs
    out = {k: (v / nz if nz else 0.0) for k, v in acc.items()}
    out["AP"] = acc["AP"] / nz_ap if nz_ap else 0.0
    return out


def java_novelty_entropy(topn, rec_items, rec_counts, purchased, I):
    U = len(rec_counts)
    info, reco = 0.0, [0] * I
    for u in range(U):
        for key in rec_items[u][:min(topn, rec_counts[u])]:
            reco[int(key)] += 1
            c = int(purchased[int(key)])
            if c > 0:
                info += -math.log(c / U)
    ent = 0.0
    for c in reco:
        if c > 0:
            p = c / U
            ent += p * (-math.log(p))
    return info / (U * math.log(2)), ent / math.log(2)


def _random_case(rng, U, I, topn, max_test):
    rowptr, col, val = [0], [], []
    for u in range(U):
        n = int(rng.integers(0, max_test + 1))
        items = np.sort(rng.choice(I, size=n, replace=False))
        col += items.tolist(); val += rng.integers(1, 6, n).astype(float).tolist()
        rowptr.append(len(col))
    counts = rng.integers(0, topn + 1, U).astype(np.int32)
    rec = np.full((U, topn), -1, np.int32)
    for u in range(U):
        pool = np.asarray(col[rowptr[u]:rowptr[u + 1]], np.int64)
        cand = np.unique(np.concatenate([pool, rng.choice(I, size=3 * topn, replace=False)]))
        rec[u, :counts[u]] = rng.permutation(cand)[:counts[u]]
    return (np.asarray(rowptr, np.int64), np.asarray(col, np.int32), np.asarray(val, np.float64), rec, counts,
            rng.integers(topn + max_test, I, U).astype(np.int32))


def test_ranking_evaluators_match_java_replay(O):
    rng = np.random.default_rng(0)
    for U, I, topn, max_test in [(40, 300, 10, 30), (25, 200000, 5, 60), (10, 50, 10, 3), (30, 100000, 20, 120)]:
        rowptr, col, val, rec, counts, dropped = _random_case(rng, U, I, topn, max_test)
        out = np.zeros(8)
        purchased = rng.integers(0, 2 * U, I).astype(np.int32)
        safe = np.where(rec >= 0, rec, 0).astype(np.int32)
        O.lib().lro_eval_ranking(U, topn, safe, counts, rowptr, col, val, dropped, purchased, I, out)
        exp = java_eval(topn, rec, counts, rowptr, col, val, dropped)
        exp["Novelty"], exp["Entropy"] = java_novelty_entropy(topn, rec, counts, purchased, I)
        for name, got in zip(O.RANKING_MEASURES, out):
            assert abs(got - exp[name]) <= 1e-12 * max(1.0, abs(exp[name])), (name, got, exp[name])


def test_ranking_evaluators_hand_worked(O):
    # one user, test items {2:5.0, 7:3.0, 9:1.0}, list [7, 4, 2], topN 5, numDropped 20
    rowptr = np.array([0, 3, 3], np.int64); col = np.array([2, 7, 9], np.int32); val = np.array([5.0, 3.0, 1.0])
    rec = np.array([[7, 4, 2, -1, -1], [1, 2, 3, 4, 5]], np.int32); counts = np.array([3, 5], np.int32)
    out = np.zeros(8)
    rec = np.where(rec >= 0, rec, 0).astype(np.int32)
    O.lib().lro_eval_ranking(2, 5, rec, counts, rowptr, col, val, np.array([20, 20], np.int32), np.ones(12, np.int32), 12, out)
    m = dict(zip(O.RANKING_MEASURES, out))
    assert m["Precision"] == 2 / 5.0 and m["Recall"] == 2 / 3.0 and m["RR"] == 1.0          # user 1 has no test items
    assert abs(m["AP"] - (1.0 / 1 + 2.0 / 3) / 3) < 1e-15                                    # min(|test| = 3, topK = 3)
    dcg = 3.0 / 1.0 + 5.0 / 2.0; idcg = 5.0 / 1.0 + 3.0 / (math.log(3) / math.log(2))
    assert abs(m["NDCG"] - dcg / idcg) < 1e-15
    # AUC: iteration order of {2,7,9} in a 16-bucket table is 2,7,9; hits so far when 9 (not recommended) is met: 2;
    # numDroppedItems = 20 - 3 = 17, numMiss = 1 -> correct = 2 + 2 * 16 = 34, pairs = (17 + 3 - 2) * 2 = 36
    assert abs(m["AUC"] - 34 / 36) < 1e-15
    # every item bought once, 2 users: 8 recommended entries x -ln(1/2) / (2 ln 2) = 4 bits; list frequencies
    # {1:1, 2:2, 3:1, 4:2, 5:1, 7:1} of 2 lists -> 4 items at p = 1/2 -> 2 bits
    assert abs(m["Novelty"] - 4.0) < 1e-14 and abs(m["Entropy"] - 2.0) < 1e-14

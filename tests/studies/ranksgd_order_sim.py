#!/usr/bin/env python
"""CPU study behind DESIGN.md 4.4b2 (i): RankSGD on config C1 with the ORACLE's sequential arithmetic, varying only the order in
which the train entries are visited (negatives drawn by popularity with numpy; 30 epochs, lr 0.01, k = 10).
    python tests/studies/ranksgd_order_sim.py        -> one line per order: final loss, Precision@10 on the C1 test split
Orders: csr (the reference), shuffled (fresh permutation per epoch), fixedshuffle (one permutation), hotfirst (the device
stream walked front to back: entries of items with >= 128 ratings first, then the rest)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from oracle import oracle as O
    z = np.load(os.path.join(ROOT, "tests", "golden", "ml100k_seed1_split.npz"))
    full = O.Csr(int(z["U"]), int(z["I"]), z["rowptr"].astype(np.int64), z["col"].astype(np.int32), z["val"].astype(np.float64))
    tr, te = full.select(z["flags"] == 1), full.select(z["flags"] == 0)
    users = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)
    rows = tr.rows().astype(np.int32)
    pop = np.bincount(tr.col, minlength=tr.I).astype(np.float64)
    cdf = np.cumsum(pop) / pop.sum()
    rated = np.zeros((tr.U, tr.I), bool)
    rated[rows, tr.col] = True
    hot = pop[tr.col] >= 128

    def p10(P, Q):
        items, _, counts = O.recommend_rank(O.BPR, tr.U, tr.I, 10, P, Q, None, None, 0.0, tr, 10, users=users)
        hits = sum(np.intersect1d(items[r, :counts[r]], te.col[te.rowptr[u]:te.rowptr[u + 1]]).shape[0] for r, u in enumerate(users))
        return hits / (10.0 * users.shape[0])

    def run(mode, seed):
        O.lib().lro_rng_set_state(int(z["rng_seed"]), int(z["rng_have"]), float(z["rng_nextg"]))
        P, Q, _, _ = O.mf_setup(tr.U, tr.I, 10, False)
        rng = np.random.default_rng(seed)
        base, loss = None, 0.0
        for _ in range(30):
            if mode == "csr":
                order = np.arange(tr.nnz)
            elif mode == "shuffled":
                order = rng.permutation(tr.nnz)
            else:
                if base is None:
                    base = (np.concatenate([rng.permutation(np.flatnonzero(hot)), rng.permutation(np.flatnonzero(~hot))])
                            if mode == "hotfirst" else rng.permutation(tr.nnz))
                order = base
            u, i = rows[order], tr.col[order]
            j = np.minimum(np.searchsorted(cdf, rng.random(tr.nnz), side="left"), tr.I - 1).astype(np.int32)
            bad = rated[u, j]
            while bad.any():
                j[bad] = np.minimum(np.searchsorted(cdf, rng.random(int(bad.sum())), side="left"), tr.I - 1)
                bad = rated[u, j]
            trip = np.ascontiguousarray(np.stack([u, i, j], 1).astype(np.int32).reshape(-1))
            loss = O.lib().lro_ranksgd_epoch(tr.U, tr.I, tr.rowptr, tr.col, tr.val, 10, P, Q, 0.01, tr.nnz, trip.ctypes.data, None)
        return loss, p10(P, Q)

    for mode in ("csr", "shuffled", "fixedshuffle", "hotfirst"):
        for seed in (1, 2):
            loss, p = run(mode, seed)
            print("%-13s seed %d: loss_30 %.1f  Precision@10 %.4f" % (mode, seed, loss, p), flush=True)


if __name__ == "__main__":
    main()

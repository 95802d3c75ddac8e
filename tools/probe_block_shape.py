#!/usr/bin/env python
"""One-GPU probe for the DSGD item partition: time the SGD epoch kernel on each of the G item
blocks of the ML-20M-shaped matrix, for (a) contiguous blocks balanced by rating count (what dsgd.cuh does today)
and (b) striped blocks (items dealt to blocks in popularity order, so every block has the same popularity mix).
A DSGD sub-epoch lasts as long as its slowest block, so max/mean over blocks is the skew the partition costs.
    python tools/probe_block_shape.py [--striped] [G ...]      (G = 1: the whole matrix)
"""
import os
import sys
import json

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sub_csr(d, rows, keep_local):
    """ratings whose item maps to keep_local[item] >= 0, item ids replaced by the local ones"""
    loc = keep_local[d["col"]]
    m = loc >= 0
    cnt = np.bincount(rows[m], minlength=d["U"])
    rowptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    col = loc[m].astype(np.int32)
    val = d["val"][m]
    # rows stay grouped (mask keeps order); sort columns inside each row as CSR requires
    key = rows[m].astype(np.int64) * (int(keep_local.max()) + 1) + col
    o = np.argsort(key, kind="stable")
    return rowptr, np.ascontiguousarray(col[o]), np.ascontiguousarray(val[o])


def main():
    from librec_b200 import capi, synth
    capi.load()
    striped = "--striped" in sys.argv
    only = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--block=")]
    Gs = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [8, 4]
    d = synth.make_ratings("ml-20m", shard=0)
    U, I = d["U"], d["I"]
    nnz = int(d["rowptr"][-1])
    rows = np.repeat(np.arange(U, dtype=np.int64), np.diff(d["rowptr"]))
    cnt = np.bincount(d["col"], minlength=I).astype(np.int64)
    k = 64
    rng = np.random.default_rng(5)
    P0 = rng.normal(0, 0.001, (U, k)); bu0 = rng.normal(0, 0.001, U)
    mu = float(d["val"].mean())
    out = []
    for G in Gs:
        # (a) contiguous, balanced by count (dsgd_item_bounds)
        cum = np.cumsum(cnt)
        bounds = [0] + [int(np.searchsorted(cum, cum[-1] * b / G, side="left")) + 1 for b in range(1, G)] + [I]
        parts = {"contiguous": [np.arange(bounds[b], bounds[b + 1]) for b in range(G)]}
        # (b) striped: popularity order dealt in snake order
        order = np.argsort(-cnt, kind="stable")
        pos = np.arange(I)
        rnd, within = pos // G, pos % G
        blk = np.where(rnd % 2 == 0, within, G - 1 - within)
        if striped:
            parts["striped"] = [order[blk == b] for b in range(G)]
        for name, plist in parts.items():
            times, sizes, guard, loss = [], [], [], []
            for b, items in enumerate(plist):
                if only and b not in only:
                    continue
                keep = np.full(I, -1, np.int64)
                keep[items] = np.arange(items.shape[0])
                rowptr, col, val = sub_csr(d, rows, keep)
                Ib = int(items.shape[0])
                with capi.Handle(capi.MODEL_BIASEDMF, k, device=0, seed=1) as h:
                    h.set_train_csr(U, Ib, rowptr, col, val)
                    h.set_factors(P0, rng.normal(0, 0.001, (Ib, k)), bu0, rng.normal(0, 0.001, Ib), mu)
                    ms, ls = [], []
                    for e in range(8):
                        ls.append(h.sgd_epoch(0.002, 0.01, 0.01, 0.01, e + 1))
                        ms.append(h.last_epoch_ms())
                    guard.append(h.sgd_safeguard()["rollbacks"])
                    loss.append(ls[-1])
                times.append(float(np.median(ms[3:])))
                sizes.append(int(col.shape[0]))
            line = {"G": G, "partition": name, "ratings": sizes, "items": [int(p.shape[0]) for p in plist], "blocks": only or list(range(G)), "kernel_ms": times, "rollbacks": guard, "loss_8": loss,
                    "sum_ms": float(np.sum(times)), "G_x_max_ms": float(G * np.max(times))}
            print(json.dumps(line), flush=True)
            out.append(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())

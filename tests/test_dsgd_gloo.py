"""CPU, world_size 2 over gloo: the DSGD stratum schedule (librec_b200/dsgd_plan.py == csrc/dsgd.cuh).

Two processes own one user shard each, walk the sub-epochs of the plan, and exchange item blocks with
torch.distributed send/recv exactly like the NCCL ring in the library; the per-stratum arithmetic is
the oracle's BiasedMF epoch.  Because the strata of a sub-epoch are disjoint in users and items, the
result must equal -- bit for bit -- a single-process sequential walk over the strata.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, rng_csr


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _problem(O):
    full = rng_csr(O, 60, 45, 0.25, 21)
    rng = np.random.default_rng(3)
    k = 6
    P = rng.normal(0, 0.1, (full.U, k)); Q = rng.normal(0, 0.1, (full.I, k))
    bu = rng.normal(0, 0.1, full.U); bi = rng.normal(0, 0.1, full.I)
    return full, k, P, Q, bu, bi


def _shard_rows(U, world):
    return [(r * U) // world for r in range(world + 1)]


def _stratum_epoch(O, csr_rows, full, sel, k, P, Q, bu, bi, mu):
    """oracle BiasedMF pass over the entries `sel` (CSR order) -- one stratum"""
    if len(sel) == 0:
        return 0.0
    order = np.ascontiguousarray(sel, np.int64)
    # oracle returns 0.5 * loss of the visited entries
    return O.lib().lro_biasedmf_epoch(full.U, _rowptr_for(full, order), full.col, full.val, k,
                                      P, Q, bu, bi, mu, 0.01, 0.02, 0.02, 0.03, order.ctypes.data, csr_rows.ctypes.data)


def _rowptr_for(full, order):
    # lro_biasedmf_epoch visits rowptr[U] entries of `order`: hand it a rowptr whose last value is len(order)
    rp = full.rowptr.copy()
    rp[-1] = len(order)
    return rp


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from librec_b200 import dsgd_plan as plan
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full, k, P, Q, bu, bi = _problem(O)
    rows = full.rows()
    shard = _shard_rows(full.U, world)
    mine = np.nonzero((rows >= shard[rank]) & (rows < shard[rank + 1]))[0]
    # global item counts via all-reduce (the library does this with ncclAllReduce)
    cnt = torch.from_numpy(np.bincount(full.col[mine], minlength=full.I).astype(np.int64))
    dist.all_reduce(cnt)
    bounds = plan.item_bounds(cnt.numpy(), world)
    segs = plan.segments(full.col[mine], bounds)
    held = plan.block_at(rank, world, 0)
    qblk = Q[bounds[held]:bounds[held + 1]].copy(); bblk = bi[bounds[held]:bounds[held + 1]].copy()
    loss = 0.0
    for epoch in range(2):
        for sub in range(world):
            b = plan.block_at(rank, world, sub)
            assert b == held
            Q[bounds[b]:bounds[b + 1]] = qblk; bi[bounds[b]:bounds[b + 1]] = bblk
            loss += _stratum_epoch(O, rows, full, mine[segs[b]], k, P, Q, bu, bi, 3.0)
            qblk = Q[bounds[b]:bounds[b + 1]].copy(); bblk = bi[bounds[b]:bounds[b + 1]].copy()
            nxt = plan.block_at(rank, world, sub + 1)
            payload = torch.from_numpy(np.concatenate([qblk.ravel(), bblk]))
            incoming = torch.zeros((bounds[nxt + 1] - bounds[nxt]) * (k + 1), dtype=torch.float64)
            reqs = [dist.isend(payload, plan.send_peer(rank, world)), dist.irecv(incoming, plan.recv_peer(rank, world))]
            for r in reqs:
                r.wait()
            n = bounds[nxt + 1] - bounds[nxt]
            qblk = incoming[:n * k].numpy().reshape(n, k).copy(); bblk = incoming[n * k:].numpy().copy()
            held = nxt
    t = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(t)
    Q[bounds[held]:bounds[held + 1]] = qblk; bi[bounds[held]:bounds[held + 1]] = bblk
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), P=P[shard[rank]:shard[rank + 1]], bu=bu[shard[rank]:shard[rank + 1]],
             Q=qblk, bi=bblk, held=held, loss=t.item(), bounds=np.asarray(bounds))
    dist.destroy_process_group()


def test_plan_functions():
    from librec_b200 import dsgd_plan as plan
    for world in (1, 2, 3, 4, 8):
        for sub in range(world):
            blocks = [plan.block_at(r, world, sub) for r in range(world)]
            assert sorted(blocks) == list(range(world))                      # a sub-epoch covers every block once
            for r in range(world):                                           # what I send is what my peer trains next
                assert plan.block_at(plan.send_peer(r, world), world, sub + 1) == plan.block_at(r, world, sub)
                assert plan.block_at(r, world, sub + 1) == plan.block_at(plan.recv_peer(r, world), world, sub)
        assert all(plan.block_at(r, world, world) == r for r in range(world))  # ring closes after G hops
    cnt = np.array([100, 1, 1, 1, 50, 50, 1, 1, 95, 0])
    b = plan.item_bounds(cnt, 3)
    assert b[0] == 0 and b[-1] == 10 and all(x <= y for x, y in zip(b, b[1:]))
    sums = [int(cnt[b[i]:b[i + 1]].sum()) for i in range(3)]
    assert sum(sums) == 300 and max(sums) <= 155
    assert plan.item_bounds(np.ones(2), 4)[-1] == 2                          # more ranks than items: empty blocks allowed
    segs = plan.segments(np.array([0, 9, 4, 5, 8]), b)
    assert sorted(np.concatenate(segs).tolist()) == [0, 1, 2, 3, 4]


def test_dsgd_world2_equals_sequential_strata(O, tmp_path):
    import torch.multiprocessing as mp
    from librec_b200 import dsgd_plan as plan
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # sequential reference: strata in (epoch, sub-epoch, rank) order
    full, k, P, Q, bu, bi = _problem(O)
    rows = full.rows()
    shard = _shard_rows(full.U, world)
    bounds = plan.item_bounds(np.bincount(full.col, minlength=full.I), world)
    loss = 0.0
    for epoch in range(2):
        for sub in range(world):
            for r in range(world):
                mine = np.nonzero((rows >= shard[r]) & (rows < shard[r + 1]))[0]
                segs = plan.segments(full.col[mine], bounds)
                loss += _stratum_epoch(O, rows, full, mine[segs[plan.block_at(r, world, sub)]], k, P, Q, bu, bi, 3.0)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert z["bounds"].tolist() == bounds and int(z["held"]) == r
        assert np.array_equal(z["P"], P[shard[r]:shard[r + 1]]) and np.array_equal(z["bu"], bu[shard[r]:shard[r + 1]])
        assert np.array_equal(z["Q"], Q[bounds[r]:bounds[r + 1]]) and np.array_equal(z["bi"], bi[bounds[r]:bounds[r + 1]])
        assert abs(float(z["loss"]) - loss) < 1e-9 * abs(loss)

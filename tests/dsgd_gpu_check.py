"""Multi-GPU DSGD parity check -- run under torchrun with one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dsgd_gpu_check.py
(1) conflict-free matrix: one DSGD epoch == the oracle's epoch up to fp32 rounding;
(2) config C1 (seeded ml-100k split, biasedmf-test.properties): RMSE / MAE within 1e-3 of the oracle.
Prints "DSGD-CHECK OK" on rank 0.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def shard_rows(U, world):
    return [(r * U) // world for r in range(world + 1)]


def local_csr(O, full, lo, hi):
    a, b = full.rowptr[lo], full.rowptr[hi]
    return O.Csr(hi - lo, full.I, full.rowptr[lo:hi + 1] - a, full.col[a:b], full.val[a:b])


def run_dsgd(capi, dist, torch, O, model, full, k, P, Q, bu, bi, mu, hyper, iters, rank, world, local):
    sh = shard_rows(full.U, world)
    mine = local_csr(O, full, sh[rank], sh[rank + 1])
    h = capi.Handle(model, k, device=local, seed=1)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    h.comm_init(rank, world, uid[0])
    h.set_train_csr(mine.U, mine.I, mine.rowptr, mine.col, mine.val)
    h.set_factors(P[sh[rank]:sh[rank + 1]], Q, None if bu is None else bu[sh[rank]:sh[rank + 1]], bi, mu)
    losses = [h.sgd_epoch(*hyper, it + 1) for it in range(iters)]
    gP, gQ, gbu, gbi = h.get_factors()
    h.close()
    parts = [None] * world
    dist.all_gather_object(parts, (gP, gbu))
    allP = np.concatenate([p[0] for p in parts])
    allbu = None if bu is None else np.concatenate([p[1] for p in parts])
    return allP, gQ, allbu, gbi, losses


def main():
    import torch
    import torch.distributed as dist
    from librec_b200 import capi
    from oracle import oracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.load()

    # (1) conflict-free: every user and item exactly once
    n, I, k = 4000, 5000, 64
    rng = np.random.default_rng(3)
    items = rng.permutation(I)[:n].astype(np.int32)
    vals = rng.integers(1, 11, n).astype(np.float64) / 2.0
    cf = O.Csr(n, I, np.arange(n + 1, dtype=np.int64), items, vals)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    P, Q = f32(rng.normal(0, 0.1, (n, k))), f32(rng.normal(0, 0.1, (I, k)))
    bu, bi = f32(rng.normal(0, 0.1, n)), f32(rng.normal(0, 0.1, I))
    gP, gQ, gbu, gbi, losses = run_dsgd(capi, dist, torch, O, capi.MODEL_BIASEDMF, cf, k, P, Q, bu, bi, 3.0,
                                        (0.01, 0.02, 0.03, 0.04), 1, rank, world, local)
    oP, oQ, obu, obi = P.copy(), Q.copy(), bu.copy(), bi.copy()
    oloss = O.lib().lro_biasedmf_epoch(cf.U, cf.rowptr, cf.col, cf.val, k, oP, oQ, obu, obi, 3.0, 0.01, 0.02, 0.03, 0.04, None, None)
    ok1 = (np.allclose(gP, oP, rtol=0, atol=2e-6) and np.allclose(gQ, oQ, rtol=0, atol=2e-6) and
           np.allclose(gbu, obu, rtol=0, atol=2e-6) and np.allclose(gbi, obi, rtol=0, atol=2e-6) and
           abs(losses[0] - oloss) <= 2e-5 * abs(oloss))

    # (2) C1
    z = np.load(os.path.join(ROOT, "tests", "golden", "ml100k_seed1_split.npz"))
    full = O.Csr(int(z["U"]), int(z["I"]), z["rowptr"].astype(np.int64), z["col"].astype(np.int32), z["val"].astype(np.float64))
    tr, te = full.select(z["flags"] == 1), full.select(z["flags"] == 0)
    pins = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_c1.json")))
    O.lib().lro_rng_set_state(int(z["rng_seed"]), int(z["rng_have"]), float(z["rng_nextg"]))
    P, Q, bu, bi = O.mf_setup(tr.U, tr.I, 20, True)
    mu = pins["global_mean"]
    gP, gQ, gbu, gbi, losses = run_dsgd(capi, dist, torch, O, capi.MODEL_BIASEDMF, tr, 20, P, Q, bu, bi, mu,
                                        (0.002, 0.01, 0.01, 0.01), 100, rank, world, local)
    rmse, mae = O.eval_rating(O.BIASEDMF, te, 20, gP, gQ, gbu, gbi, mu, 1.0, 5.0)
    ok2 = abs(rmse - pins["biasedmf"]["rmse"]) < 1e-3 and abs(mae - pins["biasedmf"]["mae"]) < 1e-3
    if rank == 0:
        print("conflict-free ok=%s  loss %.6f vs oracle %.6f" % (ok1, losses[0] if False else 0.0, oloss))
        print("C1 DSGD world=%d: rmse %.6f (oracle %.6f)  mae %.6f (oracle %.6f)  loss_100 %.2f (oracle %.2f)" % (
            world, rmse, pins["biasedmf"]["rmse"], mae, pins["biasedmf"]["mae"], losses[-1], pins["biasedmf"]["loss_100"]))
        print("DSGD-CHECK OK" if (ok1 and ok2) else "DSGD-CHECK FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if (ok1 and ok2) else 1)


if __name__ == "__main__":
    main()

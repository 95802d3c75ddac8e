"""DSGD parity checks shared by tests/dsgd_gpu_check.py (torchrun script) and bench.py --gpus N (which runs them BEFORE its timed
region and exits non-zero if they fail).  Test infrastructure: the oracle is the checker.
  conflict_free : every user and item exactly once -> one DSGD epoch == the oracle's epoch up to fp32 rounding (2e-6)
  c1            : the seeded ml-100k split, biasedmf-test.properties, 100 epochs -> RMSE / MAE within 1e-3 of the oracle pins
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_rows(U, world):
    return [(r * U) // world for r in range(world + 1)]


def local_csr(O, full, lo, hi):
    a, b = full.rowptr[lo], full.rowptr[hi]
    return O.Csr(hi - lo, full.I, full.rowptr[lo:hi + 1] - a, full.col[a:b], full.val[a:b])


def run_dsgd(capi, dist, O, model, full, k, P, Q, bu, bi, mu, hyper, iters, rank, world, local, uid=None):
    """train `iters` DSGD epochs with the users split into `world` contiguous blocks; returns the gathered factors on every rank"""
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
    guard = h.sgd_safeguard()
    h.close()
    parts = [None] * world
    dist.all_gather_object(parts, (gP, gbu))
    allP = np.concatenate([p[0] for p in parts])
    allbu = None if bu is None else np.concatenate([p[1] for p in parts])
    return allP, gQ, allbu, gbi, losses, guard


def conflict_free(capi, dist, O, rank, world, local):
    n, I, k = 4000, 5000, 64
    rng = np.random.default_rng(3)
    items = rng.permutation(I)[:n].astype(np.int32)
    vals = rng.integers(1, 11, n).astype(np.float64) / 2.0
    cf = O.Csr(n, I, np.arange(n + 1, dtype=np.int64), items, vals)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    P, Q = f32(rng.normal(0, 0.1, (n, k))), f32(rng.normal(0, 0.1, (I, k)))
    bu, bi = f32(rng.normal(0, 0.1, n)), f32(rng.normal(0, 0.1, I))
    gP, gQ, gbu, gbi, losses, _ = run_dsgd(capi, dist, O, capi.MODEL_BIASEDMF, cf, k, P, Q, bu, bi, 3.0,
                                           (0.01, 0.02, 0.03, 0.04), 1, rank, world, local)
    oP, oQ, obu, obi = P.copy(), Q.copy(), bu.copy(), bi.copy()
    oloss = O.lib().lro_biasedmf_epoch(cf.U, cf.rowptr, cf.col, cf.val, k, oP, oQ, obu, obi, 3.0, 0.01, 0.02, 0.03, 0.04, None, None)
    err = max(float(np.abs(gP - oP).max()), float(np.abs(gQ - oQ).max()), float(np.abs(gbu - obu).max()), float(np.abs(gbi - obi).max()))
    ok = bool(err <= 2e-6 and abs(losses[0] - oloss) <= 2e-5 * abs(oloss))
    return {"max_abs_factor_error": err, "tol": 2e-6, "loss": losses[0], "loss_oracle": float(oloss), "ok": ok}


def load_c1(O):
    z = np.load(os.path.join(ROOT, "tests", "golden", "ml100k_seed1_split.npz"))
    full = O.Csr(int(z["U"]), int(z["I"]), z["rowptr"].astype(np.int64), z["col"].astype(np.int32), z["val"].astype(np.float64))
    tr, te = full.select(z["flags"] == 1), full.select(z["flags"] == 0)
    pins = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_c1.json")))
    return z, tr, te, pins


def c1(capi, dist, O, rank, world, local):
    z, tr, te, pins = load_c1(O)
    O.lib().lro_rng_set_state(int(z["rng_seed"]), int(z["rng_have"]), float(z["rng_nextg"]))
    P, Q, bu, bi = O.mf_setup(tr.U, tr.I, 20, True)
    mu = pins["global_mean"]
    gP, gQ, gbu, gbi, losses, guard = run_dsgd(capi, dist, O, capi.MODEL_BIASEDMF, tr, 20, P, Q, bu, bi, mu,
                                               (0.002, 0.01, 0.01, 0.01), 100, rank, world, local)
    rmse, mae = O.eval_rating(O.BIASEDMF, te, 20, gP, gQ, gbu, gbi, mu, 1.0, 5.0)
    ok = bool(abs(rmse - pins["biasedmf"]["rmse"]) < 1e-3 and abs(mae - pins["biasedmf"]["mae"]) < 1e-3)
    return {"rmse": rmse, "rmse_oracle": pins["biasedmf"]["rmse"], "mae": mae, "mae_oracle": pins["biasedmf"]["mae"], "tol": 1e-3,
            "loss_100": losses[-1], "loss_100_oracle": pins["biasedmf"]["loss_100"], "rollbacks": guard["rollbacks"], "ok": ok}

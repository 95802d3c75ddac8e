"""At-scale statistical parity of the fast SGD kernels against the oracle's sequential restatement
(test infrastructure; imported by tests/test_gpu_parity_scale.py and by bench.py's `parity` key).

The reference walks the ratings one at a time (BiasedMFRecommender.java:67-107,
PMFSimilarityRecommender.java:59-90); the fast kernels keep thousands of ratings in flight, with item-run
tiles, the staleness-aware step and the curvature factor.  ml-100k (config C1) never
exercises those at full concurrency -- the grid is capped there -- so this module trains BOTH on the
benchmark shapes from identical initial factors on a seeded 80/20 split of the synthetic matrix and compares
the held-out RMSE / MAE (`lrk_eval_rating` for the device factors, the oracle's evaluator for its own).

  "c2":  BiasedMF k=64, lr 0.002, reg 0.01 on the ML-20M shape (BASELINE configs[1])
  "c4p": PMF k=128, lr 0.01, reg 0.08 on a 10 M-rating Netflix-shaped matrix (same 17 770-item catalogue and
         popularity law as configs[3], one tenth of the users -- the oracle needs ~3 s per epoch on it)
  "c3":  BPR k=128, lr 0.01, reg 0.01 on a binarised 2 M-rating ML-shaped matrix: AUC / Precision@10
"""
import time

import numpy as np

TOL = {"c2": 2e-3, "c4p": 5e-3}      # |held-out RMSE(device) - RMSE(oracle)| and the same for MAE


def seeded_split(d, seed, ratio=0.8):
    """one uniform draw per entry in CSR order, < ratio -> train (the rule of RatioDataSplitter.java:136-156, numpy stream)"""
    rng = np.random.default_rng(seed)
    return rng.random(int(d["rowptr"][-1])) < ratio


def rating_parity(capi, O, which, epochs=10, device=0, update_mode=None, world_handle_factory=None):
    from librec_b200 import synth
    if which == "c2":
        shape, model_c, model_o, k, lr, reg, reg_b = "ml-20m", capi.MODEL_BIASEDMF, O.BIASEDMF, 64, 0.002, 0.01, 0.01
    elif which == "c4p":
        shape, model_c, model_o, k, lr, reg, reg_b = "netflix-10m", capi.MODEL_PMF, O.PMF, 128, 0.01, 0.08, 0.0
    else:
        raise ValueError(which)
    biased = model_c == capi.MODEL_BIASEDMF
    t_all = time.perf_counter()
    d = synth.make_ratings(shape)
    full = O.Csr(d["U"], d["I"], d["rowptr"], d["col"], d["val"])
    mask = seeded_split(d, 0x5EED + len(which))
    tr, te = full.select(mask), full.select(~mask)
    mu = float(tr.val.mean())
    P0, Q0, bu0, bi0 = synth.init_factors(full.U, full.I, k, 21, biased)
    # oracle: the reference's sequential loop, fp64
    oP, oQ = P0.copy(), Q0.copy()
    obu, obi = (bu0.copy(), bi0.copy()) if biased else (None, None)
    t0 = time.perf_counter()
    _, olosses = O.train(model_o, tr, k, oP, oQ, obu, obi, mu, lr, 1000.0, reg, reg, reg_b, epochs)
    t_oracle = time.perf_counter() - t0
    o_rmse, o_mae = O.eval_rating(model_o, te, k, oP, oQ, obu, obi, mu, 0.5, 5.0)
    # device: the fast kernel through the C ABI
    kw = {} if update_mode is None else {"update_mode": update_mode}
    with capi.Handle(model_c, k, device=device, seed=1, **kw) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P0, Q0, bu0, bi0, mu)
        glosses = [h.sgd_epoch(lr, reg, reg, reg_b, it + 1) for it in range(epochs)]
        g_rmse, g_mae = h.eval_rating(te.U, te.rowptr, te.col, te.val, 0.5, 5.0)
        guard = h.sgd_safeguard()
    tol = TOL[which]
    return {"config": which, "epochs": epochs, "train_ratings": tr.nnz, "test_ratings": te.nnz,
            "rmse": g_rmse, "rmse_oracle": o_rmse, "mae": g_mae, "mae_oracle": o_mae,
            "d_rmse": g_rmse - o_rmse, "d_mae": g_mae - o_mae, "tol": tol,
            "loss_last": glosses[-1], "loss_last_oracle": float(olosses[-1]),
            "rollbacks": guard["rollbacks"], "oracle_s": t_oracle, "total_s": time.perf_counter() - t_all,
            "ok": bool(abs(g_rmse - o_rmse) <= tol and abs(g_mae - o_mae) <= tol and guard["rollbacks"] == 0)}


def ranking_quality(O, tr, te, k, P, Q, topn=10, max_users=4000, seed=5):
    """AUC and Precision@N of factors (P, Q) on a test split, by the oracle's restatement of recommendRank + the reference evaluators,
    over a seeded sample of users that have test items (the evaluators are per-user means, so a sample is unbiased)"""
    users = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)
    if users.shape[0] > max_users:
        users = np.sort(np.random.default_rng(seed).choice(users, max_users, replace=False)).astype(np.int32)
    items, _, counts = O.recommend_rank(O.BPR, tr.U, tr.I, k, P, Q, None, None, 0.0, tr, topn, users=users)
    # restrict both matrices to the sampled users so that eval_ranking (which walks all rows) sees exactly them
    keep = np.zeros(tr.U, bool); keep[users] = True
    sub = lambda m: O.Csr(users.shape[0], m.I, np.concatenate([[0], np.cumsum(np.diff(m.rowptr)[users])]),
                          m.col[np.repeat(keep, np.diff(m.rowptr))], m.val[np.repeat(keep, np.diff(m.rowptr))])
    m = O.eval_ranking(sub(te), sub(tr), topn, items, counts)
    return m["AUC"], m["Precision"]


def bpr_parity(capi, O, tr, te, k, lr, reg, epochs, device=0, seed_o=7, tol_auc=0.02, tol_prec=0.03):
    """BPRRecommender.java:45-99: the device sampler is Philox, the reference's java.util.Random -- the streams differ, so the gate is
    distributional: AUC / Precision@10 of the learned model within a stated tolerance of the oracle's BPR from the same initial factors"""
    rng = np.random.default_rng(31)
    sd = 0.01
    P0, Q0 = rng.normal(0, sd, (tr.U, k)), rng.normal(0, sd, (tr.I, k))
    ones = O.Csr(tr.U, tr.I, tr.rowptr, tr.col, np.ones_like(tr.val))
    oP, oQ = P0.copy(), Q0.copy()
    O.lib().lro_seed(seed_o)
    t0 = time.perf_counter()
    _, ol = O.train(O.BPR, ones, k, oP, oQ, None, None, 0.0, lr, 1000.0, reg, reg, 0.0, epochs)
    t_oracle = time.perf_counter() - t0
    with capi.Handle(capi.MODEL_BPR, k, device=device, seed=1) as h:
        h.set_train_csr(ones.U, ones.I, ones.rowptr, ones.col, ones.val)
        h.set_factors(P0, Q0)
        gl = [h.sgd_epoch(lr, reg, reg, 0.0, it + 1) for it in range(epochs)]
        gP, gQ, _, _ = h.get_factors()
    o_auc, o_prec = ranking_quality(O, ones, te, k, oP, oQ)
    g_auc, g_prec = ranking_quality(O, ones, te, k, gP, gQ)
    return {"epochs": epochs, "auc": g_auc, "auc_oracle": o_auc, "precision10": g_prec, "precision10_oracle": o_prec,
            "loss_last": gl[-1], "loss_last_oracle": float(ol[-1]), "tol_auc": tol_auc, "tol_precision": tol_prec,
            "oracle_s": t_oracle,
            "ok": bool(abs(g_auc - o_auc) <= tol_auc and abs(g_prec - o_prec) <= tol_prec)}

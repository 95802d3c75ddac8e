"""Generates the committed fixtures under tests/golden/ from the reference's bundled data.

Run in the BUILD container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs
  ml100k_seed1_split.npz : config C1 -- data/movielens/ml-100k/ratings.txt loaded with the oracle's
        restatement of TextDataConvertor/DataFrame, split by RatioDataSplitter.getRatioByRating(0.8)
        on java.util.Random(seed=1), plus the RNG state right after the split (so the Gaussian
        factor init of MatrixFactorizationRecommender.setup can be replayed anywhere).
  matrix4by4.txt         : copy of the reference's 13-line loader fixture data/test/datamodeltest/matrix4by4.txt
  datamodeltest/         : the other DATA files of the reference's loader / splitter tests (UIRT, CSV, 4x4A, ratings-date.txt,
        the test-convert-dir tree), byte for byte
  oracle_c1.json         : the oracle's own results on C1 (regression pins; NOT reference outputs --
        the Java reference cannot run here, parity is unpinned).
"""
import ctypes as C
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"


def main():
    L = O.lib()
    L.lro_seed(1)                                     # job/RecommenderJob.java:72-77, rec.random.seed=1
    full = O.load_text(REF + "/data/movielens/ml-100k/ratings.txt")
    flags = np.zeros(full.nnz, np.uint8)
    L.lro_split_ratio(full.nnz, full.val, 0.8, flags)
    seed, have, nextg = C.c_uint64(), C.c_int32(), C.c_double()
    L.lro_rng_get_state(C.byref(seed), C.byref(have), C.byref(nextg))
    np.savez_compressed(os.path.join(HERE, "ml100k_seed1_split.npz"),
                        U=np.int32(full.U), I=np.int32(full.I),
                        rowptr=full.rowptr.astype(np.int32), col=full.col.astype(np.int16),
                        val=full.val.astype(np.int8), flags=flags,
                        rng_seed=np.uint64(seed.value), rng_have=np.int32(have.value), rng_nextg=np.float64(nextg.value))
    shutil.copyfile(REF + "/data/test/datamodeltest/matrix4by4.txt", os.path.join(HERE, "matrix4by4.txt"))
    # the other data files the reference's loader / splitter tests run on (TextDataModelTestCase, *DataSplitterTestCase)
    dm = os.path.join(HERE, "datamodeltest")
    for rel in ("matrix4by4-date.txt", "matrix4by4A.txt", "testCSV.txt", "ratings-date.txt"):
        os.makedirs(dm, exist_ok=True)
        shutil.copyfile(REF + "/data/test/datamodeltest/" + rel, os.path.join(dm, rel))
    for rel in ("sytTest4by4.txt", "subdir1/sytTest4by4A.txt", "subdir2/sytTest4by4.txt"):
        dst = os.path.join(dm, "test-convert-dir", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(REF + "/data/test/test-convert-dir/" + rel, dst)
    for root, _, files in os.walk(dm):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)

    tr, te = full.select(flags == 1), full.select(flags == 0)
    mu, mn, mx = O.matrix_setup(tr)
    out = {"train_nnz": tr.nnz, "test_nnz": te.nnz, "global_mean": mu, "min_rate": mn, "max_rate": mx}
    # C1: biasedmf-test.properties (k=20, 100 iters, lr 0.002f, reg 0.01f, regB 0.01, maxlr 0.01f)
    P, Q, bu, bi = O.mf_setup(tr.U, tr.I, 20, True)
    out["biasedmf_P00"] = P[0, 0]
    out["biasedmf_bi_last"] = bi[-1]
    done, losses = O.train(O.BIASEDMF, tr, 20, P, Q, bu, bi, mu, 0.002, 0.01, 0.01, 0.01, 0.01, 100)
    rmse, mae = O.eval_rating(O.BIASEDMF, te, 20, P, Q, bu, bi, mu, mn, mx)
    out["biasedmf"] = {"iters": done, "loss_1": losses[0], "loss_100": losses[-1], "rmse": rmse, "mae": mae}
    # vanilla PMF on the same split (pmf-test.properties hyper-parameters: k=6, 70 iters, lr 0.01f, reg 0.08f)
    L.lro_rng_set_state(seed.value, have.value, nextg.value)
    P, Q, _, _ = O.mf_setup(tr.U, tr.I, 6, False)
    done, losses = O.train(O.PMF, tr, 6, P, Q, None, None, mu, 0.01, 0.01, 0.08, 0.08, 0.0, 70)
    rmse, mae = O.eval_rating(O.PMF, te, 6, P, Q, None, None, mu, mn, mx)
    out["pmf"] = {"iters": done, "loss_1": losses[0], "loss_70": losses[-1], "rmse": rmse, "mae": mae}
    # RankSGD on the same split (ranksgd-test.properties: k=10, 30 iters, lr 0.01f), sequential reference order + RNG
    L.lro_rng_set_state(seed.value, have.value, nextg.value)
    P, Q, _, _ = O.mf_setup(tr.U, tr.I, 10, False)
    done, losses = O.train(O.RANKSGD, tr, 10, P, Q, None, None, 0.0, 0.01, 0.01, 0.01, 0.01, 0.0, 30)
    users = np.flatnonzero(np.diff(te.rowptr) > 0).astype(np.int32)
    items, _, counts = O.recommend_rank(O.BPR, tr.U, tr.I, 10, P, Q, None, None, 0.0, tr, 10, users=users)
    hits = sum(np.intersect1d(items[r, :counts[r]], te.col[te.rowptr[u]:te.rowptr[u + 1]]).shape[0] for r, u in enumerate(users))
    out["ranksgd"] = {"iters": done, "loss_1": losses[0], "loss_30": losses[-1], "precision_at_10": hits / (10.0 * users.shape[0])}
    with open(os.path.join(HERE, "oracle_c1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

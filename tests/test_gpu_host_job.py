"""GPU: the host mirror driven like the reference's own test cases
(core/src/test/java/net/librec/recommender/cf/rating/BiasedMFTestCase.java:50-56 -- load the algorithm's
properties, `new RecommenderJob(conf).runJob()`), but WITH assertions against the oracle."""
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BIASEDMF_PROPS = """
rec.recommender.class=biasedmf
rec.iterator.learnrate=0.002
rec.iterator.learnrate.maximum=0.01
rec.iterator.maximum=100
rec.user.regularization=0.01
rec.item.regularization=0.01
rec.bias.regularization=0.01
rec.factor.number=20
rec.learnrate.bolddriver=false
rec.learnrate.decay=1.0
rec.random.seed=1
rec.recommender.isranking=false
"""


def _job(props, c1, extra=None):
    from librec_b200.host.binding import RecommenderJob
    text = props + "\n" + "\n".join("%s=%s" % kv for kv in (extra or {}).items())
    job = RecommenderJob(text)
    job.set_data(c1["train"].U, c1["train"].I, c1["train"], c1["test"])
    return job


def _replay_split_draws(O, c1):
    """the job seeds Randoms with rec.random.seed; in the reference the splitter then consumes one draw per rating
    before the factors are initialised -- replay that so the init matches a full reference run"""
    O.lib().lro_seed(1)
    flags = np.zeros(c1["full"].nnz, np.uint8)
    O.lib().lro_split_ratio(c1["full"].nnz, c1["full"].val, 0.8, flags)


def test_biasedmf_job_reference_order_matches_oracle_bit_for_bit(O, capi, c1):
    """rec.cuda.order=reference: the whole job (Gaussian init on java.util.Random, 100 iterations, RMSE/MAE) == oracle"""
    from librec_b200.host import binding
    H = binding.load()
    job = _job(BIASEDMF_PROPS, c1, {"rec.cuda.order": "reference"})
    # RecommenderJob's constructor seeded the RNG; burn the splitter's draws like the reference's data model does
    H.lrh_randoms_seed(1)
    for _ in range(c1["full"].nnz):
        H.lrh_randoms_uniform()
    job.run_job()
    pins = c1["pins"]["biasedmf"]
    assert job.metric("RMSE") == pytest.approx(pins["rmse"], abs=1e-12)
    assert job.metric("MAE") == pytest.approx(pins["mae"], abs=1e-12)
    log = job.log()
    it1 = [l for l in log if " iter 1:" in l][0]
    m = re.match(r"BiasedMFCudaRecommender iter 1: loss = ([0-9.E-]+), delta_loss = (-?[0-9.E-]+)", it1)
    assert m and abs(float(m.group(1)) - pins["loss_1"]) < 1e-6
    assert np.float32(float(m.group(2))) == np.float32(0.0 - float(m.group(1)))    # (float)(lastLoss - loss), Float.toString
    assert sum(" iter " in l for l in log) == 100
    assert any(l.startswith("Evaluator value:RMSE is 0.93324") for l in log) and any(l.startswith("Evaluator value:MAE is ") for l in log)
    P, Q, bu, bi, mu = job.factors(20, True)
    assert mu == c1["pins"]["global_mean"]
    O.lib().lro_rng_set_state(*c1["rng_state"])
    oP, oQ, obu, obi = O.mf_setup(c1["train"].U, c1["train"].I, 20, True)
    O.train(O.BIASEDMF, c1["train"], 20, oP, oQ, obu, obi, mu, 0.002, 0.01, 0.01, 0.01, 0.01, 100)
    assert np.array_equal(P, oP) and np.array_equal(Q, oQ) and np.array_equal(bu, obu) and np.array_equal(bi, obi)


def test_biasedmf_job_fast_mode_within_1e3(O, capi, c1):
    from librec_b200.host import binding
    H = binding.load()
    job = _job(BIASEDMF_PROPS, c1)
    H.lrh_randoms_seed(1)
    for _ in range(c1["full"].nnz):
        H.lrh_randoms_uniform()
    job.run_job()
    assert abs(job.metric("RMSE") - c1["pins"]["biasedmf"]["rmse"]) < 1e-3
    assert abs(job.metric("MAE") - c1["pins"]["biasedmf"]["mae"]) < 1e-3


def test_bolddriver_earlystop_and_divergence(O, capi, c1):
    from librec_b200.host.binding import LibrecException
    job = _job(BIASEDMF_PROPS, c1, {"rec.learnrate.bolddriver": "true", "rec.iterator.maximum": "5",
                                    "rec.iterator.learnrate.maximum": "1000", "rec.cuda.order": "reference"})
    job.run_job()
    P, Q, bu, bi, mu = job.factors(20, True)
    # oracle with bold driver from the same init (the job seeded with 1 and did not replay the split here)
    O.lib().lro_seed(1)
    oP, oQ, obu, obi = O.mf_setup(c1["train"].U, c1["train"].I, 20, True)
    O.train(O.BIASEDMF, c1["train"], 20, oP, oQ, obu, obi, mu, 0.002, 1000.0, 0.01, 0.01, 0.01, 5, bold_driver=True)
    assert np.array_equal(P, oP) and np.array_equal(bi, obi)
    bad = _job(BIASEDMF_PROPS, c1, {"rec.iterator.learnrate": "50", "rec.iterator.learnrate.maximum": "1000", "rec.iterator.maximum": "30"})
    with pytest.raises(LibrecException) as e:
        bad.run_job()
    assert "Loss = NaN or Infinity" in str(e.value)


def test_bpr_ranking_job(O, capi, c1):
    """bpr-test.properties on the C1 split: recommendRank() lists == oracle's for the factors the job learned"""
    props = """
rec.recommender.class=bpr
rec.iterator.learnrate=0.01
rec.iterator.learnrate.maximum=0.01
rec.iterator.maximum=20
rec.user.regularization=0.01
rec.item.regularization=0.01
rec.factor.number=10
rec.recommender.isranking=true
rec.recommender.ranking.topn=10
rec.random.seed=1
"""
    tr = c1["train"]
    ones = O.Csr(tr.U, tr.I, tr.rowptr, tr.col, np.ones(tr.nnz))          # data.convert.binarize.threshold=0.0
    from librec_b200.host.binding import RecommenderJob
    job = RecommenderJob(props)
    job.set_data(tr.U, tr.I, ones, c1["test"])
    job.run_job()
    counts, keys, vals = job.recommended_list()
    P, Q, _, _, _ = job.factors(10, False)
    oi, os_, oc = O.recommend_rank(O.BPR, tr.U, tr.I, 10, P, Q, None, None, 0.0, ones, 10)
    assert np.array_equal(counts, oc)
    assert np.array_equal(keys, oi[oi >= 0]) and np.array_equal(vals.view(np.int64), os_[oi >= 0].view(np.int64))
    # the model learned something: hit-rate of held-out items in the top-10 well above chance
    te = c1["test"]
    hits = 0
    off = np.concatenate([[0], np.cumsum(counts)])
    for u in range(tr.U):
        held = set(te.col[te.rowptr[u]:te.rowptr[u + 1]].tolist())
        hits += len(held.intersection(keys[off[u]:off[u + 1]].tolist()))
    precision = hits / (10.0 * tr.U)
    assert precision > 0.05, precision          # random guessing: ~ 21/1682 = 0.012
    losses = [float(re.search(r"loss = ([0-9.E-]+)", l).group(1)) for l in job.log() if " iter " in l]
    assert len(losses) == 20 and losses[-1] < losses[0]
    # the job's ranking measures (job/RecommenderJob.java:229-262) come from the device evaluators and equal the oracle's
    exp = O.eval_ranking(te, ones, 10, oi, oc)
    for java_name, name in [("PRECISION", "Precision"), ("RECALL", "Recall"), ("AUC", "AUC"), ("AP", "AP"), ("NDCG", "NDCG"), ("RR", "RR"),
                            ("Novelty", "Novelty"), ("Entropy", "Entropy")]:
        got = job.metric("%s top 10" % java_name)
        assert abs(got - exp[name]) <= 1e-12, (java_name, got, exp[name])
        assert any(l.startswith("Evaluator value:%s top 10 is " % java_name) for l in job.log())
    assert abs(exp["Precision"] - precision * tr.U / np.count_nonzero(np.diff(te.rowptr))) < 1e-12


def test_job_from_a_ratings_file_end_to_end(O, capi, c1, tmp_path):
    """`librec rec -exec` shape of a run: properties name a ratings file; the native TextDataModel loads and splits it
    (SURVEY.md 8f N2), the CUDA recommender trains and evaluates, saveResult writes `user,item,value` lines with raw ids
    (N4).  Checked against the oracle run on the SAME file with the SAME java.util.Random stream
    (seed -> split draws -> factor init, job/RecommenderJob.java:74-77)."""
    import os
    from librec_b200.host.binding import RecommenderJob
    full = c1["full"]
    rows = full.rows()
    path = os.path.join(str(tmp_path), "ratings.txt")
    with open(path, "w") as f:                                             # raw ids: user 1000+u, item 5000+i, tab separated
        for u, i, r in zip(rows.tolist(), full.col.tolist(), full.val.tolist()):
            f.write("%d\t%d\t%s\n" % (1000 + u, 5000 + i, repr(float(r))))
    props = BIASEDMF_PROPS + "\ndfs.data.dir=%s\ndata.input.path=ratings.txt\ndfs.result.dir=%s\ndata.splitter.trainset.ratio=0.8\n" % (
        str(tmp_path), os.path.join(str(tmp_path), "result"))
    job = RecommenderJob(props)
    job.run_job()
    # oracle on the same file, same RNG stream
    O.lib().lro_seed(1)
    ofull = O.load_text(path)
    otr, ote = O.split_ratio(ofull, 0.8)
    mu, mn, mx = O.matrix_setup(otr)
    P, Q, bu, bi = O.mf_setup(otr.U, otr.I, 20, True)
    O.train(O.BIASEDMF, otr, 20, P, Q, bu, bi, mu, 0.002, 0.01, 0.01, 0.01, 0.01, 100)
    ormse, omae = O.eval_rating(O.BIASEDMF, ote, 20, P, Q, bu, bi, mu, mn, mx)
    assert abs(job.metric("RMSE") - ormse) < 1e-3 and abs(job.metric("MAE") - omae) < 1e-3
    assert any(l.startswith("Dataset: [") for l in job.log()) and any("user number: 943" in l for l in job.log())
    # result sink: one line per test entry, raw ids, Java's Double.toString for the value
    out = job.save_result()
    assert out.endswith("ratings.txt-biasedmf-output/biasedmf") and os.path.exists(out)
    lines = open(out).read().splitlines()
    assert len(lines) == ote.nnz
    u0, i0, v0 = lines[0].split(",")
    assert 1000 <= int(u0) < 1000 + full.U and 5000 <= int(i0) < 5000 + full.I and 1.0 <= float(v0) <= 5.0


def test_ranksgd_ranking_job(O, capi, c1):
    """ranksgd-test.properties on the C1 split: the job trains on the device, its lists are the reference's
    lists for the factors it learned, and the ranking measures come out of the device evaluators"""
    props = """
rec.recommender.class=ranksgd
rec.iterator.learnrate=0.01
rec.iterator.learnrate.maximum=0.01
rec.iterator.maximum=30
rec.user.regularization=0.01
rec.item.regularization=0.01
rec.factor.number=10
rec.learnrate.bolddriver=false
rec.learnrate.decay=1.0
rec.recommender.isranking=true
rec.recommender.ranking.topn=10
rec.random.seed=1
"""
    tr, te = c1["train"], c1["test"]
    from librec_b200.host.binding import RecommenderJob
    job = RecommenderJob(props)
    job.set_data(tr.U, tr.I, tr, te)
    job.run_job()
    counts, keys, vals = job.recommended_list()
    P, Q, _, _, _ = job.factors(10, False)
    oi, os_, oc = O.recommend_rank(O.BPR, tr.U, tr.I, 10, P, Q, None, None, 0.0, tr, 10)
    assert np.array_equal(counts, oc)
    assert np.array_equal(keys, oi[oi >= 0]) and np.array_equal(vals.view(np.int64), os_[oi >= 0].view(np.int64))
    losses = [float(re.search(r"loss = ([0-9.E-]+)", l).group(1)) for l in job.log() if " iter " in l]
    assert len(losses) == 30 and losses[-1] < 0.75 * losses[0]
    assert any(l.startswith("RankSGDCudaRecommender iter 1: loss = ") for l in job.log())
    exp = O.eval_ranking(te, tr, 10, oi, oc)
    assert abs(job.metric("PRECISION top 10") - exp["Precision"]) <= 1e-12 and exp["Precision"] > 0.10
    assert abs(job.metric("NDCG top 10") - exp["NDCG"]) <= 1e-12


def test_kcv_job_cross_validation(O, capi, c1, tmp_path):
    """data.model.splitter=kcv: RecommenderJob.java:125-133 trains and evaluates once per fold on one recommender instance and
    prints the averages (printCVAverageResult :311-326).  The folds themselves are checked entry for entry against the oracle on
    the CPU (tests/test_host_datamodel.py); here: three folds run, every fold's RMSE is in the plausible range for 15 iterations,
    and the reported metric is the mean of the fold lines."""
    import os
    from librec_b200.host.binding import RecommenderJob
    full = c1["full"]
    path = os.path.join(str(tmp_path), "ratings.txt")
    with open(path, "w") as f:
        for u, i, r in zip(full.rows().tolist(), full.col.tolist(), full.val.tolist()):
            f.write("%d %d %s\n" % (u, i, repr(float(r))))
    props = BIASEDMF_PROPS.replace("rec.iterator.maximum=100", "rec.iterator.maximum=15") + (
        "\ndfs.data.dir=%s\ndata.input.path=ratings.txt\ndata.model.splitter=kcv\ndata.splitter.cv.number=3\n" % str(tmp_path))
    job = RecommenderJob(props)
    job.run_job()
    log = job.log()
    rmse_lines = [float(l.split(" is ")[1]) for l in log if l.startswith("Evaluator value:RMSE is ")]
    assert len(rmse_lines) == 4                                            # 3 folds + the average
    assert all(0.85 < v < 1.2 for v in rmse_lines)
    assert abs(rmse_lines[3] - sum(rmse_lines[:3]) / 3.0) < 1e-12
    assert abs(job.metric("RMSE") - rmse_lines[3]) < 1e-12
    assert "Average Evaluation Result of Cross Validation:" in log
    assert sum(" iter 15:" in l for l in log) == 3 and sum(" iter 1:" in l for l in log) == 3

"""GPU: rec.eval.classes in the host job (job/RecommenderJob.java:219-231) -- only the designated evaluators run and are logged
as "Evaluator info:<SimpleName> is <value>".  (Runs last: the wiring was written after the round's GPU budget was spent.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BPR_PROPS = """
rec.recommender.class=bpr
rec.iterator.learnrate=0.01
rec.iterator.learnrate.maximum=0.01
rec.iterator.maximum=10
rec.user.regularization=0.01
rec.item.regularization=0.01
rec.factor.number=10
rec.recommender.isranking=true
rec.recommender.ranking.topn=10
rec.random.seed=1
"""


def test_designated_ranking_evaluators(O, capi, c1):
    from librec_b200.host.binding import RecommenderJob
    tr, te = c1["train"], c1["test"]
    ones = O.Csr(tr.U, tr.I, tr.rowptr, tr.col, np.ones(tr.nnz))
    job = RecommenderJob(BPR_PROPS + "rec.eval.classes=auc,precision arhr,idcg\n")
    job.set_data(tr.U, tr.I, ones, te)
    job.run_job()
    P, Q, _, _, _ = job.factors(10, False)
    oi, _, oc = O.recommend_rank(O.BPR, tr.U, tr.I, 10, P, Q, None, None, 0.0, ones, 10)
    exp = O.eval_ranking(te, ones, 10, oi, oc)
    extra = O.eval_ranking_extra(te, 10, oi, oc)
    assert abs(job.metric("AUCEvaluator") - exp["AUC"]) <= 1e-12
    assert abs(job.metric("PrecisionEvaluator") - exp["Precision"]) <= 1e-12
    assert abs(job.metric("AverageReciprocalHitRankEvaluator") - extra["ARHR"]) <= 1e-12
    assert abs(job.metric("IdealDCGEvaluator") - extra["IDCG"]) <= 1e-12
    assert job.metric("RecallEvaluator") == -1.0                          # not designated -> not evaluated
    info = [l for l in job.log() if l.startswith("Evaluator info:")]
    assert len(info) == 4 and not any(l.startswith("Evaluator value:") for l in job.log())


def test_hitrate_needs_leave_one_out(O, capi, c1):
    from librec_b200.host.binding import RecommenderJob, LibrecException
    tr, te = c1["train"], c1["test"]
    ones = O.Csr(tr.U, tr.I, tr.rowptr, tr.col, np.ones(tr.nnz))
    job = RecommenderJob(BPR_PROPS + "rec.eval.classes=hitrate\n")
    job.set_data(tr.U, tr.I, ones, te)
    with pytest.raises(Exception) as e:
        job.run_job()
    assert "leave-one-out" in str(e.value)

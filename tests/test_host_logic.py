"""CPU: the C++ host mirror's own logic (Configuration, Randoms, Java number formatting) against the oracle /
JDK known answers.  No GPU: nothing here touches a compute entry point."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def H():
    from librec_b200.host import binding
    return binding.load()


def test_host_randoms_match_jdk_and_oracle(H, O):
    L = O.lib()
    H.lrh_randoms_seed(42)
    assert H.lrh_randoms_gaussian(0.0, 1.0) == 1.1419053154730547
    for seed in (0, 1, 20251018):
        H.lrh_randoms_seed(seed); L.lro_seed(seed)
        for _ in range(50):
            assert H.lrh_randoms_uniform_int(943) == L.lro_uniform_int(943)
            assert H.lrh_randoms_uniform() == L.lro_uniform()
            assert H.lrh_randoms_gaussian(0.0, float(np.float32(0.001))) == L.lro_gaussian(0.0, float(np.float32(0.001)))


def test_java_number_formatting(H):
    def fd(v):
        b = C.create_string_buffer(64); H.lrh_format_double(v, b, 64); return b.value.decode()

    def ff(v):
        b = C.create_string_buffer(64); H.lrh_format_float(v, b, 64); return b.value.decode()
    # Double.toString / Float.toString known answers
    assert fd(46802.94873760634) == "46802.94873760634"
    assert fd(1.0) == "1.0" and fd(100.0) == "100.0" and fd(0.001) == "0.001" and fd(-2.5) == "-2.5"
    assert fd(1.0e7) == "1.0E7" and fd(12345678.9) == "1.23456789E7" and fd(1.0e-4) == "1.0E-4" and fd(0.00012345) == "1.2345E-4"
    assert fd(float("nan")) == "NaN" and fd(float("inf")) == "Infinity" and fd(-0.0) == "-0.0"
    assert ff(0.002) == "0.002" and ff(1234.5) == "1234.5" and ff(1.0e10) == "1.0E10"


def test_configuration_getters(H):
    props = b"# comment\nrec.iterator.learnrate=0.002\nrec.factor.number = 20\nrec.learnrate.bolddriver=TRUE\nrec.empty=\n"
    assert H.lrh_conf_probe(props, b"rec.factor.number", 0, 10) == 20
    assert H.lrh_conf_probe(props, b"rec.iterator.learnrate", 1, 0.01) == float(np.float32(0.002))      # Float.valueOf then widened
    assert H.lrh_conf_probe(props, b"rec.iterator.learnrate", 2, 0.01) == 0.002
    assert H.lrh_conf_probe(props, b"rec.learnrate.bolddriver", 3, 0) == 1.0                              # Boolean.valueOf ignores case
    assert H.lrh_conf_probe(props, b"rec.missing", 0, 7) == 7 and H.lrh_conf_probe(props, b"rec.empty", 0, 9) == 9   # blank -> default


def test_unknown_recommender_class_is_reported(H):
    from librec_b200.host.binding import RecommenderJob, LibrecException
    from oracle import oracle as O
    tr = O.Csr(2, 2, [0, 1, 2], [0, 1], [1.0, 2.0])
    job = RecommenderJob({"rec.recommender.class": "svdpp", "rec.random.seed": "1"})
    job.set_data(2, 2, tr, tr)
    with pytest.raises(LibrecException) as e:
        job.run_job()
    assert "ClassNotFoundException" in str(e.value)


def test_default_rating_measures(H, O):
    """eval/Measure.java:100-107: a rating job reports RMSE, MSE, MAE and MPE (share of entries with |error| > rec.measure.mpe).
    The host evaluators zip the test CSR with the predictions like eval/rating/*Evaluator.java; checked against numpy and, for
    RMSE / MAE, the oracle's evaluator arithmetic."""
    from conftest import rng_csr
    te = rng_csr(O, 60, 40, 0.2, 9)
    rng = np.random.default_rng(1)
    pred = te.val + rng.normal(0, 0.5, te.nnz)
    pred[::7] = te.val[::7]                                     # exact hits: never counted by MPE
    pred[1::7] = te.val[1::7] + 0.005                           # below the default threshold 0.01
    out = np.zeros(4)
    H.lrh_probe_rating_measures.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    rc = H.lrh_probe_rating_measures(te.U, te.I, te.rowptr.ctypes.data, te.col.ctypes.data, te.val.ctypes.data, pred.ctypes.data, 0.01, out.ctypes.data)
    assert rc == 0
    d = te.val - pred
    se = 0.0
    for x in d:
        se += x * x                                             # sequential sum like the evaluators
    assert out[1] == se / te.nnz and out[0] == np.sqrt(se / te.nnz)
    assert abs(out[2] - np.abs(d).sum() / te.nnz) < 1e-15
    assert out[3] == np.count_nonzero(np.abs(d) > 0.01) / te.nnz and 0.5 < out[3] < 0.75
    rc = H.lrh_probe_rating_measures(te.U, te.I, te.rowptr.ctypes.data, te.col.ctypes.data, te.val.ctypes.data, pred.ctypes.data, 0.75, out.ctypes.data)
    assert rc == 0 and out[3] == np.count_nonzero(np.abs(d) > 0.75) / te.nnz


def test_host_ranking_measures_match_the_oracle(H, O):
    """the host evaluators that take over when rec.recommender.ranking.topn > 64 (the device evaluator keeps 64 entries per
    thread): all eight measures equal the oracle's on random train / test splits and random lists, topN 5 ... 100"""
    from conftest import rng_csr
    H.lrh_probe_ranking_measures.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(4)
    for U, I, dens, topn in [(60, 300, 0.08, 10), (40, 500, 0.1, 100), (25, 70000, 0.001, 5), (30, 200, 0.3, 80)]:
        full = rng_csr(O, U, I, dens, int(rng.integers(1, 1000)))
        flags = rng.random(full.nnz) < 0.7
        tr, te = full.select(flags), full.select(~flags)
        items = np.zeros((U, topn), np.int32)
        counts = np.zeros(U, np.int32)
        for u in range(U):
            cand = np.setdiff1d(np.arange(I), tr.col[tr.rowptr[u]:tr.rowptr[u + 1]])
            n = int(min(topn, cand.shape[0], rng.integers(0, topn + 1)))
            # half of the picks come from the test row so that hits occur
            trow = te.col[te.rowptr[u]:te.rowptr[u + 1]]
            pick = list(rng.permutation(trow)[:n // 2])
            rest = np.setdiff1d(cand, pick)
            pick += list(rng.permutation(rest)[:n - len(pick)])
            pick = rng.permutation(pick)
            items[u, :len(pick)] = pick
            counts[u] = len(pick)
        exp = O.eval_ranking(te, tr, topn, items, counts)
        out = np.zeros(8)
        rc = H.lrh_probe_ranking_measures(U, I, tr.rowptr.ctypes.data, tr.col.ctypes.data, te.rowptr.ctypes.data, te.col.ctypes.data,
                                          te.val.ctypes.data, topn, items.ctypes.data, counts.ctypes.data, out.ctypes.data)
        assert rc == 0
        got = dict(zip(("AUC", "AP", "NDCG", "Precision", "Recall", "RR", "Novelty", "Entropy"), out.tolist()))
        for name in O.RANKING_MEASURES:
            assert abs(got[name] - exp[name]) <= 1e-12 * max(1.0, abs(exp[name])), (name, got[name], exp[name], U, I, topn)
        assert got["Precision"] > 0 and got["AUC"] > 0


def test_hitrate_arhr_idcg_match_the_oracle_and_a_hand_worked_case(H, O):
    """the three ranking evaluators outside the default list (rec.eval.classes = hitrate, arhr, idcg)"""
    H.lrh_probe_ranking_extra.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    # hand-worked, leave-one-out: user 0 test {7} found at rank 3; user 1 test {2} missed; user 2 no test item; user 3 test {5} at rank 1
    rowptr = np.array([0, 1, 2, 2, 3], np.int64); col = np.array([7, 2, 5], np.int32)
    items = np.array([[4, 9, 7, 1], [1, 3, 4, 5], [1, 2, 3, 4], [5, 6, 7, 8]], np.int32); counts = np.array([4, 4, 4, 4], np.int32)
    out = np.zeros(3)
    assert H.lrh_probe_ranking_extra(4, 10, rowptr.ctypes.data, col.ctypes.data, 4, items.ctypes.data, counts.ctypes.data, 1, out.ctypes.data) == 0
    assert out[0] == 2 / 3.0 and abs(out[1] - (1 / 3.0 + 0.0 + 1.0) / 3) < 1e-15 and out[2] == 1.0      # IDCG of one item = 1 / log2(2)
    te = O.Csr(4, 10, rowptr, col, np.ones(3))
    exp = O.eval_ranking_extra(te, 4, items, counts)
    assert [exp["HitRate"], exp["ARHR"], exp["IDCG"]] == out.tolist()
    # with top-2 lists the rank-3 hit is gone
    assert H.lrh_probe_ranking_extra(4, 10, rowptr.ctypes.data, col.ctypes.data, 2, np.ascontiguousarray(items[:, :2]).ctypes.data,
                                     np.array([2, 2, 2, 2], np.int32).ctypes.data, 1, out.ctypes.data) == 0
    assert out[0] == 1 / 3.0 and abs(out[1] - 1.0 / 3) < 1e-15
    # not leave-one-out: HitRate throws like the reference, the other two still work and equal the oracle
    from conftest import rng_csr
    te = rng_csr(O, 50, 80, 0.1, 3)
    rng = np.random.default_rng(0)
    items = np.stack([rng.permutation(80)[:10] for _ in range(50)]).astype(np.int32); counts = np.full(50, 10, np.int32)
    assert H.lrh_probe_ranking_extra(50, 80, te.rowptr.ctypes.data, te.col.ctypes.data, 10, items.ctypes.data, counts.ctypes.data, 1, out.ctypes.data) == -1
    assert b"leave-one-out" in H.lrh_last_error()
    assert H.lrh_probe_ranking_extra(50, 80, te.rowptr.ctypes.data, te.col.ctypes.data, 10, items.ctypes.data, counts.ctypes.data, 0, out.ctypes.data) == 0
    exp = O.eval_ranking_extra(te, 10, items, counts)
    assert np.isnan(exp["HitRate"]) and abs(out[1] - exp["ARHR"]) < 1e-15 and abs(out[2] - exp["IDCG"]) < 1e-12

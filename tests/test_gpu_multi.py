"""GPU: the single-process multi-GPU handle (lrk_create_multi, rec.cuda.devices in the Java shim) -- full matrices in and out, sharding
inside.  With one device it must behave exactly like a plain handle; with two or more (skipped on a one-GPU lease) the DSGD epoch and
the sharded top-N must match the oracle / the single-device handle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _conflict_free(O, n, I, seed):
    rng = np.random.default_rng(seed)
    items = rng.permutation(I)[:n].astype(np.int32)
    vals = rng.integers(1, 11, n).astype(np.float64) / 2.0
    return O.Csr(n, I, np.arange(n + 1, dtype=np.int64), items, vals)


def _devices(capi, n):
    if capi.device_count() < n:
        pytest.skip("needs %d GPUs" % n)
    return list(range(n))


@pytest.mark.parametrize("ndev", [1, 2, 4])
def test_multi_handle_conflict_free_epoch_and_topn(O, capi, ndev):
    devs = _devices(capi, ndev)
    n, I, k = 4000, 9000, 64
    tr = _conflict_free(O, n, I, 3)
    rng = np.random.default_rng(1)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    P, Q = f32(rng.normal(0, 0.1, (n, k))), f32(rng.normal(0, 0.1, (I, k)))
    bu, bi = f32(rng.normal(0, 0.1, n)), f32(rng.normal(0, 0.1, I))
    with capi.Handle(capi.MODEL_BIASEDMF, k, devices=devs) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q, bu, bi, 3.0)
        loss = h.sgd_epoch(0.01, 0.02, 0.03, 0.04)
        gP, gQ, gbu, gbi = h.get_factors()
        users = np.array([5, 3999, 17, 2000, 1999, 0, 3000], np.int32)
        items, scores, counts = h.topn(10, users=users)
        all_items, all_scores, all_counts = h.topn(10)
        assert h.launch_count() > 0 and h.stage_stats()["ratings"] == n
    oP, oQ, obu, obi = P.copy(), Q.copy(), bu.copy(), bi.copy()
    oloss = O.lib().lro_biasedmf_epoch(tr.U, tr.rowptr, tr.col, tr.val, k, oP, oQ, obu, obi, 3.0, 0.01, 0.02, 0.03, 0.04, None, None)
    assert np.allclose(gP, oP, rtol=0, atol=2e-6) and np.allclose(gQ, oQ, rtol=0, atol=2e-6)
    assert np.allclose(gbu, obu, rtol=0, atol=2e-6) and np.allclose(gbi, obi, rtol=0, atol=2e-6)
    assert abs(loss - oloss) <= 2e-5 * abs(oloss)
    # ranking for the trained factors: bit-identical to the oracle, whichever device served the user
    oi, os_, oc = O.recommend_rank(O.BIASEDMF, n, I, k, gP, gQ, gbu, gbi, 3.0, tr, 10, users=users)
    assert np.array_equal(items, oi) and np.array_equal(scores.view(np.int64), os_.view(np.int64)) and np.array_equal(counts, oc)
    assert np.array_equal(all_items[users], oi) and np.array_equal(all_counts[users], oc)


@pytest.mark.parametrize("ndev", [2])
def test_multi_handle_c1_rmse_within_1e3(O, capi, c1, ndev):
    devs = _devices(capi, ndev)
    tr, te, pins = c1["train"], c1["test"], c1["pins"]
    O.lib().lro_rng_set_state(*c1["rng_state"])
    P, Q, bu, bi = O.mf_setup(tr.U, tr.I, 20, True)
    mu = pins["global_mean"]
    with capi.Handle(capi.MODEL_BIASEDMF, 20, devices=devs) as h:
        h.set_train_csr(tr.U, tr.I, tr.rowptr, tr.col, tr.val)
        h.set_factors(P, Q, bu, bi, mu)
        losses = h.sgd_epochs(100, 0.002, 0.01, 0.01, 0.01)
        gP, gQ, gbu, gbi = h.get_factors()
    rmse, mae = O.eval_rating(O.BIASEDMF, te, 20, gP, gQ, gbu, gbi, mu, 1.0, 5.0)
    assert abs(rmse - pins["biasedmf"]["rmse"]) < 1e-3 and abs(mae - pins["biasedmf"]["mae"]) < 1e-3
    assert np.isfinite(losses).all()


def test_multi_handle_rejects_what_it_does_not_shard(O, capi):
    devs = _devices(capi, 1)
    with capi.Handle(capi.MODEL_PMF, 8, devices=devs) as h:
        with pytest.raises(capi.LibrecException) as e:
            h.predict_pairs(np.zeros(1, np.int32), np.zeros(1, np.int32))
        assert e.value.status == capi.ERR_INVALID and "multi-device" in str(e.value)
    with pytest.raises(capi.LibrecException):
        capi.Handle(capi.MODEL_PMF, 8, devices=[0, 0])

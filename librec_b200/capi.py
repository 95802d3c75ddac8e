"""ctypes binding of include/librec_b200.h -- the same entry points a JNI / Panama shim binds.

There is no CPU fallback: importing works anywhere (so symbol checks can run without a GPU), but
every compute call needs a B200 and raises LibrecException otherwise.
"""
import ctypes as C
import os

import numpy as np

from . import _build

MODEL_BIASEDMF, MODEL_PMF, MODEL_BPR, MODEL_RANKSGD, MODEL_GBPR, MODEL_SVDPP, MODEL_AOBPR, MODEL_WRMF, MODEL_EALS = 0, 1, 2, 3, 4, 5, 6, 7, 8
UPDATE_ATOMIC, UPDATE_HOGWILD, UPDATE_REFERENCE_ORDER = 0, 1, 2
OK, ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_NOMEM, ERR_DIVERGED = 0, -1, -2, -3, -4, -5


class LibrecException(Exception):
    """mirrors net.librec.common.LibrecException (common/LibrecException.java)"""

    def __init__(self, msg, status=None):
        super().__init__(msg)
        self.status = status


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("model", C.c_int32), ("num_factors", C.c_int32),
                ("update_mode", C.c_int32), ("seed", C.c_uint64), ("topn_path", C.c_int32),
                ("reserved", C.c_int32 * 7)]


_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

# name -> (restype, argtypes); also the list the symbol-export test walks
SIGNATURES = {
    "lrk_version": (C.c_char_p, []),
    "lrk_abi_version": (C.c_int32, []),
    "lrk_device_count": (C.c_int32, []),
    "lrk_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "lrk_create_multi": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "lrk_destroy": (C.c_int, [C.c_void_p]),
    "lrk_last_error": (C.c_char_p, [C.c_void_p]),
    "lrk_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lrk_synchronize": (C.c_int, [C.c_void_p]),
    "lrk_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "lrk_host_free": (C.c_int, [C.c_void_p]),
    "lrk_set_train_csr": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrk_set_factors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]),
    "lrk_get_factors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrk_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "lrk_set_matrix": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p]),
    "lrk_get_matrix": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p]),
    "lrk_sgd_epoch": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_double, C.c_int32, C.POINTER(C.c_double)]),
    "lrk_sgd_epochs": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double, C.c_int32,
                                 C.c_void_p]),
    "lrk_stage_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "lrk_debug_stream": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "lrk_last_epoch_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "lrk_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "lrk_bpr_peek_samples": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, _i32p]),
    "lrk_predict_pairs": (C.c_int, [C.c_void_p, _i32p, _i32p, C.c_int64, _f64p]),
    "lrk_eval_rating": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                  C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lrk_topn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _i32p, _f64p, _i32p]),
    "lrk_eval_ranking": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_double)]),
    "lrk_topn_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "lrk_topn_phase_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "lrk_sgd_safeguard_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "lrk_probe_l2": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(C.c_double)]),
    "lrk_comm_unique_id": (C.c_int, [C.c_void_p]),
    "lrk_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """dlopen the in-tree CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB_PATH):
        raise LibrecException(
            "librec_b200 CUDA library is missing (%s). Build it with `python -m librec_b200._build` "
            "(needs nvcc); there is no CPU fallback." % _build.LIB_PATH)
    L = C.CDLL(_build.LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError == missing export
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc, h=None):
    if rc != OK:
        msg = load().lrk_last_error(h)
        raise LibrecException((msg or b"?").decode(), status=rc)


class Handle:
    """one native handle == one recommender instance (not thread-safe)"""

    def __init__(self, model, num_factors, device=0, update_mode=UPDATE_ATOMIC, seed=1, topn_path=0, devices=None):
        """devices=[d0, d1, ...]: a single-process multi-GPU handle (lrk_create_multi, rec.cuda.devices in the Java shim)"""
        L = load()
        cfg = Config(device=device, model=model, num_factors=num_factors, update_mode=update_mode,
                     seed=seed, topn_path=topn_path)
        h = C.c_void_p()
        if devices is not None:
            dv = np.ascontiguousarray(devices, np.int32)
            _check(L.lrk_create_multi(C.byref(cfg), _ptr(dv), dv.shape[0], C.byref(h)))
        else:
            _check(L.lrk_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.k = num_factors
        self.model = model
        self.U = self.I = 0

    def close(self):
        if getattr(self, "_h", None):
            load().lrk_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- staging
    def set_stream(self, cuda_stream):
        _check(load().lrk_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None), self._h)

    def synchronize(self):
        _check(load().lrk_synchronize(self._h), self._h)

    def set_train_csr(self, U, I, rowptr, col, val):
        rowptr = np.ascontiguousarray(rowptr, np.int64)
        col = np.ascontiguousarray(col, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        _check(load().lrk_set_train_csr(self._h, U, I, _ptr(rowptr), _ptr(col), _ptr(val)), self._h)
        self.U, self.I = U, I

    def set_factors(self, P, Q, bu=None, bi=None, mu=0.0):
        P = np.ascontiguousarray(P, np.float64)
        Q = np.ascontiguousarray(Q, np.float64)
        bu = None if bu is None else np.ascontiguousarray(bu, np.float64)
        bi = None if bi is None else np.ascontiguousarray(bi, np.float64)
        _check(load().lrk_set_factors(self._h, _ptr(P), _ptr(Q), _ptr(bu), _ptr(bi), float(mu)), self._h)

    def get_factors(self):
        P = np.empty((self.U, self.k), np.float64)
        Q = np.empty((self.I, self.k), np.float64)
        biased = self.model in (MODEL_BIASEDMF, MODEL_GBPR, MODEL_SVDPP)
        bu = np.empty(self.U, np.float64) if biased else None
        bi = np.empty(self.I, np.float64) if biased else None
        _check(load().lrk_get_factors(self._h, _ptr(P), _ptr(Q), _ptr(bu), _ptr(bi)), self._h)
        return P, Q, bu, bi

    def set_param(self, name, value):
        _check(load().lrk_set_param(self._h, name.encode(), float(value)), self._h)

    def set_matrix(self, name, values):
        values = np.ascontiguousarray(values, np.float64)
        _check(load().lrk_set_matrix(self._h, name.encode(), _ptr(values)), self._h)

    def get_matrix(self, name, shape):
        out = np.empty(shape, np.float64)
        _check(load().lrk_get_matrix(self._h, name.encode(), _ptr(out)), self._h)
        return out

    # -- training
    def sgd_epoch(self, lr, reg_u, reg_i, reg_b=0.0, epoch_idx=1):
        loss = C.c_double()
        rc = load().lrk_sgd_epoch(self._h, lr, reg_u, reg_i, float(reg_b), epoch_idx, C.byref(loss))
        _check(rc, self._h)
        return loss.value

    def sgd_epochs(self, n, lr, reg_u, reg_i, reg_b=0.0, first_epoch_idx=1, decay=1.0, max_lr=0.0):
        losses = np.zeros(n, np.float64)
        _check(load().lrk_sgd_epochs(self._h, n, lr, decay, max_lr, reg_u, reg_i, float(reg_b), first_epoch_idx, _ptr(losses)), self._h)
        return losses

    def stage_stats(self):
        out = (C.c_int64 * 4)()
        _check(load().lrk_stage_stats(self._h, out), self._h)
        return {"ratings": out[0], "run_tile_ratings": out[1], "run_tile_share": (out[1] / out[0]) if out[0] else 0.0,
                "max_item_degree": out[2], "run_min_degree": out[3]}

    def debug_stream(self, nnz):
        """-> (su, si, sr, units[n,4] or None): the staged stream and, for the unit-ordered stream, its unit table"""
        su, si, sr = np.empty(nnz, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float32)
        n = C.c_int64()
        _check(load().lrk_debug_stream(self._h, None, None, None, None, 0, C.byref(n)), self._h)
        units = np.empty((n.value, 4), np.int32) if n.value else None
        _check(load().lrk_debug_stream(self._h, _ptr(su), _ptr(si), _ptr(sr), _ptr(units), n.value, C.byref(n)), self._h)
        return su, si, sr, units

    def last_epoch_ms(self):
        ms = C.c_float()
        _check(load().lrk_last_epoch_ms(self._h, C.byref(ms)), self._h)
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        _check(load().lrk_launch_count(self._h, C.byref(n)), self._h)
        return n.value

    def bpr_peek_samples(self, epoch_idx, first, n):
        out = np.empty((n, 11 if self.model == MODEL_GBPR else 3), np.int32)
        _check(load().lrk_bpr_peek_samples(self._h, epoch_idx, first, n, out), self._h)
        return out

    # -- prediction / ranking
    def predict_pairs(self, users, items):
        users = np.ascontiguousarray(users, np.int32)
        items = np.ascontiguousarray(items, np.int32)
        out = np.empty(users.shape[0], np.float64)
        _check(load().lrk_predict_pairs(self._h, users, items, users.shape[0], out), self._h)
        return out

    def eval_rating(self, U, t_rowptr, t_col, t_val, min_rate, max_rate, want_pred=False):
        t_rowptr = np.ascontiguousarray(t_rowptr, np.int64)
        t_col = np.ascontiguousarray(t_col, np.int32)
        t_val = np.ascontiguousarray(t_val, np.float64)
        pred = np.empty(t_col.shape[0], np.float64) if want_pred else None
        rmse, mae = C.c_double(), C.c_double()
        _check(load().lrk_eval_rating(self._h, U, _ptr(t_rowptr), _ptr(t_col), _ptr(t_val), float(min_rate),
                                      float(max_rate), _ptr(pred), C.byref(rmse), C.byref(mae)), self._h)
        return (rmse.value, mae.value, pred) if want_pred else (rmse.value, mae.value)

    def topn(self, topn, users=None, nq=None, exclude_train=True, out=None):
        """out = (items int32[nq,topn], scores float64[nq,topn], counts int32[nq]) reuses caller buffers (e.g. pinned)"""
        if users is not None:
            users = np.ascontiguousarray(users, np.int32)
            nq = users.shape[0]
        elif nq is None:
            nq = self.U
        if out is not None:
            items, scores, counts = out
            assert items.shape == (nq, topn) and scores.shape == (nq, topn) and counts.shape == (nq,)
        else:
            items = np.empty((nq, topn), np.int32)
            scores = np.empty((nq, topn), np.float64)
            counts = np.empty(nq, np.int32)
        _check(load().lrk_topn(self._h, _ptr(users), nq, topn, int(bool(exclude_train)), items, scores, counts), self._h)
        return items, scores, counts

    def eval_ranking(self, topn, t_rowptr, t_col, t_val, want_lists=False):
        """-> dict(AUC, AP, NDCG, Precision, Recall, RR, Novelty, Entropy) [, (items, scores, counts)]"""
        t_rowptr = np.ascontiguousarray(t_rowptr, np.int64)
        t_col = np.ascontiguousarray(t_col, np.int32)
        t_val = np.ascontiguousarray(t_val, np.float64)
        out = (C.c_double * 8)()
        items = np.empty((self.U, topn), np.int32) if want_lists else None
        scores = np.empty((self.U, topn), np.float64) if want_lists else None
        counts = np.empty(self.U, np.int32) if want_lists else None
        _check(load().lrk_eval_ranking(self._h, topn, _ptr(t_rowptr), _ptr(t_col), _ptr(t_val), _ptr(items), _ptr(scores),
                                       _ptr(counts), out), self._h)
        m = dict(zip(("AUC", "AP", "NDCG", "Precision", "Recall", "RR", "Novelty", "Entropy"), list(out)))
        return (m, (items, scores, counts)) if want_lists else m

    def sgd_safeguard(self):
        d, r = C.c_int32(), C.c_int64()
        _check(load().lrk_sgd_safeguard_state(self._h, C.byref(d), C.byref(r)), self._h)
        return {"conc_div": d.value, "rollbacks": r.value}

    def topn_stats(self):
        a, b, ms = C.c_int64(), C.c_int64(), C.c_float()
        _check(load().lrk_topn_stats(self._h, C.byref(a), C.byref(b), C.byref(ms)), self._h)
        ph = (C.c_float * 6)()
        _check(load().lrk_topn_phase_ms(self._h, ph), self._h)
        return {"fast_users": a.value, "fallback_users": b.value, "ms": ms.value,
                "phase_ms": {"operands": ph[0], "sweep": ph[1], "rescore": ph[2], "fallback": ph[3]},
                "sweep_error_over_bound": ph[4], "resweep_users": int(ph[5])}

    def probe_l2(self, working_set_bytes, row_floats):
        """-> dict(gather, red, mix) GB/s of row bytes through L2 (measurement aid, include/librec_b200.h)"""
        out = (C.c_double * 3)()
        _check(load().lrk_probe_l2(self._h, int(working_set_bytes), int(row_floats), out), self._h)
        return {"gather": out[0], "red": out[1], "mix": out[2]}

    # -- DSGD
    def comm_init(self, rank, world, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        _check(load().lrk_comm_init(self._h, rank, world, buf), self._h)


def comm_unique_id():
    buf = (C.c_uint8 * 128)()
    _check(load().lrk_comm_unique_id(buf))
    return bytes(buf)


def device_count():
    return load().lrk_device_count()

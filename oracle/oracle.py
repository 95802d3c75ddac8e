"""ctypes binding of the CPU oracle (oracle/lrk_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (librec_b200/) never imports this.
PARITY UNPINNED (see the header of lrk_oracle.cpp): the Java reference cannot run here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblrk_oracle.so")

BIASEDMF, PMF, BPR, RANKSGD = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "lrk_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def _opt(a):
    """numpy array or None -> void* (NULL for None)"""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    L.lro_strictmath_log.restype = C.c_double
    L.lro_strictmath_log.argtypes = [C.c_double]
    L.lro_seed.argtypes = [C.c_int64]
    L.lro_next_int.restype = C.c_int32
    L.lro_uniform_int.restype = C.c_int32
    L.lro_uniform_int.argtypes = [C.c_int32]
    L.lro_uniform.restype = C.c_double
    L.lro_next_gaussian.restype = C.c_double
    L.lro_gaussian.restype = C.c_double
    L.lro_gaussian.argtypes = [C.c_double, C.c_double]
    L.lro_gaussian_fill.argtypes = [_f64p, C.c_int64, C.c_double, C.c_double]
    L.lro_rng_get_state.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.lro_rng_set_state.argtypes = [C.c_uint64, C.c_int32, C.c_double]
    L.lro_float_promote.restype = C.c_double
    L.lro_float_promote.argtypes = [C.c_char_p]
    L.lro_csr_load_text.restype = C.c_void_p
    L.lro_csr_load_text.argtypes = [C.c_char_p, C.c_double]
    L.lro_csr_load_testset.restype = C.c_int32
    L.lro_csr_load_testset.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.lro_csr_load_paths.restype = C.c_void_p
    L.lro_csr_load_paths.argtypes = [C.c_char_p, C.c_char_p, C.c_double]
    L.lro_csr_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.lro_csr_copy.argtypes = [C.c_void_p, _i64p, _i32p, _f64p]
    L.lro_csr_outer_id.restype = C.c_int64
    L.lro_csr_outer_id.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    L.lro_csr_free.argtypes = [C.c_void_p]
    L.lro_split_ratio.argtypes = [C.c_int64, _f64p, C.c_double, _u8p]
    L.lro_matrix_setup.argtypes = [C.c_int64, _f64p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.lro_mf_setup.argtypes = [C.c_int32, C.c_int32, C.c_int32, _f64p, _f64p, C.c_void_p, C.c_void_p]
    L.lro_biasedmf_epoch.restype = C.c_double
    L.lro_biasedmf_epoch.argtypes = [C.c_int32, _i64p, _i32p, _f64p, C.c_int32, _f64p, _f64p, _f64p, _f64p, C.c_double,
                                     C.c_float, C.c_float, C.c_float, C.c_double, C.c_void_p, C.c_void_p]
    L.lro_pmf_epoch.restype = C.c_double
    L.lro_pmf_epoch.argtypes = [C.c_int32, _i64p, _i32p, _f64p, C.c_int32, _f64p, _f64p,
                                C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.lro_bpr_epoch.restype = C.c_double
    L.lro_bpr_epoch.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, C.c_int32, _f64p, _f64p,
                                C.c_float, C.c_float, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]
    L.lro_is_converged.restype = C.c_int32
    L.lro_is_converged.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_float)]
    L.lro_update_lrate.restype = C.c_float
    L.lro_update_lrate.argtypes = [C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_float, C.c_double, C.POINTER(C.c_double)]
    L.lro_train.restype = C.c_int32
    L.lro_train.argtypes = [C.c_int32, C.c_int32, C.c_int32, _i64p, _i32p, _f64p, C.c_int32, _f64p, _f64p,
                            C.c_void_p, C.c_void_p, C.c_double, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double,
                            C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    L.lro_split_ratio_item.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, C.c_double, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")]
    L.lro_split_ratio_valid.argtypes = [C.c_int64, C.c_double, C.c_double, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")]
    L.lro_split_kcv.argtypes = [C.c_int64, C.c_int32, _i32p]
    L.lro_csr_copy_dates.restype = C.c_int32
    L.lro_csr_copy_dates.argtypes = [C.c_void_p, _i64p]
    L.lro_split_by_date.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, _f64p, _i64p, C.c_int32, C.c_double, C.c_int32,
                                    np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")]
    L.lro_split_loocv.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, C.c_int32, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")]
    L.lro_split_givenn.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, C.c_int32, C.c_int32, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")]
    L.lro_svdpp_epoch.restype = C.c_double
    L.lro_svdpp_epoch.argtypes = [C.c_int32, _i64p, _i32p, _f64p, C.c_int32, _f64p, _f64p, _f64p, _f64p, _f64p, C.c_double,
                                  C.c_float, C.c_float, C.c_float, C.c_double, C.c_double]
    L.lro_svdpp_predict_pairs.argtypes = [C.c_int32, _f64p, _f64p, _f64p, _f64p, _f64p, C.c_double, _i64p, _i32p, _i32p, _i32p,
                                          C.c_int64, _f64p]
    L.lro_aobpr_train.restype = C.c_int32
    L.lro_aobpr_train.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, C.c_int32, _f64p, _f64p, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_int32, C.c_void_p, C.c_void_p]
    L.lro_gbpr_epoch.restype = C.c_double
    L.lro_gbpr_epoch.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, C.c_int32, _f64p, _f64p, _f64p, C.c_float, C.c_float, C.c_float,
                                 C.c_double, C.c_float, C.c_int32, C.c_void_p, C.c_void_p]
    L.lro_dense_inverse_export.argtypes = [_f64p, C.c_int32, _f64p]
    L.lro_wrmf_weight.restype = C.c_double
    L.lro_wrmf_weight.argtypes = [C.c_double, C.c_float]
    L.lro_wrmf_epoch.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, _f64p, C.c_int32, _f64p, _f64p, C.c_float, C.c_float]
    L.lro_eals_confidences.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, C.c_float, C.c_float, C.c_int32, _f64p]
    L.lro_eals_weight.restype = C.c_double
    L.lro_eals_weight.argtypes = [C.c_double, C.c_float, C.c_int32]
    L.lro_eals_epoch.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, _f64p, C.c_int32, _f64p, _f64p, _f64p, C.c_float, C.c_float]
    L.lro_ranksgd_item_probs.restype = C.c_int32
    L.lro_ranksgd_item_probs.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, _i32p, _f64p]
    L.lro_ranksgd_epoch.restype = C.c_double
    L.lro_ranksgd_epoch.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, _f64p, C.c_int32, _f64p, _f64p, C.c_float,
                                    C.c_int64, C.c_void_p, C.c_void_p]
    L.lro_eval_rating.argtypes = [C.c_int32, C.c_int32, _i64p, _i32p, _f64p, C.c_int32, _f64p, _f64p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
    L.lro_eval_ranking.argtypes = [C.c_int32, C.c_int32, _i32p, _i32p, _i64p, _i32p, _f64p, _i32p, _i32p, C.c_int32, _f64p]
    L.lro_eval_ranking_extra.argtypes = [C.c_int32, C.c_int32, _i32p, _i32p, _i64p, _i32p, _f64p]
    L.lro_predict_pairs.argtypes = [C.c_int32, C.c_int32, _f64p, _f64p, C.c_void_p, C.c_void_p, C.c_double,
                                    _i32p, _i32p, C.c_int64, _f64p]
    L.lro_recommend_rank.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f64p, _f64p, C.c_void_p, C.c_void_p,
                                     C.c_double, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                     _i32p, _f64p, _i32p, C.c_int32]
    L.lro_heap_trace.restype = C.c_int32
    L.lro_heap_trace.argtypes = [_f64p, C.c_int32, C.c_int32, _i32p]
    L.lro_sgd_epoch_hogwild_f32.restype = C.c_double
    L.lro_sgd_epoch_hogwild_f32.argtypes = [C.c_int32, _i32p, _i32p, _f32p, C.c_int64, C.c_int32, _f32p, _f32p,
                                            C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32]
    L.lro_max_threads.restype = C.c_int32
    L.lro_version.restype = C.c_char_p
    _lib = L
    return L


# ------------------------------------------------------------------ convenience layer
class Csr:
    """flat CSR: rowptr int64[U+1], col int32[nnz], val float64[nnz]"""

    def __init__(self, U, I, rowptr, col, val):
        self.U, self.I = int(U), int(I)
        self.rowptr = np.ascontiguousarray(rowptr, np.int64)
        self.col = np.ascontiguousarray(col, np.int32)
        self.val = np.ascontiguousarray(val, np.float64)

    @property
    def nnz(self):
        return int(self.col.shape[0])

    def rows(self):
        return np.repeat(np.arange(self.U, dtype=np.int32), np.diff(self.rowptr)).astype(np.int32)

    def select(self, mask):
        rows = self.rows()[mask]
        rowptr = np.zeros(self.U + 1, np.int64)
        np.add.at(rowptr, rows.astype(np.int64) + 1, 1)
        return Csr(self.U, self.I, np.cumsum(rowptr), self.col[mask], self.val[mask])


def load_text(path, bin_thold=-1.0, column_format="UIR"):
    """path: file, directory, or ':'-separated list of them (data.input.path, already prefixed with dfs.data.dir)"""
    L = lib()
    h = L.lro_csr_load_paths(path.encode(), column_format.encode(), float(bin_thold))
    if not h:
        raise FileNotFoundError(path)
    U, I, n = C.c_int32(), C.c_int32(), C.c_int64()
    L.lro_csr_dims(h, C.byref(U), C.byref(I), C.byref(n))
    rowptr = np.zeros(U.value + 1, np.int64)
    col = np.zeros(n.value, np.int32)
    val = np.zeros(n.value, np.float64)
    L.lro_csr_copy(h, rowptr, col, val)
    date = np.zeros(n.value, np.int64)
    has_date = L.lro_csr_copy_dates(h, date)
    L.lro_csr_free(h)
    m = Csr(U.value, I.value, rowptr, col, val)
    m.date = date if has_date else None            # datetime matrix values aligned with the entries (UIRT)
    return m


def _take(h):
    L = lib()
    U, I, n = C.c_int32(), C.c_int32(), C.c_int64()
    L.lro_csr_dims(h, C.byref(U), C.byref(I), C.byref(n))
    rowptr = np.zeros(U.value + 1, np.int64); col = np.zeros(n.value, np.int32); val = np.zeros(n.value, np.float64)
    L.lro_csr_copy(h, rowptr, col, val)
    L.lro_csr_free(h)
    return Csr(U.value, I.value, rowptr, col, val)


def load_testset(path, test_path, bin_thold=-1.0, column_format="UIR"):
    """data.model.splitter=testset (GivenTestSetDataSplitter): -> (preference, train, test), all with the final dimensions"""
    p, tr, te = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ok = lib().lro_csr_load_testset(path.encode(), test_path.encode(), column_format.encode(), float(bin_thold),
                                    C.byref(p), C.byref(tr), C.byref(te))
    if not ok:
        raise FileNotFoundError(path + " / " + test_path)
    return _take(p), _take(tr), _take(te)


def split_ratio(csr, ratio=0.8):
    """RatioDataSplitter.getRatioByRating on the global RNG -> (train Csr, test Csr)"""
    flags = np.zeros(csr.nnz, np.uint8)
    lib().lro_split_ratio(csr.nnz, csr.val, float(ratio), flags)
    return csr.select(flags == 1), csr.select(flags == 0)


def _two_way(csr, flags):
    """(train, test) from per-entry flags; exact zeros vanish from both (reshape())"""
    nz = csr.val != 0.0
    return csr.select((flags == 1) & nz), csr.select((flags == 0) & nz)


def split(csr, splitter="ratio", by="rating", ratio=0.8, n_given=1, k_fold=5, valid_ratio=0.0):
    """the reference's splitters on the global RNG.  ratio / loocv / givenn -> (train, test); kcv -> list of (train, test)"""
    L = lib()
    flags = np.zeros(csr.nnz, np.uint8)
    if splitter == "ratio" and by.endswith("date"):
        L.lro_split_by_date(csr.U, csr.I, csr.rowptr, csr.col, csr.val, csr.date, {"ratingdate": 0, "userdate": 1, "itemdate": 2}[by],
                            ratio, n_given, flags)
        return _two_way(csr, flags)
    if splitter == "ratio" and by == "valid":            # -> (train, valid, test)
        L.lro_split_ratio_valid(csr.nnz, ratio, valid_ratio, flags)
        nz = csr.val != 0.0
        return tuple(csr.select((flags == w) & nz) for w in (0, 1, 2))
    if splitter == "ratio":
        if by in ("rating", "user"):
            L.lro_split_ratio(csr.nnz, csr.val, ratio, flags)
        elif by == "item":
            L.lro_split_ratio_item(csr.U, csr.I, csr.rowptr, csr.col, ratio, flags)
        else:
            raise ValueError(by)
        return _two_way(csr, flags)
    date_modes = {("ratio", "ratingdate"): 0, ("ratio", "userdate"): 1, ("ratio", "itemdate"): 2, ("loocv", "userdate"): 3,
                  ("loocv", "itemdate"): 4, ("givenn", "userdate"): 5, ("givenn", "itemdate"): 6}
    if (splitter, by) in date_modes:
        L.lro_split_by_date(csr.U, csr.I, csr.rowptr, csr.col, csr.val, csr.date, date_modes[(splitter, by)], ratio, n_given, flags)
        return _two_way(csr, flags)
    if splitter == "loocv":
        L.lro_split_loocv(csr.U, csr.I, csr.rowptr, csr.col, int(by == "item"), flags)
        return _two_way(csr, flags)
    if splitter == "givenn":
        L.lro_split_givenn(csr.U, csr.I, csr.rowptr, csr.col, int(by == "item"), n_given, flags)
        return _two_way(csr, flags)
    if splitter == "kcv":
        fold = np.zeros(csr.nnz, np.int32)
        L.lro_split_kcv(csr.nnz, k_fold, fold)
        return [_two_way(csr, (fold != k).astype(np.uint8)) for k in range(1, min(k_fold, csr.nnz) + 1)]
    raise ValueError(splitter)


def matrix_setup(csr):
    mu, mn, mx = C.c_double(), C.c_double(), C.c_double()
    lib().lro_matrix_setup(csr.nnz, csr.val, C.byref(mu), C.byref(mn), C.byref(mx))
    return mu.value, mn.value, mx.value


def mf_setup(U, I, k, biased):
    P = np.zeros((U, k), np.float64)
    Q = np.zeros((I, k), np.float64)
    bu = np.zeros(U, np.float64) if biased else None
    bi = np.zeros(I, np.float64) if biased else None
    lib().lro_mf_setup(U, I, k, P, Q, _opt(bu), _opt(bi))
    return P, Q, bu, bi


def train(model, tr, k, P, Q, bu, bi, mu, lr, max_lr, reg_u, reg_i, reg_b, num_iter,
          early_stop=False, bold_driver=False, decay=1.0):
    losses = np.zeros(num_iter, np.float64)
    done = lib().lro_train(model, tr.U, tr.I, tr.rowptr, tr.col, tr.val, k, P, Q, _opt(bu), _opt(bi), float(mu),
                           lr, max_lr, reg_u, reg_i, float(reg_b), num_iter, int(early_stop), int(bold_driver),
                           decay, _opt(losses))
    return done, losses


def eval_rating(model, te, k, P, Q, bu, bi, mu, min_rate, max_rate, want_pred=False):
    rmse, mae = C.c_double(), C.c_double()
    pred = np.zeros(te.nnz, np.float64) if want_pred else None
    lib().lro_eval_rating(model, te.U, te.rowptr, te.col, te.val, k, P, Q, _opt(bu), _opt(bi), float(mu),
                          float(min_rate), float(max_rate), C.byref(rmse), C.byref(mae), _opt(pred))
    return (rmse.value, mae.value, pred) if want_pred else (rmse.value, mae.value)


RANKING_MEASURES = ("AUC", "AP", "NDCG", "Precision", "Recall", "RR", "Novelty", "Entropy")


def eval_ranking(te, tr, topn, items, counts):
    """ranking evaluators of the reference over top-N lists for ALL users -> dict over RANKING_MEASURES"""
    out = np.zeros(8, np.float64)
    num_dropped = (te.I - np.diff(tr.rowptr)).astype(np.int32)
    purchased = (np.bincount(tr.col, minlength=te.I) + np.bincount(te.col, minlength=te.I)).astype(np.int32)
    lib().lro_eval_ranking(te.U, topn, np.ascontiguousarray(items, np.int32), np.ascontiguousarray(counts, np.int32),
                           te.rowptr, te.col, te.val, num_dropped, purchased, te.I, out)
    return dict(zip(RANKING_MEASURES, out.tolist()))


def eval_ranking_extra(te, topn, items, counts):
    """HitRate (NaN unless leave-one-out), ARHR, IDCG -- the ranking evaluators outside the default list"""
    out = np.zeros(3, np.float64)
    lib().lro_eval_ranking_extra(te.U, topn, np.ascontiguousarray(items, np.int32), np.ascontiguousarray(counts, np.int32),
                                 te.rowptr, te.col, out)
    return dict(zip(("HitRate", "ARHR", "IDCG"), out.tolist()))


def recommend_rank(model, U, I, k, P, Q, bu, bi, mu, tr, topn, users=None, nthreads=None):
    nq = U if users is None else len(users)
    items = np.full((nq, topn), -1, np.int32)
    scores = np.zeros((nq, topn), np.float64)
    counts = np.zeros(nq, np.int32)
    if users is not None:
        users = np.ascontiguousarray(users, np.int32)
    nt = nthreads or lib().lro_max_threads()
    lib().lro_recommend_rank(model, U, I, k, P, Q, _opt(bu), _opt(bi), float(mu),
                             _opt(tr.rowptr) if tr is not None else None, _opt(tr.col) if tr is not None else None,
                             topn, _opt(users), nq, items, scores, counts, nt)
    return items, scores, counts

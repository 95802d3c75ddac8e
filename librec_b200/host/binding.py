"""ctypes access to the C++ host mirror (RecommenderJob and friends) for the tests and examples."""
import ctypes as C

import numpy as np

from . import build as _hb

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    _hb.build()
    L = C.CDLL(_hb.LIB_PATH)
    L.lrh_last_error.restype = C.c_char_p
    L.lrh_job_create.restype = C.c_void_p
    L.lrh_job_create.argtypes = [C.c_char_p]
    L.lrh_job_destroy.argtypes = [C.c_void_p]
    L.lrh_job_set_data.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 6
    L.lrh_job_run.argtypes = [C.c_void_p]
    L.lrh_job_metric.restype = C.c_double
    L.lrh_job_metric.argtypes = [C.c_void_p, C.c_char_p]
    L.lrh_job_log.restype = C.c_char_p
    L.lrh_job_log.argtypes = [C.c_void_p]
    L.lrh_job_factors.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.POINTER(C.c_double)]
    L.lrh_job_list_size.restype = C.c_int64
    L.lrh_job_list_size.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.lrh_job_list_copy.argtypes = [C.c_void_p] * 4
    L.lrh_datamodel_build.restype = C.c_void_p
    L.lrh_datamodel_build.argtypes = [C.c_char_p]
    L.lrh_datamodel_destroy.argtypes = [C.c_void_p]
    L.lrh_datamodel_next_fold.argtypes = [C.c_void_p]
    L.lrh_datamodel_next_fold.restype = C.c_int
    L.lrh_datamodel_num_folds.argtypes = [C.c_void_p]
    L.lrh_datamodel_num_folds.restype = C.c_int
    L.lrh_datamodel_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.lrh_datamodel_copy.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
    L.lrh_datamodel_raw_id.restype = C.c_char_p
    L.lrh_datamodel_raw_id.argtypes = [C.c_void_p, C.c_int, C.c_int32]
    L.lrh_job_save_result.restype = C.c_char_p
    L.lrh_job_save_result.argtypes = [C.c_void_p]
    L.lrh_randoms_seed.argtypes = [C.c_longlong]
    L.lrh_randoms_uniform_int.restype = C.c_int
    L.lrh_randoms_uniform_int.argtypes = [C.c_int]
    L.lrh_randoms_uniform.restype = C.c_double
    L.lrh_randoms_gaussian.restype = C.c_double
    L.lrh_randoms_gaussian.argtypes = [C.c_double, C.c_double]
    L.lrh_format_double.argtypes = [C.c_double, C.c_char_p, C.c_int]
    L.lrh_format_float.argtypes = [C.c_float, C.c_char_p, C.c_int]
    L.lrh_conf_probe.restype = C.c_double
    L.lrh_conf_probe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double]
    _lib = L
    return L


class LibrecException(Exception):
    pass


class TextDataModel:
    """data/model/TextDataModel.java: properties (dfs.data.dir, data.input.path, data.column.format,
    data.convert.binarize.threshold, data.model.splitter = ratio | kcv | loocv | givenn with their keys, rec.random.seed)
    -> preference / train / test as flat CSR; next_fold() walks the folds (one for everything but kcv)"""

    def __init__(self, properties):
        text = properties if isinstance(properties, str) else "\n".join("%s=%s" % kv for kv in properties.items())
        self._L = load()
        self._h = self._L.lrh_datamodel_build(text.encode())
        if not self._h:
            raise LibrecException(self._L.lrh_last_error().decode())

    def matrix(self, which):
        """which: 'preference' | 'train' | 'test' | 'valid' -> (U, I, rowptr, col, val)"""
        w = {"preference": 0, "train": 1, "test": 2, "valid": 3}[which]
        U, I, n = C.c_int32(), C.c_int32(), C.c_int64()
        self._L.lrh_datamodel_dims(self._h, w, C.byref(U), C.byref(I), C.byref(n))
        rowptr = np.zeros(U.value + 1, np.int64); col = np.zeros(n.value, np.int32); val = np.zeros(n.value, np.float64)
        self._L.lrh_datamodel_copy(self._h, w, rowptr.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p))
        return U.value, I.value, rowptr, col, val

    def next_fold(self):
        """AbstractDataModel.hasNextFold(): True = matrix('train') / matrix('test') now hold the next fold"""
        r = self._L.lrh_datamodel_next_fold(self._h)
        if r < 0:
            raise LibrecException(self._L.lrh_last_error().decode())
        return bool(r)

    @property
    def num_folds(self):
        return self._L.lrh_datamodel_num_folds(self._h)

    def raw_id(self, is_item, inner):
        return self._L.lrh_datamodel_raw_id(self._h, int(is_item), inner).decode()

    def close(self):
        if self._h:
            self._L.lrh_datamodel_destroy(self._h)
            self._h = None

    __del__ = close


class RecommenderJob:
    """new RecommenderJob(conf).runJob() -- job/RecommenderJob.java:72-90 (data model handed in as flat CSR)"""

    def __init__(self, properties):
        text = properties if isinstance(properties, str) else "\n".join("%s=%s" % kv for kv in properties.items())
        self._L = load()
        self._h = self._L.lrh_job_create(text.encode())
        if not self._h:
            raise LibrecException(self._L.lrh_last_error().decode())
        self._keep = []

    def set_data(self, U, I, train, test):
        arrs = []
        for m in (train, test):
            arrs += [np.ascontiguousarray(m.rowptr, np.int64), np.ascontiguousarray(m.col, np.int32), np.ascontiguousarray(m.val, np.float64)]
        self._keep = arrs
        self.U, self.I = U, I
        rc = self._L.lrh_job_set_data(self._h, U, I, *[a.ctypes.data_as(C.c_void_p) for a in arrs])
        if rc:
            raise LibrecException(self._L.lrh_last_error().decode())

    def run_job(self):
        rc = self._L.lrh_job_run(self._h)
        if rc == -1:
            raise LibrecException(self._L.lrh_last_error().decode())
        if rc:
            raise IndexError(self._L.lrh_last_error().decode())

    def save_result(self):
        p = self._L.lrh_job_save_result(self._h)
        if p is None:
            raise LibrecException(self._L.lrh_last_error().decode())
        return p.decode()

    def metric(self, name):
        return self._L.lrh_job_metric(self._h, name.encode())

    def log(self):
        return self._L.lrh_job_log(self._h).decode().splitlines()

    def factors(self, k, biased):
        P = np.zeros((self.U, k)); Q = np.zeros((self.I, k))
        bu = np.zeros(self.U) if biased else None
        bi = np.zeros(self.I) if biased else None
        mu = C.c_double()
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self._L.lrh_job_factors(self._h, p(P), p(Q), p(bu), p(bi), C.byref(mu))
        return P, Q, bu, bi, mu.value

    def recommended_list(self):
        n_ctx = C.c_int32()
        n = self._L.lrh_job_list_size(self._h, C.byref(n_ctx))
        counts = np.zeros(n_ctx.value, np.int32); keys = np.zeros(n, np.int32); vals = np.zeros(n, np.float64)
        self._L.lrh_job_list_copy(self._h, counts.ctypes.data_as(C.c_void_p), keys.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p))
        return counts, keys, vals

    def close(self):
        if self._h:
            self._L.lrh_job_destroy(self._h)
            self._h = None

    __del__ = close

"""ctypes binding of baseline/libcpu_gamma.so (cpu_gamma.c): the CPU arm of bench.py -- gamma's CPU search
path restated for speed (FMA / AVX-512 dispatch, blocked sgemm coarse search), timed on the GPU box's host
cores.  Not the correctness checker (oracle/ is) and never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcpu_gamma.so")
METRIC_IP, METRIC_L2 = 0, 1
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "cpu_gamma.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.cg_isa.restype = C.c_char_p
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bm(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def physical_cores():
    """Cores this process may really use: min(affinity mask, cgroup cpu.max quota), not omp_get_max_threads()."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    return n


def set_threads(n=None):
    """Use n OpenMP threads (default: every core the cgroup grants), whatever OMP_NUM_THREADS says --
    torchrun exports OMP_NUM_THREADS=1 to its ranks."""
    n = n or physical_cores()
    lib().cg_set_num_threads(C.c_int(n))
    return n


def num_threads():
    return int(lib().cg_num_threads())


def isa():
    return lib().cg_isa().decode()


def flat_search(db, xq, k, metric, del_bitmap=None, filter_bitmap=None, min_score=-3.4028235e38, max_score=3.4028235e38):
    db, xq = _f32(db), _f32(xq)
    n, d = db.shape
    nq = xq.shape[0]
    dis, ids = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    delb, filb = _bm(del_bitmap), _bm(filter_bitmap)
    rc = lib().cg_flat_search(_p(db), C.c_int64(d), C.c_int64(n), C.c_int(d), _p(xq), C.c_int(nq), C.c_int(k), C.c_int(metric),
                              _p(delb), _p(filb), C.c_float(min_score), C.c_float(max_score), _p(dis), _p(ids))
    assert rc == 0
    return dis, ids


def coarse_search(centroids, xq, nprobe, metric):
    centroids, xq = _f32(centroids), _f32(xq)
    L, d = centroids.shape
    nq = xq.shape[0]
    dis, ids = np.empty((nq, nprobe), np.float32), np.empty((nq, nprobe), np.int64)
    rc = lib().cg_coarse_search(_p(centroids), C.c_int(L), C.c_int(d), _p(xq), C.c_int(nq), C.c_int(nprobe), C.c_int(metric),
                                _p(dis), _p(ids))
    assert rc == 0
    return dis, ids


def ivfflat_search_preassigned(list_off, list_vecs, list_ids, xq, k, keys, metric, del_bitmap=None, filter_bitmap=None,
                               min_score=-3.4028235e38, max_score=3.4028235e38):
    list_off = np.ascontiguousarray(list_off, np.int64)
    list_vecs, xq = _f32(list_vecs), _f32(xq)
    list_ids, keys = np.ascontiguousarray(list_ids, np.int64), np.ascontiguousarray(keys, np.int64)
    nq, nprobe = keys.shape
    d = xq.shape[1]
    dis, ids = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    delb, filb = _bm(del_bitmap), _bm(filter_bitmap)
    rc = lib().cg_ivfflat_search_preassigned(_p(list_off), _p(list_vecs), _p(list_ids), C.c_int(len(list_off) - 1), C.c_int(d),
                                             _p(xq), C.c_int(nq), C.c_int(k), _p(keys), C.c_int(nprobe), C.c_int(metric),
                                             _p(delb), _p(filb), C.c_float(min_score), C.c_float(max_score), _p(dis), _p(ids))
    assert rc == 0
    return dis, ids


def ivfpq_search_preassigned(list_off, list_codes, list_ids, coarse, pq, T, xq, k, keys, coarse_dis, metric, recall_num=0,
                             raw=None, del_bitmap=None, filter_bitmap=None, min_score=-3.4028235e38, max_score=3.4028235e38):
    list_off = np.ascontiguousarray(list_off, np.int64)
    list_codes = np.ascontiguousarray(list_codes, np.uint8)
    list_ids, keys = np.ascontiguousarray(list_ids, np.int64), np.ascontiguousarray(keys, np.int64)
    coarse, pq, xq, coarse_dis = _f32(coarse), _f32(pq), _f32(xq), _f32(coarse_dis)
    T = None if T is None else _f32(T)
    raw = None if raw is None else _f32(raw)
    nq, nprobe = keys.shape
    d = xq.shape[1]
    M = list_codes.shape[1]
    dis, ids = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    delb, filb = _bm(del_bitmap), _bm(filter_bitmap)
    rc = lib().cg_ivfpq_search_preassigned(_p(list_off), _p(list_codes), _p(list_ids), C.c_int(len(list_off) - 1), C.c_int(d),
                                           C.c_int(M), _p(coarse), _p(pq), _p(T), _p(xq), C.c_int(nq), C.c_int(k), _p(keys),
                                           _p(coarse_dis), C.c_int(nprobe), C.c_int(metric), C.c_int(recall_num), _p(raw),
                                           C.c_int64(raw.shape[1] if raw is not None else 0), _p(delb), _p(filb),
                                           C.c_float(min_score), C.c_float(max_score), _p(dis), _p(ids))
    assert rc == 0, rc
    return dis, ids

"""ctypes binding of oracle/_ref/libshine_ref.so — TEST INFRASTRUCTURE ONLY.

The library is the reference's own src/hnsw/hnsw.hh (insert :40, knn :253) compiled unmodified with
oracle/shim/ standing in for ibverbs (see oracle/ref_harness.cc).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libshine_ref.so")

STAT_FIELDS = ("distcomps", "rdma_reads_in_bytes", "rdma_writes_in_bytes", "processed", "remote_allocations",
               "allocation_size", "visited_nodes", "visited_nodes_l0", "visited_neighborlists", "max_level",
               "cache_hits", "cache_misses")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.shine_ref_dist.restype = C.c_float
        _lib.shine_ref_stats_words.restype = C.c_uint32
        _lib.shine_ref_select_heuristic.restype = C.c_uint32
        assert _lib.shine_ref_stats_words() == len(STAT_FIELDS)
    return _lib


def _stats_dict(arr):
    return {k: int(v) for k, v in zip(STAT_FIELDS, arr)}


def build(base, m=16, efc=200, seed=1234, ip=False, threads=1, coroutines=4, num_mn=1, quiet=True):
    """Returns (list of dump bytes objects, one per memory node; stats dict; seconds)."""
    base = np.ascontiguousarray(base, dtype=np.float32)
    n, dim = base.shape
    dumps = (C.POINTER(C.c_uint8) * num_mn)()
    sizes = (C.c_uint64 * num_mn)()
    stats = (C.c_uint64 * len(STAT_FIELDS))()
    secs = C.c_double()
    rc = lib().shine_ref_build(base.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(dim), C.c_uint32(m),
                               C.c_uint32(efc), C.c_uint32(seed), C.c_int(int(ip)), C.c_uint32(threads),
                               C.c_uint32(coroutines), C.c_uint32(num_mn), dumps, sizes, stats, C.byref(secs),
                               C.c_int(int(quiet)))
    if rc != 0:
        raise RuntimeError(f"shine_ref_build failed: {rc}")
    out = []
    for i in range(num_mn):
        out.append(C.string_at(dumps[i], sizes[i]))
        lib().shine_ref_free(dumps[i])
    return out, _stats_dict(stats), secs.value


def search(dumps, dim, m, queries, k, ef, ip=False, threads=1, coroutines=4, cache_ratio_pct=0,
           per_query_stats=False, quiet=True):
    """Returns (ids [nq,k] u32 in the reference's heap-array order, dists [nq,k] f32, counts [nq], stats, seconds).

    stats is a dict of totals, or with per_query_stats a dict of np.uint64 arrays of length nq."""
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    nq = queries.shape[0]
    num_mn = len(dumps)
    bufs = [np.frombuffer(d, dtype=np.uint8) for d in dumps]
    ptrs = (C.c_void_p * num_mn)(*[b.ctypes.data for b in bufs])
    sizes = (C.c_uint64 * num_mn)(*[b.size for b in bufs])
    ids = np.empty((nq, k), dtype=np.uint32)
    dists = np.empty((nq, k), dtype=np.float32)
    counts = np.zeros(nq, dtype=np.uint32)
    nstat = nq if per_query_stats else 1
    stats = np.zeros((nstat, len(STAT_FIELDS)), dtype=np.uint64)
    secs = C.c_double()
    rc = lib().shine_ref_search(ptrs, sizes, C.c_uint32(num_mn), C.c_uint32(dim), C.c_uint32(m), C.c_uint32(k),
                                C.c_uint32(ef), C.c_int(int(ip)), queries.ctypes.data_as(C.c_void_p), C.c_uint32(nq),
                                C.c_uint32(threads), C.c_uint32(coroutines), C.c_uint32(cache_ratio_pct),
                                C.c_int(int(per_query_stats)), ids.ctypes.data_as(C.c_void_p),
                                dists.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                                stats.ctypes.data_as(C.c_void_p), C.byref(secs), C.c_int(int(quiet)))
    if rc != 0:
        raise RuntimeError(f"shine_ref_search failed: {rc}")
    if per_query_stats:
        st = {k_: stats[:, i].copy() for i, k_ in enumerate(STAT_FIELDS)}
    else:
        st = _stats_dict(stats[0])
    return ids, dists, counts, st, secs.value


def dist(a, b, ip=False):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(lib().shine_ref_dist(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                      C.c_uint32(a.size), C.c_int(int(ip))))


def select_heuristic(uids, dists, vectors, m, ip=False):
    """The reference's HNSW<D>::select_heuristic (hnsw.hh:482-522): returns (selected candidate indices in the
    heap-array order the reference leaves them in, distcomps)."""
    uids = np.ascontiguousarray(uids, np.uint32)
    dists = np.ascontiguousarray(dists, np.float32)
    vectors = np.ascontiguousarray(vectors, np.float32)
    c, dim = vectors.shape
    sel = np.empty(max(c, 1), np.uint32)
    dc = C.c_uint64()
    n = lib().shine_ref_select_heuristic(uids.ctypes.data_as(C.c_void_p), dists.ctypes.data_as(C.c_void_p),
                                         vectors.ctypes.data_as(C.c_void_p), C.c_uint32(c), C.c_uint32(dim), C.c_uint32(m),
                                         C.c_int(int(ip)), sel.ctypes.data_as(C.c_void_p), C.byref(dc))
    return sel[:n].copy(), dc.value

"""ctypes binding of oracle/liboracle.so (the plain-C restatement, oracle/hnsw_oracle.c).

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

COUNTER_FIELDS = ("distcomps", "visited_nodes", "visited_nodes_l0", "lists_l0", "lists_upper",
                  "rdma_reads_in_bytes", "tie")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_dist.restype = C.c_float
        _lib.orc_load.restype = C.c_void_p
        _lib.orc_num_nodes.restype = C.c_uint32
        _lib.orc_entry_row.restype = C.c_uint32
        _lib.orc_max_level.restype = C.c_uint32
        _lib.orc_neighbors.restype = C.c_uint32
        _lib.orc_select_heuristic.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def dist(a, b, ip=False):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(lib().orc_dist(_p(a), _p(b), C.c_uint32(a.size), C.c_int(int(ip))))


class Index:
    """Parsed index dump(s) in the reference format."""

    def __init__(self, dumps, dim, m):
        bufs = [np.frombuffer(d, dtype=np.uint8) for d in dumps]
        ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        sizes = (C.c_uint64 * len(bufs))(*[b.size for b in bufs])
        self.dim, self.m = dim, m
        self._h = lib().orc_load(ptrs, sizes, C.c_uint32(len(bufs)), C.c_uint32(dim), C.c_uint32(m))
        if not self._h:
            raise ValueError("not an index dump for this dim / m")
        self.n = lib().orc_num_nodes(C.c_void_p(self._h))
        self.entry_row = lib().orc_entry_row(C.c_void_p(self._h))
        self.max_level = lib().orc_max_level(C.c_void_p(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_free(C.c_void_p(self._h))
            self._h = None

    def export(self):
        n, d, m = self.n, self.dim, self.m
        out = dict(uid=np.empty(n, np.uint32), level=np.empty(n, np.uint32), vectors=np.empty((n, d), np.float32),
                   l0_cnt=np.empty(n, np.uint32), l0_adj=np.empty((n, 2 * m), np.uint32))
        lib().orc_export(C.c_void_p(self._h), _p(out["uid"]), _p(out["level"]), _p(out["vectors"]),
                         _p(out["l0_cnt"]), _p(out["l0_adj"]))
        return out

    def neighbors(self, row, level):
        buf = np.empty(2 * self.m, np.uint32)
        c = lib().orc_neighbors(C.c_void_p(self._h), C.c_uint32(row), C.c_uint32(level), _p(buf))
        return None if c == 0xFFFFFFFF else buf[:c].copy()

    def knn(self, queries, k, ef, ip=False, counters=False, track_ties=False, threads=1):
        """ids/dists [nq,k] in the reference's heap-array order (hnsw.hh:300-303), counts [nq], counters dict."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, k), np.uint32)
        dists = np.empty((nq, k), np.float32)
        counts = np.empty(nq, np.uint32)
        ct = np.zeros((nq, len(COUNTER_FIELDS)), np.uint64) if counters or track_ties else None
        rc = lib().orc_knn(C.c_void_p(self._h), _p(q), C.c_uint32(nq), C.c_uint32(k), C.c_uint32(ef),
                           C.c_int(int(ip)), _p(ids), _p(dists), _p(counts),
                           _p(ct) if ct is not None else None, C.c_int(int(track_ties)), C.c_int(threads))
        if rc != 0:
            raise ValueError("orc_knn: invalid arguments (ef >= k required, hnsw.hh:36)")
        cd = {f: ct[:, i].copy() for i, f in enumerate(COUNTER_FIELDS)} if ct is not None else None
        return ids, dists, counts, cd


def select_heuristic(uids, dists, vectors, m, ip=False):
    uids = np.ascontiguousarray(uids, np.uint32)
    dists = np.ascontiguousarray(dists, np.float32)
    vectors = np.ascontiguousarray(vectors, np.float32)
    c, dim = vectors.shape
    sel = np.empty(max(c, 1), np.uint32)
    dc = C.c_uint64()
    n = lib().orc_select_heuristic(_p(uids), _p(dists), _p(vectors), C.c_uint32(c), C.c_uint32(dim), C.c_uint32(m),
                                   C.c_int(int(ip)), _p(sel), C.byref(dc))
    return sel[:n].copy(), dc.value


def sorted_results(ids, dists):
    """Sort each row ascending by (distance, id): the canonical form results are compared in."""
    order = np.lexsort((ids, dists), axis=1)
    return np.take_along_axis(ids, order, 1), np.take_along_axis(dists, order, 1)


def build(base, m=16, efc=200, seed=1234, ip=False):
    """HNSW::insert restated (single thread / single coroutine / one memory node).  Returns (dump bytes, distcomps)."""
    base = np.ascontiguousarray(base, dtype=np.float32)
    n, dim = base.shape
    dump = C.POINTER(C.c_uint8)()
    size = C.c_uint64()
    dc = C.c_uint64()
    rc = lib().orc_build(_p(base), C.c_uint32(n), C.c_uint32(dim), C.c_uint32(m), C.c_uint32(efc), C.c_uint32(seed),
                         C.c_int(int(ip)), C.byref(dump), C.byref(size), C.byref(dc))
    if rc != 0:
        raise ValueError("orc_build: invalid arguments")
    out = C.string_at(dump, size.value)
    lib().orc_free_buffer(dump)
    return out, dc.value

"""ctypes binding of libshn_b200.so (include/shn.h).  No compute happens in Python; there is no CPU fallback:
every entry point raises ShnError when the library is missing or no sm_100 device is usable."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

L2, IP = 0, 1
INVALID_ID = 0xFFFFFFFF


class ShnError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"shn error {code}: {message}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("distcomps", "visited_nodes", "visited_nodes_l0", "visited_neighborlists",
                                          "lists_l0", "lists_upper", "algorithmic_bytes", "reference_layout_bytes",
                                          "overflow_queries", "processed")] + \
               [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double)] + \
               [(n, C.c_uint64) for n in ("rows_hot", "rows_local", "rows_remote", "rows_halo")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def library_path():
    return os.environ.get("SHN_LIB") or os.path.join(_HERE, "libshn_b200.so")


def build_library():
    """Compile the CUDA library in-tree (nvcc cross-compiles for sm_100a without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc")])


def lib():
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise ShnError(-3, f"{path} is missing: run __graft_entry__.build() (there is no CPU path)")
        L = C.CDLL(path)
        L.shn_last_error.restype = C.c_char_p
        L.shn_version.restype = C.c_char_p
        for f in ("shn_index_size", "shn_index_hbm_bytes", "shn_index_dump_bytes"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("shn_index_dim", "shn_index_m", "shn_index_max_level"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [C.c_void_p]
        L.shn_index_free.argtypes = [C.c_void_p]
        L.shn_index_free.restype = None
        L.shn_index_load.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_char_p), C.c_int, C.c_uint32, C.c_uint32,
                                     C.c_int, C.c_int]
        L.shn_index_load_mem.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int,
                                         C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.shn_index_store.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int]
        L.shn_index_store_mem.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int]
        L.shn_dump_repartition.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int, C.c_uint32, C.c_uint32,
                                           C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.shn_index_build.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.c_int, C.c_uint32, C.c_int]
        L.shn_index_build_device.argtypes = L.shn_index_build.argtypes
        L.shn_index_build_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.shn_bruteforce_topk.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32,
                                          C.c_void_p, C.c_void_p, C.c_int]
        L.shn_bruteforce_topk_device.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int,
                                                 C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.shn_index_count_visits.argtypes = [C.c_void_p, C.c_int]
        L.shn_index_visit_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.shn_index_partition.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
        L.shn_placement_fit.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.shn_route_queries.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.c_double, C.c_void_p, C.c_int]
        L.shn_index_partition_export.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.shn_index_partition_attach.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.shn_index_partition_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.shn_index_partition_build_halo.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
        L.shn_router_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_double, C.c_uint64, C.c_uint32]
        L.shn_router_free.argtypes = [C.c_void_p]
        L.shn_router_free.restype = None
        L.shn_router_export.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.shn_router_attach.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
        L.shn_router_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.shn_router_search.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(Stats)]
        L.shn_router_results.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.shn_router_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.shn_router_destinations.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.shn_draw_levels.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.shn_set_build_option.argtypes = [C.c_char_p, C.c_int64]
        L.shn_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.shn_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                 C.POINTER(Stats)]
        L.shn_search_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise ShnError(rc, lib().shn_last_error().decode())


class Index:
    """An HNSW index resident in the HBM of one B200 (handle of include/shn.h)."""

    def __init__(self, handle, metric):
        self._h = handle
        self.metric = metric

    # -- lifetime ----------------------------------------------------------------------------------------------
    @classmethod
    def load(cls, paths, dim, m, ip=False, gpu=0):
        """From dump/index_m<M>_efc<efC>_node<i>_of<n>.dat files (reference format)."""
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        h = C.c_void_p()
        _check(lib().shn_index_load(C.byref(h), arr, len(paths), dim, m, IP if ip else L2, gpu))
        return cls(h, IP if ip else L2)

    @classmethod
    def from_dumps(cls, dumps, dim, m, ip=False, gpu=0):
        """From dumps held in memory (bytes-like, one per memory node)."""
        bufs = [np.frombuffer(d, dtype=np.uint8) for d in dumps]
        ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        sizes = (C.c_uint64 * len(bufs))(*[b.size for b in bufs])
        h = C.c_void_p()
        _check(lib().shn_index_load_mem(C.byref(h), ptrs, sizes, len(bufs), dim, m, IP if ip else L2, gpu))
        return cls(h, IP if ip else L2)

    @classmethod
    def build(cls, base, m, efc, ip=False, seed=1234, ids=None, gpu=0):
        """GPU construction over host rows base[n][dim] (shn_index_build)."""
        b = np.ascontiguousarray(base, dtype=np.float32)
        i = np.ascontiguousarray(ids, dtype=np.uint32) if ids is not None else None
        h = C.c_void_p()
        _check(lib().shn_index_build(C.byref(h), b.ctypes.data, i.ctypes.data if i is not None else None, b.shape[0],
                                     b.shape[1], m, efc, IP if ip else L2, seed, gpu))
        return cls(h, IP if ip else L2)

    @classmethod
    def build_device(cls, d_base, n, dim, m, efc, ip=False, seed=1234, d_ids=0, gpu=0):
        """GPU construction over rows already in HBM (raw device pointers)."""
        h = C.c_void_p()
        _check(lib().shn_index_build_device(C.byref(h), d_base, d_ids or None, n, dim, m, efc, IP if ip else L2, seed, gpu))
        return cls(h, IP if ip else L2)

    # -- multi-GPU partitioning -----------------------------------------------------------------------------------
    def count_visits(self, enable=True):
        _check(lib().shn_index_count_visits(self._h, int(enable)))

    def visit_counts(self, d_counts, write_back=False):
        """Copy the per-node visit counters to / from a device buffer of n u32 (raw pointer)."""
        _check(lib().shn_index_visit_counts(self._h, d_counts, int(write_back)))

    def partition(self, rank, world, cache_ratio_pct=5, d_owner=0):
        """d_owner: device pointer to n bytes from placement_fit (placement by cluster); 0 = round-robin."""
        h = C.c_void_p()
        _check(lib().shn_index_partition(C.byref(h), self._h, rank, world, cache_ratio_pct, d_owner or None))
        return Index(h, self.metric)

    def placement_fit(self, world, d_owner, seed=1234, slack=0.05):
        """k-means over the upper-level nodes + balanced assignment of every node; returns (centroids [world,dim], sizes)."""
        cent = np.empty((world, self.dim), np.float32)
        sizes = np.zeros(world, np.uint64)
        _check(lib().shn_placement_fit(self._h, world, seed, slack, cent.ctypes.data, d_owner, sizes.ctypes.data))
        return cent, sizes

    def partition_export(self, want_fds=False):
        """(fds or None, sizes, raw pointers): this GPU's share as two POSIX fds (caller closes them) / raw pointers."""
        fds = (C.c_int * 2)(-1, -1)
        sizes = (C.c_uint64 * 2)()
        raw = (C.c_uint64 * 2)()
        _check(lib().shn_index_partition_export(self._h, fds if want_fds else None, sizes, raw))
        return ([int(fds[0]), int(fds[1])] if want_fds else None), [int(sizes[0]), int(sizes[1])], (int(raw[0]), int(raw[1]))

    def partition_attach(self, peer, fds=None, sizes=None, raw_ptrs=None):
        raw = (C.c_uint64 * 2)(*raw_ptrs) if raw_ptrs is not None else None
        cf = (C.c_int * 2)(*fds) if fds is not None else None
        cs = (C.c_uint64 * 2)(*sizes) if sizes is not None else None
        _check(lib().shn_index_partition_attach(self._h, peer, cf, cs, raw))

    def build_stats(self):
        st = Stats()
        _check(lib().shn_index_build_stats(self._h, C.byref(st)))
        return st.as_dict()

    def close(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.shn_index_free(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- introspection -----------------------------------------------------------------------------------------
    @property
    def n(self):
        return lib().shn_index_size(self._h)

    @property
    def dim(self):
        return lib().shn_index_dim(self._h)

    @property
    def m(self):
        return lib().shn_index_m(self._h)

    @property
    def max_level(self):
        return lib().shn_index_max_level(self._h)

    @property
    def hbm_bytes(self):
        return lib().shn_index_hbm_bytes(self._h)

    @property
    def dump_bytes(self):
        return lib().shn_index_dump_bytes(self._h)

    def build_halo(self, ratio_pct):
        """After count_visits(True) + warm-up searches on this partition: cache the most-read peer-owned rows locally."""
        rows = C.c_uint64()
        _check(lib().shn_index_partition_build_halo(self._h, ratio_pct, C.byref(rows)))
        return int(rows.value)

    def partition_info(self):
        """(hot, own, entry_row) of a partition handle."""
        h, o, e = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().shn_index_partition_info(self._h, C.byref(h), C.byref(o), C.byref(e)))
        return int(h.value), int(o.value), int(e.value)

    def set_option(self, key, value):
        _check(lib().shn_set_option(self._h, key.encode(), int(value)))

    # -- store -------------------------------------------------------------------------------------------------
    def store(self, paths):
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        _check(lib().shn_index_store(self._h, arr, len(paths)))

    def to_dumps(self, n_parts=1):
        sizes = (C.c_uint64 * n_parts)()
        _check(lib().shn_index_store_mem(self._h, None, sizes, n_parts))
        bufs = [np.empty(int(s), dtype=np.uint8) for s in sizes]
        ptrs = (C.c_void_p * n_parts)(*[b.ctypes.data for b in bufs])
        _check(lib().shn_index_store_mem(self._h, ptrs, sizes, n_parts))
        return bufs

    # -- search ------------------------------------------------------------------------------------------------
    def search(self, queries, k, ef, out_ids=None, out_dists=None):
        """Host buffers in, host buffers out (numpy).  Returns (ids [nq,k] u32, dists [nq,k] f32, stats dict)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ShnError(-1, f"queries must be [nq, {self.dim}]")
        nq = q.shape[0]
        ids = out_ids if out_ids is not None else np.empty((nq, k), np.uint32)
        dists = out_dists if out_dists is not None else np.empty((nq, k), np.float32)
        st = Stats()
        _check(lib().shn_search(self._h, q.ctypes.data, nq, k, ef, ids.ctypes.data, dists.ctypes.data, C.byref(st)))
        return ids, dists, st.as_dict()

    def search_device(self, d_queries, nq, k, ef, d_ids, d_dists=0, d_counters=0, stream=0, want_stats=True):
        """Raw device pointers (ints).  Synchronises only when want_stats."""
        st = Stats()
        _check(lib().shn_search_device(self._h, d_queries, nq, k, ef, d_ids, d_dists or None, d_counters or None,
                                       stream or None, C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None


class _DevicePointer:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


def device_view(ptr, shape, typestr):
    """A torch tensor aliasing device memory owned by the library (e.g. a router's landing buffers): '<i4', '<f4', '|u1'."""
    import torch
    return torch.as_tensor(_DevicePointer(ptr, shape, typestr), device="cuda")


class Router:
    """Query routing fused with the exchange (include/shn.h shn_router_*): scatter -> barrier -> search -> barrier."""

    def __init__(self, index, centroids, slack=0.25, max_batch=1 << 20, k_max=10):
        c = np.ascontiguousarray(centroids, dtype=np.float32)
        self._h = C.c_void_p()
        self.index, self.world = index, c.shape[0]
        _check(lib().shn_router_create(C.byref(self._h), index._h, c.ctypes.data, slack, max_batch, k_max))

    def export(self, want_fd=False):
        """(fd or None, size, raw device pointer) of this rank's exchange block."""
        fd, size, raw = C.c_int(-1), C.c_uint64(), C.c_uint64()
        _check(lib().shn_router_export(self._h, C.byref(fd) if want_fd else None, C.byref(size), C.byref(raw)))
        return (int(fd.value) if want_fd else None), int(size.value), int(raw.value)

    def attach(self, peer, fd=-1, size=0, raw_ptr=0):
        _check(lib().shn_router_attach(self._h, peer, fd, size, raw_ptr))

    def scatter(self, d_queries, nq, stream=0):
        _check(lib().shn_router_scatter(self._h, d_queries, nq, stream or None))

    def search(self, k, ef, stream=0, want_stats=True):
        st = Stats()
        _check(lib().shn_router_search(self._h, k, ef, stream or None, C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def results(self):
        """Raw device pointers (ids, dists) of this rank's landing buffers."""
        i, d = C.c_void_p(), C.c_void_p()
        _check(lib().shn_router_results(self._h, C.byref(i), C.byref(d)))
        return int(i.value), int(d.value)

    def counts(self, stream=0):
        sent = np.zeros(self.world, np.uint32)
        recv = np.zeros(self.world, np.uint32)
        _check(lib().shn_router_counts(self._h, sent.ctypes.data, recv.ctypes.data, stream or None))
        return sent, recv

    def destinations(self):
        p = C.c_void_p()
        _check(lib().shn_router_destinations(self._h, C.byref(p)))
        return int(p.value)

    def close(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.shn_router_free(self._h)
            self._h = None

    __del__ = close


def repartition_dumps(dumps, dim, m, n_parts_out):
    """Host-only: rewrite reference-format dumps for a different memory-node count."""
    bufs = [np.frombuffer(d, dtype=np.uint8) for d in dumps]
    ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
    sizes = (C.c_uint64 * len(bufs))(*[b.size for b in bufs])
    osz = (C.c_uint64 * n_parts_out)()
    _check(lib().shn_dump_repartition(ptrs, sizes, len(bufs), dim, m, n_parts_out, None, osz))
    out = [np.empty(int(s), dtype=np.uint8) for s in osz]
    optrs = (C.c_void_p * n_parts_out)(*[b.ctypes.data for b in out])
    _check(lib().shn_dump_repartition(ptrs, sizes, len(bufs), dim, m, n_parts_out, optrs, osz))
    return out


def set_build_option(key, value):
    _check(lib().shn_set_build_option(key.encode(), int(value)))


def bruteforce_topk(base, queries, k, ip=False, gpu=0):
    """Exact top-k on the GPU (host arrays in, host arrays out): ids [nq,k] u32 (row numbers), dists [nq,k] f32."""
    b = np.ascontiguousarray(base, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    ids = np.empty((q.shape[0], k), np.uint32)
    dists = np.empty((q.shape[0], k), np.float32)
    _check(lib().shn_bruteforce_topk(b.ctypes.data, b.shape[0], q.ctypes.data, q.shape[0], b.shape[1], IP if ip else L2, k,
                                     ids.ctypes.data, dists.ctypes.data, gpu))
    return ids, dists


def bruteforce_topk_device(d_base, n, d_queries, nq, dim, k, d_ids, d_dists=0, ip=False, gpu=0, stream=0):
    _check(lib().shn_bruteforce_topk_device(d_base, n, d_queries, nq, dim, IP if ip else L2, k, d_ids, d_dists or None, gpu,
                                            stream or None))


def route_queries(centroids, d_queries, nq, ip=False, slack=0.25, gpu=0):
    """dest[q] = rank of the nearest centroid still under its share of the batch (host uint8 array)."""
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    dest = np.empty(nq, np.uint8)
    _check(lib().shn_route_queries(c.ctypes.data, c.shape[0], c.shape[1], IP if ip else L2, d_queries, nq, slack,
                                   dest.ctypes.data, gpu))
    return dest


def draw_levels(n, m, seed=1234):
    """Host-only: node levels as the GPU builder (and a single-coroutine reference build) draws them."""
    out = np.empty(n, np.uint32)
    _check(lib().shn_draw_levels(n, m, seed, out.ctypes.data))
    return out

"""ctypes binding of libvidx_b200.so (include/vidx_b200.h).

The library is hand-written CUDA for sm_100a; there is no CPU path.  Every compute call
raises when no B200 is usable.  Loading the library and resolving its symbols works on
any host (cudart is linked statically), which is what the CPU-side tests check.
"""
import ctypes as C
import os
import re

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# VIDX_B200_LIB: another build of the same library (the role-timer build of tools/README.md), experiments only
LIB_PATH = os.environ.get("VIDX_B200_LIB") or os.path.join(_ROOT, "lib", "libvidx_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_ROOT), "include", "vidx_b200.h")

VIDX_OK, INVALID_INPUT, NOT_FOUND, INVALID_DATA, OTHER, CUDA, UNSUPPORTED = range(7)

u64, u32, i32 = C.c_uint64, C.c_uint32, C.c_int
f32p = C.POINTER(C.c_float)
u64p, u32p, i64p, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
vp = C.c_void_p


class SearchStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_coarse", "ms_select", "ms_group", "ms_scan", "ms_merge", "ms_total")] + [
        (n, u64) for n in ("scan_bytes_algorithmic", "scan_bytes_logical", "scan_flops", "coarse_flops", "n_pairs",
                           "n_dense_items", "n_sparse_items", "kernel_launches")] + [("ms_scan_tc", C.c_double)] + [
        (n, u64) for n in ("n_tc_items", "n_tc_survivors", "n_tc_overflow", "tc_mma_flops", "n_tc_submin_slots")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class VidxError(RuntimeError):
    """Mirrors std::io::Error: .code is one of the VIDX_ERR_* values (= io::ErrorKind)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class InvalidInput(VidxError, ValueError):
    pass


_SIGS = {
    "vidx_create": (i32, [u32, i32, C.POINTER(vp)]),
    "vidx_free": (None, [vp]),
    "vidx_last_error": (C.c_char_p, []),
    "vidx_set_limits": (i32, [vp, u64, u64, u64, u64]),
    "vidx_build": (i32, [vp, f32p, u64p, u64p, u64, u64, u64, u64]),
    "vidx_build_device": (i32, [vp, vp, u64p, u64p, u64, u64, u64, u64]),
    "vidx_build_from_vector_file": (i32, [vp, C.c_char_p, u64, u64, u64]),
    "vidx_vector_file_read": (i32, [C.c_char_p, u64, u64, f32p, u64p, u64p, u64p]),
    "vidx_vector_file_write": (i32, [C.c_char_p, f32p, u64p, u64p, u64, u64, u64]),
    "vidx_train": (i32, [vp, f32p, u64, u64, u64, u64]),
    "vidx_add": (i32, [vp, f32p, u64p, u64p, u64]),
    "vidx_build_from_labels": (i32, [vp, f32p, u64p, u64, f32p, u64, u64p, u64p]),
    "vidx_search": (i32, [vp, f32p, u64, u64, u64, f32p, i64p]),
    "vidx_search_device": (i32, [vp, vp, u64, u64, u64, vp, vp, vp]),
    "vidx_search_with_vectors": (i32, [vp, f32p, u64, u64, u64, f32p, i64p, f32p]),
    "vidx_coarse_probes": (i32, [vp, f32p, u64, u64, u32p, f32p]),
    "vidx_dimension": (u32, [vp]),
    "vidx_ntotal": (u64, [vp]),
    "vidx_nlist": (u64, [vp]),
    "vidx_num_shards": (u64, [vp]),
    "vidx_k_trained": (u64, [vp]),
    "vidx_get_centroids": (i32, [vp, f32p]),
    "vidx_get_centroids_to_shard": (i32, [vp, u64p]),
    "vidx_get_list_sizes": (i32, [vp, u64p]),
    "vidx_get_list_members": (i32, [vp, u64, u64p]),
    "vidx_get_train_labels": (i32, [vp, u64p]),
    "vidx_get_train_centroids": (i32, [vp, f32p]),
    "vidx_resident_vectors": (u64, [vp]),
    "vidx_resident_bytes": (u64, [vp]),
    "vidx_load_warning_count": (u64, [vp]),
    "vidx_load_warning": (C.c_char_p, [vp, u64]),
    "vidx_comm_unique_id": (i32, [C.POINTER(C.c_uint8)]),
    "vidx_comm_init": (i32, [vp, i32, i32, C.POINTER(C.c_uint8)]),
    "vidx_comm_destroy": (i32, [vp]),
    "vidx_comm_version": (C.c_char_p, [vp]),
    "vidx_search_multi": (i32, [vp, f32p, u64, u64, u64, f32p, i64p]),
    "vidx_search_multi_device": (i32, [vp, vp, u64, u64, u64, vp, vp, vp]),
    "vidx_kmeans_mini_batch": (i32, [i32, f32p, u64, u64, u64, u64, C.c_float, u64, f32p, u64p, u64p]),
    "vidx_kmeans_parallel": (i32, [i32, f32p, u64, u64, u64, u64, C.c_float, u64, f32p, u64p, u64p]),
    "vidx_assign_points": (i32, [i32, f32p, u64, u64, f32p, u64, u64, u64p]),
    "vidx_kmeans_pp_init": (i32, [i32, f32p, u64, u64, u64, u64, f32p]),
    "vidx_kmeans_last_profile": (i32, [C.POINTER(C.c_double)]),
    "vidx_calculate_num_clusters": (u64, [u64]),
    "vidx_calculate_max_iterations": (u64, [u64]),
    "vidx_save": (i32, [vp, C.c_char_p, C.c_char_p]),
    "vidx_load": (i32, [vp, C.c_char_p, C.c_char_p]),
    "vidx_stdrng_draw": (i32, [u64, u64, i32, u64, u64, u64p]),
    "vidx_stdrng_weighted": (i32, [u64, f32p, u64, u64, u64p]),
    "vidx_index_bin_write": (i32, [C.c_char_p, f32p, u64p, u64, u32]),
    "vidx_index_bin_read": (i32, [C.c_char_p, u64, f32p, u64p, u64p, u32p]),
    "vidx_set_coarse_mode": (i32, [vp, i32]),
    "vidx_set_partition": (i32, [vp, i32, i32]),
    "vidx_set_partition_mode": (i32, [vp, i32]),
    "vidx_get_partition_kind": (i32, [vp]),
    "vidx_get_shard_owner": (i32, [vp, i32, i32p]),
    "vidx_partition_shards": (i32, [u64p, u64, i32, i32p]),
    "vidx_partition_plan": (i32, [u64p, u64p, u64, u64, i32, i32, i32p, C.POINTER(C.c_int)]),
    "vidx_merge_topk_device": (i32, [i32, vp, vp, u32, u64, u64, vp, vp, vp]),
    "vidx_merge_topk_keyed_device": (i32, [i32, vp, vp, vp, u32, u64, u64, vp, vp, vp]),
    "vidx_grid_plan": (i32, [u64, i32, i32, i32, u64p]),
    "vidx_merge_topk_grid_device": (i32, [i32, vp, vp, vp, u32, u64, u64, u64, vp, vp, vp]),
    "vidx_search_local_device": (i32, [vp, vp, u64, u64, u64, vp, vp, vp, vp]),
    "vidx_set_profiling": (i32, [vp, i32]),
    "vidx_set_scan_mode": (i32, [vp, i32]),
    "vidx_get_search_stats": (i32, [vp, C.POINTER(SearchStats)]),
    "vidx_kernel_launch_count": (u64, []),
}

_lib = None


def header_symbols():
    """Every function name declared in include/vidx_b200.h."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(vidx_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {_ROOT}` (nvcc, sm_100a). "
                              "There is no fallback implementation.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != VIDX_OK:
        msg = lib().vidx_last_error().decode(errors="replace")
        raise (InvalidInput if rc == INVALID_INPUT else VidxError)(rc, msg)


def _f(a):
    return a.ctypes.data_as(f32p)


def _u(a):
    return None if a is None else a.ctypes.data_as(u64p)


def _c32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


class Index:
    """Thin owner of a vidx_index handle (VectorIndexer in src/api.rs)."""

    def __init__(self, dimension, device=0):
        h = vp()
        check(lib().vidx_create(dimension, device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().vidx_free(self.h)
            self.h = None

    __del__ = close

    # ---- build -------------------------------------------------------------------------
    def build(self, data, ext_ids=None, timestamps=None, seed=42, nlist=0, max_iters=0):
        data = _c32(data)
        n = data.shape[0] if data.ndim == 2 else 0
        e = None if ext_ids is None else np.ascontiguousarray(ext_ids, dtype=np.uint64)
        t = None if timestamps is None else np.ascontiguousarray(timestamps, dtype=np.uint64)
        check(lib().vidx_build(self.h, _f(data), _u(e), _u(t), n, seed, nlist, max_iters))
        return self

    def build_device(self, d_data_ptr, n, ext_ids=None, timestamps=None, seed=42, nlist=0, max_iters=0):
        """vidx_build_device: the data set already sits in this GPU's memory (n x dimension fp32)."""
        e = None if ext_ids is None else np.ascontiguousarray(ext_ids, dtype=np.uint64)
        t = None if timestamps is None else np.ascontiguousarray(timestamps, dtype=np.uint64)
        check(lib().vidx_build_device(self.h, d_data_ptr, _u(e), _u(t), n, seed, nlist, max_iters))
        return self

    def build_from_vector_file(self, path, seed=42, nlist=0, max_iters=0):
        """VectorIndexer::build_from_vector_file (src/api.rs:149-186)."""
        check(lib().vidx_build_from_vector_file(self.h, os.fsencode(path), seed, nlist, max_iters))
        return self

    def train(self, data, seed=42, nlist=0, max_iters=0):
        data = _c32(data)
        check(lib().vidx_train(self.h, _f(data), data.shape[0], seed, nlist, max_iters))
        return self

    def add(self, data, ext_ids=None, timestamps=None):
        data = _c32(data)
        e = None if ext_ids is None else np.ascontiguousarray(ext_ids, dtype=np.uint64)
        t = None if timestamps is None else np.ascontiguousarray(timestamps, dtype=np.uint64)
        check(lib().vidx_add(self.h, _f(data), _u(e), _u(t), data.shape[0]))
        return self

    def build_from_labels(self, data, centroids, labels, ext_ids=None, centroid_shard=None):
        data, centroids = _c32(data), _c32(centroids)
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        e = None if ext_ids is None else np.ascontiguousarray(ext_ids, dtype=np.uint64)
        s = None if centroid_shard is None else np.ascontiguousarray(centroid_shard, dtype=np.uint64)
        check(lib().vidx_build_from_labels(self.h, _f(data), _u(e), data.shape[0], _f(centroids), centroids.shape[0],
                                           _u(labels), _u(s)))
        return self

    def set_limits(self, default_k=10, default_n_probe=20, max_k=10_000, max_n_probe=10_000):
        check(lib().vidx_set_limits(self.h, default_k, default_n_probe, max_k, max_n_probe))

    # ---- search ------------------------------------------------------------------------
    def search(self, xq, k, n_probe, include_vectors=False):
        xq = _c32(xq)
        if xq.ndim != 2:
            raise InvalidInput(INVALID_INPUT, "xq must be 2-D")
        nq = xq.shape[0]
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        if include_vectors:
            V = np.empty((nq, k, self.dimension), np.float32)
            check(lib().vidx_search_with_vectors(self.h, _f(xq), nq, k, n_probe, _f(D), I.ctypes.data_as(i64p), _f(V)))
            return D, I, V
        check(lib().vidx_search(self.h, _f(xq), nq, k, n_probe, _f(D), I.ctypes.data_as(i64p)))
        return D, I

    def search_host_ptr(self, xq_ptr, nq, k, n_probe, D_ptr, I_ptr):
        """vidx_search on raw HOST addresses (e.g. pinned torch tensors)."""
        check(lib().vidx_search(self.h, C.cast(xq_ptr, f32p), nq, k, n_probe, C.cast(D_ptr, f32p), C.cast(I_ptr, i64p)))

    def search_device(self, d_xq_ptr, nq, k, n_probe, d_D_ptr, d_I_ptr, stream_ptr=0):
        check(lib().vidx_search_device(self.h, d_xq_ptr, nq, k, n_probe, d_D_ptr, d_I_ptr, stream_ptr))

    # ---- multi-GPU: NCCL communicator + collective search --------------------------------
    def comm_init(self, rank, world, unique_id):
        """unique_id: the 128 bytes rank 0 got from comm_unique_id(), handed to every rank by the host program."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        check(lib().vidx_comm_init(self.h, rank, world, buf))

    def comm_destroy(self):
        check(lib().vidx_comm_destroy(self.h))

    @property
    def comm_version(self):
        return lib().vidx_comm_version(self.h).decode()

    def search_multi(self, xq, k, n_probe):
        xq = _c32(xq)
        nq = xq.shape[0]
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        check(lib().vidx_search_multi(self.h, _f(xq), nq, k, n_probe, _f(D), I.ctypes.data_as(i64p)))
        return D, I

    def search_multi_host_ptr(self, xq_ptr, nq, k, n_probe, D_ptr, I_ptr):
        check(lib().vidx_search_multi(self.h, C.cast(xq_ptr, f32p), nq, k, n_probe, C.cast(D_ptr, f32p), C.cast(I_ptr, i64p)))

    def search_multi_device(self, d_xq_ptr, nq, k, n_probe, d_D_ptr, d_I_ptr, stream_ptr=0):
        check(lib().vidx_search_multi_device(self.h, d_xq_ptr, nq, k, n_probe, d_D_ptr, d_I_ptr, stream_ptr))

    def search_local_device(self, d_xq_ptr, nq, k, n_probe, d_D_ptr, d_I_ptr, d_keys_ptr, stream_ptr=0):
        """The local part of a partitioned search, with (probe rank << 32 | global row) keys for the merge."""
        check(lib().vidx_search_local_device(self.h, d_xq_ptr, nq, k, n_probe, d_D_ptr, d_I_ptr, d_keys_ptr, stream_ptr))

    def coarse_probes(self, xq, n_probe):
        xq = _c32(xq)
        nq = xq.shape[0]
        lists = np.empty((nq, n_probe), np.uint32)
        dists = np.empty((nq, n_probe), np.float32)
        check(lib().vidx_coarse_probes(self.h, _f(xq), nq, n_probe, lists.ctypes.data_as(u32p), _f(dists)))
        return lists, dists

    # ---- introspection -----------------------------------------------------------------
    @property
    def dimension(self):
        return lib().vidx_dimension(self.h)

    @property
    def ntotal(self):
        return lib().vidx_ntotal(self.h)

    @property
    def nlist(self):
        return lib().vidx_nlist(self.h)

    @property
    def num_shards(self):
        return lib().vidx_num_shards(self.h)

    @property
    def k_trained(self):
        return lib().vidx_k_trained(self.h)

    @property
    def resident_vectors(self):
        return lib().vidx_resident_vectors(self.h)

    @property
    def resident_bytes(self):
        return lib().vidx_resident_bytes(self.h)

    def load_warnings(self):
        n = lib().vidx_load_warning_count(self.h)
        return [lib().vidx_load_warning(self.h, i).decode(errors="replace") for i in range(n)]

    def centroids(self):
        out = np.empty((self.nlist, self.dimension), np.float32)
        check(lib().vidx_get_centroids(self.h, _f(out)))
        return out

    def train_centroids(self):
        out = np.empty((self.k_trained, self.dimension), np.float32)
        check(lib().vidx_get_train_centroids(self.h, _f(out)))
        return out

    def train_labels(self):
        out = np.empty(self.ntotal, np.uint64)
        check(lib().vidx_get_train_labels(self.h, _u(out)))
        return out.astype(np.int64)

    def centroids_to_shard(self):
        out = np.empty(self.nlist, np.uint64)
        check(lib().vidx_get_centroids_to_shard(self.h, _u(out)))
        return out.astype(np.int64)

    def list_sizes(self):
        out = np.empty(self.nlist, np.uint64)
        check(lib().vidx_get_list_sizes(self.h, _u(out)))
        return out.astype(np.int64)

    def list_members(self, l):
        out = np.empty(int(self.list_sizes()[l]), np.uint64)
        check(lib().vidx_get_list_members(self.h, l, _u(out)))
        return out.astype(np.int64)

    # ---- persistence / partition / measurement ----------------------------------------
    def save(self, index_dir, shards_dir):
        check(lib().vidx_save(self.h, os.fsencode(index_dir), os.fsencode(shards_dir)))

    def load(self, index_dir, shards_dir):
        check(lib().vidx_load(self.h, os.fsencode(index_dir), os.fsencode(shards_dir)))
        return self

    def set_partition(self, rank, world):
        check(lib().vidx_set_partition(self.h, rank, world))

    def set_partition_mode(self, mode):
        """0 auto, 1 the reference's shards, 2 segment ranges of every list."""
        check(lib().vidx_set_partition_mode(self.h, {"auto": 0, "shards": 1, "ranges": 2}.get(mode, mode)))

    @property
    def partition_kind(self):
        return {1: "shards", 2: "ranges"}.get(lib().vidx_get_partition_kind(self.h))

    def shard_owner(self, world):
        out = np.zeros(self.num_shards, np.int32)
        check(lib().vidx_get_shard_owner(self.h, world, out.ctypes.data_as(i32p)))
        return out

    def set_profiling(self, on=True):
        check(lib().vidx_set_profiling(self.h, 1 if on else 0))

    def set_coarse_mode(self, mode):
        """0 = auto, 1 = exact FP32 coarse stage, 2 = tensor-core filter whenever it applies."""
        check(lib().vidx_set_coarse_mode(self.h, mode))

    def set_scan_mode(self, mode):
        """0 = tensor-core pre-filter + exact re-check (default), 1 = exact kernels only."""
        check(lib().vidx_set_scan_mode(self.h, mode))

    def stats(self):
        s = SearchStats()
        check(lib().vidx_get_search_stats(self.h, C.byref(s)))
        return s.asdict()


def write_vector_file(path, data, ids=None, meta=None, batch=1000):
    """The framing of generate_test_vectors_parallel (src/utils.rs:34-79) with caller-provided values."""
    data = _c32(data)
    i = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
    m = None if meta is None else np.ascontiguousarray(meta, dtype=np.uint64)
    check(lib().vidx_vector_file_write(os.fsencode(path), _f(data), _u(i), _u(m), data.shape[0], data.shape[1], batch))


def read_vector_file(path, dim):
    """read_vectors_from_file (src/utils.rs:82-107) -> (ids, data[n, dim], metadata)."""
    n = C.c_uint64(0)
    check(lib().vidx_vector_file_read(os.fsencode(path), dim, 0, None, None, None, C.byref(n)))
    data = np.zeros((n.value, dim), np.float32)
    ids = np.zeros(n.value, np.uint64)
    meta = np.zeros(n.value, np.uint64)
    check(lib().vidx_vector_file_read(os.fsencode(path), dim, n.value, _f(data), _u(ids), _u(meta), C.byref(n)))
    return ids, data, meta


def stdrng_draw(seed, kind, n, arg=0, skip_u32=0):
    """The build's random stream (csrc/rng.hpp), host only.  kind: "u32", "u64", "gen_range", "shuffle", "choose_multiple",
    "shuffle_head" (first n of a shuffle of 0..arg as the mini-batch loop takes them, plus the next u32 of the stream)."""
    kinds = ["u32", "u64", "gen_range", "shuffle", "choose_multiple", "shuffle_head"]
    extra = 1 if kind == "shuffle_head" else 0
    out = np.zeros(max(n + extra, 1), np.uint64)
    check(lib().vidx_stdrng_draw(seed, skip_u32, kinds.index(kind), arg, n, _u(out)))
    return out[:n + extra]


def stdrng_weighted(seed, weights, n):
    w = _c32(weights)
    out = np.zeros(max(n, 1), np.uint64)
    check(lib().vidx_stdrng_weighted(seed, _f(w), len(w), n, _u(out)))
    return out[:n]


def index_bin_write(index_dir, centroids, centroids_to_shard):
    """index.bin as IvfIndex::save_to writes it (src/ivf_index.rs:274-294); host only."""
    c = _c32(centroids)
    s = np.ascontiguousarray(centroids_to_shard, dtype=np.uint64)
    check(lib().vidx_index_bin_write(os.fsencode(index_dir), _f(c), _u(s), c.shape[0], c.shape[1]))


def index_bin_read(index_dir):
    """-> (centroids f32[nlist, dim], centroids_to_shard i64[nlist]) as load_index_from reads them; host only."""
    n, d = u64(0), u32(0)
    check(lib().vidx_index_bin_read(os.fsencode(index_dir), 0, None, None, C.byref(n), C.byref(d)))
    c = np.zeros((n.value, d.value), np.float32)
    s = np.zeros(n.value, np.uint64)
    check(lib().vidx_index_bin_read(os.fsencode(index_dir), n.value, _f(c), _u(s), C.byref(n), C.byref(d)))
    return c, s.astype(np.int64)


def comm_unique_id():
    """ncclGetUniqueId through the library: 128 bytes for rank 0 to hand to the other ranks."""
    buf = (C.c_uint8 * 128)()
    check(lib().vidx_comm_unique_id(buf))
    return bytes(buf)


def partition_shards(shard_sizes, world):
    sizes = np.ascontiguousarray(shard_sizes, dtype=np.uint64)
    out = np.zeros(len(sizes), np.int32)
    check(lib().vidx_partition_shards(_u(sizes), len(sizes), world, out.ctypes.data_as(i32p)))
    return out


def partition_plan(list_sizes, list_shard, num_shards, world, mode=0):
    """-> (kind "shards" | "ranges", owner rank per shard): the split vidx_set_partition would choose (host only)."""
    sizes = np.ascontiguousarray(list_sizes, dtype=np.uint64)
    shard = np.ascontiguousarray(list_shard, dtype=np.uint64)
    owner = np.zeros(max(int(num_shards), 1), np.int32)
    kind = C.c_int(0)
    check(lib().vidx_partition_plan(_u(sizes), _u(shard), len(sizes), num_shards, world, {"auto": 0, "shards": 1, "ranges": 2}.get(mode, mode),
                                    owner.ctypes.data_as(i32p), C.byref(kind)))
    return {1: "shards", 2: "ranges"}[kind.value], owner[:num_shards]


def grid_plan(nq, world, parts, rank):
    """vidx_search_multi's rank grid: dict(group, q_lo, q_hi, per_group, coarse_lo, coarse_hi) of one rank (host only)."""
    out = np.zeros(6, np.uint64)
    check(lib().vidx_grid_plan(nq, world, parts, rank, _u(out)))
    return dict(zip(("group", "q_lo", "q_hi", "per_group", "coarse_lo", "coarse_hi"), (int(v) for v in out)))


def merge_topk_grid_device(device, d_D_runs, d_I_runs, d_K_runs, parts, per_group, nq, k, d_D, d_I, stream=0):
    check(lib().vidx_merge_topk_grid_device(device, d_D_runs, d_I_runs, d_K_runs, parts, per_group, nq, k, d_D, d_I, stream))


def merge_topk_keyed_device(device, d_D_runs, d_I_runs, d_K_runs, nruns, nq, k, d_D, d_I, stream=0):
    check(lib().vidx_merge_topk_keyed_device(device, d_D_runs, d_I_runs, d_K_runs, nruns, nq, k, d_D, d_I, stream))


def merge_topk_device(device, d_D_runs, d_I_runs, nruns, nq, k, d_D, d_I, stream=0):
    check(lib().vidx_merge_topk_device(device, d_D_runs, d_I_runs, nruns, nq, k, d_D, d_I, stream))


def _kmeans(fn, data, k, max_iters, tol, seed, device):
    data = _c32(data)
    n, d = (data.shape if data.ndim == 2 else (0, 0))
    c = np.zeros((k, max(d, 1)), np.float32)
    labels = np.zeros(max(n, 1), np.uint64)
    it = u64(0)
    check(fn(device, _f(data), n, d, k, max_iters, -1.0 if tol is None else tol, seed, _f(c), _u(labels), C.byref(it)))
    return c, labels[:n].astype(np.int64), it.value


def kmeans_mini_batch(data, k, max_iters, tol=None, seed=42, device=0):
    """run_kmeans_mini_batch (src/kmeans.rs:64)."""
    return _kmeans(lib().vidx_kmeans_mini_batch, data, k, max_iters, tol, seed, device)


def kmeans_parallel(data, k, max_iters, tol=None, seed=42, device=0):
    """run_kmeans_parallel (src/kmeans.rs:15)."""
    return _kmeans(lib().vidx_kmeans_parallel, data, k, max_iters, tol, seed, device)


def assign_points(data, centroids, seed=42, device=0):
    data, centroids = _c32(data), _c32(centroids)
    labels = np.zeros(data.shape[0], np.uint64)
    check(lib().vidx_assign_points(device, _f(data), data.shape[0], data.shape[1], _f(centroids), centroids.shape[0], seed,
                                   _u(labels)))
    return labels.astype(np.int64)


def kmeans_pp_init(data, k, seed=42, device=0):
    data = _c32(data)
    c = np.zeros((k, data.shape[1]), np.float32)
    check(lib().vidx_kmeans_pp_init(device, _f(data), data.shape[0], data.shape[1], k, seed, _f(c)))
    return c


def kmeans_last_profile():
    """(seconds in the host's serial random stream, seconds blocked on the device) of this thread's k-means work since the last call."""
    out = (C.c_double * 2)()
    check(lib().vidx_kmeans_last_profile(out))
    return float(out[0]), float(out[1])


def calculate_num_clusters(n):
    return lib().vidx_calculate_num_clusters(n)


def calculate_max_iterations(n):
    return lib().vidx_calculate_max_iterations(n)


def kernel_launch_count():
    return lib().vidx_kernel_launch_count()

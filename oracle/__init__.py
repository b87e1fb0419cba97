"""ctypes loader for the CPU oracle (oracle/vidx_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(vector-indexer_b200/) never imports this module.

Parity status: "parity unpinned" for RNG-derived quantities -- see the header of
vidx_oracle.cpp and DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvidx_oracle.so")

u64, u32, f32p = C.c_uint64, C.c_uint32, C.POINTER(C.c_float)
u64p, u32p, i64p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++, seconds)."""
    src = os.path.join(_HERE, "vidx_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "MAKEFLAGS": ""})
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.vo_rng_new.restype = C.c_void_p
        L.vo_rng_new.argtypes = [u64]
        L.vo_rng_free.argtypes = [C.c_void_p]
        L.vo_rng_key.argtypes = [C.c_void_p, u32p]
        L.vo_rng_next_u32.restype = u32
        L.vo_rng_next_u32.argtypes = [C.c_void_p]
        L.vo_rng_next_u64.restype = u64
        L.vo_rng_next_u64.argtypes = [C.c_void_p]
        L.vo_rng_gen_range.restype = u64
        L.vo_rng_gen_range.argtypes = [C.c_void_p, u64]
        L.vo_rng_gen_index.restype = u64
        L.vo_rng_gen_index.argtypes = [C.c_void_p, u64]
        L.vo_rng_gen_range_f32.restype = C.c_float
        L.vo_rng_gen_range_f32.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.vo_rng_shuffle.argtypes = [C.c_void_p, u64p, u64]
        L.vo_rng_choose_multiple.restype = u64
        L.vo_rng_choose_multiple.argtypes = [C.c_void_p, u64, u64, u64p]
        L.vo_rng_weighted_index.restype = u64
        L.vo_rng_weighted_index.argtypes = [C.c_void_p, f32p, u64]
        L.vo_chacha_block.argtypes = [u32p, u64, u64, C.c_int, u32p]
        L.vo_create_deterministic_vectors.argtypes = [u64, u64, u64, f32p]
        L.vo_calculate_num_clusters.restype = u64
        L.vo_calculate_num_clusters.argtypes = [u64]
        L.vo_calculate_max_iterations.restype = u64
        L.vo_calculate_max_iterations.argtypes = [u64]
        L.vo_euclidean_distance_squared.restype = C.c_float
        L.vo_euclidean_distance_squared.argtypes = [f32p, f32p, u64]
        L.vo_compute_distance_simd.restype = C.c_float
        L.vo_compute_distance_simd.argtypes = [f32p, f32p, u64]
        L.vo_kmeans_pp_init.argtypes = [f32p, u64, u64, u64, u64, f32p, u64p]
        for fn in (L.vo_kmeans_mini_batch, L.vo_kmeans_parallel):
            fn.restype = C.c_int
            fn.argtypes = [f32p, u64, u64, u64, u64, C.c_float, u64, f32p, u64p, u64p]
        L.vo_assign_points.argtypes = [f32p, u64, u64, f32p, u64, u64, u64p]
        L.vo_assign_brute_force.argtypes = [f32p, u64, u64, f32p, u64, u64p]
        L.vo_build_hierarchy.restype = u64
        L.vo_build_hierarchy.argtypes = [f32p, u64, u64, u64, f32p, u64p]
        L.vo_update_centroids_full.argtypes = [f32p, u64, u64, u64p, u64, f32p, u64p]
        L.vo_centroid_delta.restype = C.c_float
        L.vo_centroid_delta.argtypes = [f32p, f32p, u64, u64]
        L.vo_ivf_fit.restype = C.c_void_p
        L.vo_ivf_fit.argtypes = [f32p, u64p, u64p, u64, u64, u64, u64, u64]
        L.vo_ivf_from_labels.restype = C.c_void_p
        L.vo_ivf_from_labels.argtypes = [f32p, u64p, u64, u64, f32p, u64, u64p]
        L.vo_ivf_free.argtypes = [C.c_void_p]
        L.vo_ivf_set_list_mask.argtypes = [C.c_void_p, C.c_char_p]
        for name in ("nlist", "k_trained", "num_shards", "iters_run"):
            fn = getattr(L, "vo_ivf_" + name)
            fn.restype = u64
            fn.argtypes = [C.c_void_p]
        L.vo_ivf_centroids.argtypes = [C.c_void_p, f32p]
        L.vo_ivf_centroids_all.argtypes = [C.c_void_p, f32p]
        L.vo_ivf_labels_all.argtypes = [C.c_void_p, u64p]
        L.vo_ivf_c2shard.argtypes = [C.c_void_p, u64p]
        L.vo_ivf_list_sizes.argtypes = [C.c_void_p, u64p]
        L.vo_ivf_list_members.argtypes = [C.c_void_p, u64, u64p]
        L.vo_ivf_search.restype = C.c_long
        L.vo_ivf_search.argtypes = [C.c_void_p, f32p, u64, u64, u64p, f32p]
        L.vo_ivf_probes.restype = C.c_long
        L.vo_ivf_probes.argtypes = [C.c_void_p, f32p, u64, u64p, f32p]
        L.vo_ivf_search_batch.restype = C.c_int
        L.vo_ivf_search_batch.argtypes = [C.c_void_p, f32p, u64, u64, u64, f32p, i64p, C.c_int]
        L.vo_brute_force_topk.argtypes = [f32p, u64, u64, f32p, u64, u64, i64p, f32p]
        L.vo_shard_write.restype = C.c_int
        L.vo_shard_write.argtypes = [C.c_char_p, u64, u32, u32, u64p, f32p, u32p, u64p, f32p]
        L.vo_shard_read.restype = C.c_int
        L.vo_shard_read.argtypes = [C.c_char_p, u64, u32p, u32p, u64p, u64p, f32p, u32p, u64p, f32p]
        L.vo_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(f32p)


def _u(a):
    return a.ctypes.data_as(u64p)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Rng:
    """rand 0.8.5 StdRng::seed_from_u64 restatement (ChaCha12)."""

    def __init__(self, seed):
        self.h = lib().vo_rng_new(seed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().vo_rng_free(self.h)
            self.h = None

    def key(self):
        k = np.zeros(8, np.uint32)
        lib().vo_rng_key(self.h, k.ctypes.data_as(u32p))
        return k

    def next_u32(self):
        return lib().vo_rng_next_u32(self.h)

    def next_u64(self):
        return lib().vo_rng_next_u64(self.h)

    def gen_range(self, n):
        return lib().vo_rng_gen_range(self.h, n)

    def gen_index(self, n):
        return lib().vo_rng_gen_index(self.h, n)

    def gen_range_f32(self, lo, hi):
        return lib().vo_rng_gen_range_f32(self.h, lo, hi)

    def shuffle(self, n):
        a = np.arange(n, dtype=np.uint64)
        lib().vo_rng_shuffle(self.h, _u(a), n)
        return a

    def choose_multiple(self, n, amount):
        out = np.zeros(amount, np.uint64)
        m = lib().vo_rng_choose_multiple(self.h, n, amount, _u(out))
        return out[:m]

    def weighted_index(self, w):
        w = _c32(w)
        return lib().vo_rng_weighted_index(self.h, _f(w), len(w))


def chacha_block(key, counter, stream=0, rounds=12):
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(16, np.uint32)
    lib().vo_chacha_block(key.ctypes.data_as(u32p), counter, stream, rounds, out.ctypes.data_as(u32p))
    return out


def create_deterministic_vectors(n, dim, seed):
    out = np.zeros((n, dim), np.float32)
    lib().vo_create_deterministic_vectors(n, dim, seed, _f(out))
    return out


def create_test_vectors(n, dim, scale=0.1, mod=50.0):
    """tests/test_utils/mod.rs:10-16: (x as f32 * 0.1) % 50.0 in f32 arithmetic."""
    x = np.arange(n * dim, dtype=np.float32)
    v = np.fmod(x * np.float32(scale), np.float32(mod)).astype(np.float32)
    return v.reshape(n, dim)


def calculate_num_clusters(n):
    return lib().vo_calculate_num_clusters(n)


def calculate_max_iterations(n):
    return lib().vo_calculate_max_iterations(n)


def euclidean_distance_squared(a, b):
    a, b = _c32(a), _c32(b)
    return lib().vo_euclidean_distance_squared(_f(a), _f(b), len(a))


def compute_distance_simd(a, b):
    a, b = _c32(a), _c32(b)
    return lib().vo_compute_distance_simd(_f(a), _f(b), len(a))


def kmeans_pp_init(data, k, seed):
    data = _c32(data)
    n, d = data.shape
    c = np.zeros((k, d), np.float32)
    chosen = np.zeros(k, np.uint64)
    lib().vo_kmeans_pp_init(_f(data), n, d, k, seed, _f(c), _u(chosen))
    return c, chosen.astype(np.int64)


def _kmeans(fn, data, k, max_iters, tol, seed):
    data = _c32(data)
    if data.size == 0:
        raise ValueError("Input vectors cannot be empty")
    n, d = data.shape
    c = np.zeros((k, d), np.float32)
    labels = np.zeros(n, np.uint64)
    it = u64(0)
    rc = fn(_f(data), n, d, k, max_iters, -1.0 if tol is None else tol, seed, _f(c), _u(labels), C.byref(it))
    if rc != 0:
        raise ValueError("Input vectors cannot be empty")
    return c, labels.astype(np.int64), it.value


def kmeans_mini_batch(data, k, max_iters, tol=None, seed=42):
    return _kmeans(lib().vo_kmeans_mini_batch, data, k, max_iters, tol, seed)


def kmeans_parallel(data, k, max_iters, tol=None, seed=42):
    return _kmeans(lib().vo_kmeans_parallel, data, k, max_iters, tol, seed)


def assign_points(data, cents, seed=42):
    data, cents = _c32(data), _c32(cents)
    labels = np.zeros(len(data), np.uint64)
    lib().vo_assign_points(_f(data), data.shape[0], data.shape[1], _f(cents), len(cents), seed, _u(labels))
    return labels.astype(np.int64)


def assign_brute_force(data, cents):
    data, cents = _c32(data), _c32(cents)
    labels = np.zeros(len(data), np.uint64)
    lib().vo_assign_brute_force(_f(data), data.shape[0], data.shape[1], _f(cents), len(cents), _u(labels))
    return labels.astype(np.int64)


def build_hierarchy(cents, seed=42):
    cents = _c32(cents)
    k, d = cents.shape
    mk = lib().vo_build_hierarchy(_f(cents), k, d, seed, None, None)
    meta = np.zeros((mk, d), np.float32)
    c2m = np.zeros(k, np.uint64)
    lib().vo_build_hierarchy(_f(cents), k, d, seed, _f(meta), _u(c2m))
    return meta, c2m.astype(np.int64)


def update_centroids_full(data, labels, k):
    data = _c32(data)
    labels = np.ascontiguousarray(labels, dtype=np.uint64)
    c = np.zeros((k, data.shape[1]), np.float32)
    cnt = np.zeros(k, np.uint64)
    lib().vo_update_centroids_full(_f(data), data.shape[0], data.shape[1], _u(labels), k, _f(c), _u(cnt))
    return c, cnt.astype(np.int64)


def centroid_delta(a, b):
    a, b = _c32(a), _c32(b)
    return lib().vo_centroid_delta(_f(a), _f(b), a.shape[0], a.shape[1])


class Ivf:
    """In-memory restatement of IvfIndex (src/ivf_index.rs)."""

    def __init__(self, handle, dim):
        self.h = handle
        self.dim = dim

    @classmethod
    def fit(cls, data, ext_ids=None, timestamps=None, seed=42, nlist=0, max_iters=0):
        data = _c32(data)
        if data.size == 0:
            raise ValueError("no vectors provided")
        n, d = data.shape
        e = None if ext_ids is None else np.ascontiguousarray(ext_ids, dtype=np.uint64)
        t = None if timestamps is None else np.ascontiguousarray(timestamps, dtype=np.uint64)
        h = lib().vo_ivf_fit(_f(data), None if e is None else _u(e), None if t is None else _u(t), n, d, seed, nlist,
                             max_iters)
        return cls(h, d)

    @classmethod
    def from_labels(cls, data, cents, labels, ext_ids=None):
        data, cents = _c32(data), _c32(cents)
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        e = None if ext_ids is None else np.ascontiguousarray(ext_ids, dtype=np.uint64)
        h = lib().vo_ivf_from_labels(_f(data), None if e is None else _u(e), data.shape[0], data.shape[1], _f(cents),
                                     len(cents), _u(labels))
        return cls(h, data.shape[1])

    def __del__(self):
        if getattr(self, "h", None):
            lib().vo_ivf_free(self.h)
            self.h = None

    @property
    def nlist(self):
        return lib().vo_ivf_nlist(self.h)

    @property
    def k_trained(self):
        return lib().vo_ivf_k_trained(self.h)

    @property
    def num_shards(self):
        return lib().vo_ivf_num_shards(self.h)

    @property
    def iters_run(self):
        return lib().vo_ivf_iters_run(self.h)

    def centroids(self):
        out = np.zeros((self.nlist, self.dim), np.float32)
        lib().vo_ivf_centroids(self.h, _f(out))
        return out

    def centroids_all(self):
        out = np.zeros((self.k_trained, self.dim), np.float32)
        lib().vo_ivf_centroids_all(self.h, _f(out))
        return out

    def labels_all(self, n):
        out = np.zeros(n, np.uint64)
        lib().vo_ivf_labels_all(self.h, _u(out))
        return out.astype(np.int64)

    def centroids_to_shard(self):
        out = np.zeros(self.nlist, np.uint64)
        lib().vo_ivf_c2shard(self.h, _u(out))
        return out.astype(np.int64)

    def list_sizes(self):
        out = np.zeros(self.nlist, np.uint64)
        lib().vo_ivf_list_sizes(self.h, _u(out))
        return out.astype(np.int64)

    def set_list_mask(self, mask):
        """Scan only lists with mask[l] != 0 (None clears): one rank of the sharded search."""
        lib().vo_ivf_set_list_mask(self.h, None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8).tobytes())

    def list_members(self, l):
        sz = int(self.list_sizes()[l])
        out = np.zeros(sz, np.uint64)
        lib().vo_ivf_list_members(self.h, l, _u(out))
        return out.astype(np.int64)

    def search(self, q, k, nprobe):
        """One query -> (ids, dists); raises ValueError like ErrorKind::InvalidInput."""
        q = _c32(q)
        ids = np.zeros(max(k, 1), np.uint64)
        ds = np.zeros(max(k, 1), np.float32)
        m = lib().vo_ivf_search(self.h, _f(q), k, nprobe, _u(ids), _f(ds))
        if m == -1:
            raise ValueError("k and n_probe must be greater than 0")
        if m == -2:
            raise FloatingPointError("NaN distance (reference panics)")
        return ids[:m].astype(np.int64), ds[:m]

    def probes(self, q, nprobe):
        q = _c32(q)
        ls = np.zeros(nprobe, np.uint64)
        ds = np.zeros(nprobe, np.float32)
        m = lib().vo_ivf_probes(self.h, _f(q), nprobe, _u(ls), _f(ds))
        return ls[:m].astype(np.int64), ds[:m]

    def search_batch(self, xq, k, nprobe, nthreads=1):
        xq = _c32(xq)
        nq = xq.shape[0]
        D = np.zeros((nq, k), np.float32)
        I = np.zeros((nq, k), np.int64)
        rc = lib().vo_ivf_search_batch(self.h, _f(xq), nq, k, nprobe, _f(D), I.ctypes.data_as(i64p), nthreads)
        if rc == -1:
            raise ValueError("k and n_probe must be greater than 0")
        return D, I


def brute_force_topk(data, xq, k, want_d=False):
    data, xq = _c32(data), _c32(xq)
    I = np.zeros((len(xq), k), np.int64)
    D = np.zeros((len(xq), k), np.float32)
    lib().vo_brute_force_topk(_f(data), data.shape[0], data.shape[1], _f(xq), len(xq), k, I.ctypes.data_as(i64p), _f(D))
    return (I, D) if want_d else I


def shard_write(path, shard_id, dim, centroid_ids, centroid_vecs, lens, meta, vecs):
    cid = np.ascontiguousarray(centroid_ids, dtype=np.uint64)
    cv = _c32(centroid_vecs)
    ln = np.ascontiguousarray(lens, dtype=np.uint32)
    mt = np.ascontiguousarray(meta, dtype=np.uint64)
    vv = _c32(vecs)
    return lib().vo_shard_write(path.encode(), shard_id, dim, len(cid), _u(cid), _f(cv), ln.ctypes.data_as(u32p),
                                _u(mt), _f(vv))


def shard_read(path, shard_id):
    d, nl, tot = u32(0), u32(0), u64(0)
    rc = lib().vo_shard_read(path.encode(), shard_id, C.byref(d), C.byref(nl), C.byref(tot), None, None, None, None,
                             None)
    if rc:
        raise IOError(rc)
    cid = np.zeros(nl.value, np.uint64)
    cv = np.zeros((nl.value, d.value), np.float32)
    ln = np.zeros(nl.value, np.uint32)
    mt = np.zeros((tot.value, 3), np.uint64)
    vv = np.zeros((tot.value, d.value), np.float32)
    rc = lib().vo_shard_read(path.encode(), shard_id, C.byref(d), C.byref(nl), C.byref(tot), _u(cid), _f(cv),
                             ln.ctypes.data_as(u32p), _u(mt), _f(vv))
    if rc:
        raise IOError(rc)
    return dict(dim=d.value, centroid_ids=cid, centroid_vecs=cv, lens=ln, meta=mt, vecs=vv)


def num_threads():
    return lib().vo_num_threads()

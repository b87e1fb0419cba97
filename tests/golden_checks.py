"""Known-answer values: the SAME quantities tools/golden/src/main.rs prints from the real crates, computed here from the
oracle (CPU) and -- where a GPU is needed -- from the product.  tests/test_golden.py compares the two sides whenever
tests/golden/reference_golden.json exists (it is produced by `cargo run` on a machine with a Rust toolchain; this image
has none).  Keys and layouts mirror main.rs one to one."""
import numpy as np

X8 = np.array([16.5, 2176.0, 0.203125, 6.75, 54.0, 1.8125, 19.5, 2944.0], np.float32)
X4 = np.array([1.125, 5248.0, 2.6875, 1.0], np.float32)
KMEANS_CASES = [(5000, 32, 20, 50), (5000, 32, 150, 30), (300, 13, 7, 25)]
U64_MAX = 2 ** 64 - 1


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32).ravel().tolist()


def tiny_index_records():
    ids = (1000 + np.arange(12)).astype(np.uint64)
    vals = (np.arange(12, dtype=np.float32)[:, None] * np.float32(0.37) + np.arange(3, dtype=np.float32)[None, :]).astype(np.float32)
    ts = (1_700_000_000 + np.arange(12)).astype(np.uint64)
    return ids, vals, ts


def reduce_trees(x):
    """What each candidate wide::f32x8::reduce_add tree gives for the squares of x (for the failure message)."""
    f = np.float32
    l = [f(v) * f(v) for v in x]
    seq4 = lambda a: f(f(f(a[0] + a[1]) + a[2]) + a[3])
    pair4 = lambda a: f(f(a[0] + a[1]) + f(a[2] + a[3]))
    stride4 = lambda a: f(f(a[0] + a[2]) + f(a[1] + a[3]))
    if len(l) == 4:
        return {"sequential": bits([seq4(l)])[0], "pairwise": bits([pair4(l)])[0], "strided": bits([stride4(l)])[0]}
    return {"halves_sequential": bits([f(seq4(l[:4]) + seq4(l[4:]))])[0], "halves_pairwise": bits([f(pair4(l[:4]) + pair4(l[4:]))])[0],
            "halves_strided": bits([f(stride4(l[:4]) + stride4(l[4:]))])[0],
            "avx": bits([f(f(f(l[0] + l[4]) + f(l[2] + l[6])) + f(f(l[1] + l[5]) + f(l[3] + l[7])))])[0]}


def oracle_side(O, with_kmeans=True):
    g = {}
    r = O.Rng(42)
    g["rng_u32_seed42"] = [r.next_u32() for _ in range(130)]
    r = O.Rng(42)
    g["rng_u64_seed42"] = [str(r.next_u64()) for _ in range(40)]
    r = O.Rng(7)
    for _ in range(63):
        r.next_u32()
    g["rng_u64_straddle_seed7"] = [str(r.next_u64()) for _ in range(3)]
    for seed in (0, 1309, 756, U64_MAX):
        r = O.Rng(seed)
        g[f"rng_u32_seed{seed}"] = [r.next_u32() for _ in range(8)]
    g["shuffle_10_seed42"] = O.Rng(42).shuffle(10).tolist()
    g["shuffle_1000_seed7"] = O.Rng(7).shuffle(1000).tolist()
    gr = {}
    for n in (10, 1000, 50_000, 1_000_000, 3_000_000_000):
        r = O.Rng(42)
        gr[str(n)] = [r.gen_range(n) for _ in range(20)]
    g["gen_range_usize_seed42"] = gr
    w = np.array([((i * 37) % 11 + 1) * 0.25 for i in range(16)], np.float32)
    r = O.Rng(42)
    g["weighted_index_16_seed42"] = [r.weighted_index(w) for _ in range(20)]
    x = np.array([float((i * 7919) % 1013) for i in range(1000)], np.float32)
    w2 = (((x * x).astype(np.float32) * x).astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    r = O.Rng(756)
    g["weighted_index_1000_seed756"] = [r.weighted_index(w2) for _ in range(20)]
    g["choose_multiple_50_7_seed756"] = O.Rng(756).choose_multiple(50, 7).tolist()
    g["choose_multiple_4096_64_seed756"] = O.Rng(756).choose_multiple(4096, 64).tolist()
    g["gen_range_f32_m10_10_seed7"] = bits(O.create_deterministic_vectors(4, 8, 7))
    # wide: compute_distance_simd(a, 0) = reduce_add(a^2 per lane) + 0 + 0 (kmeans.rs:377-419)
    g["wide_f32x8_reduce_add"] = bits([O.compute_distance_simd(X8, np.zeros(8, np.float32))])[0]
    a12 = np.concatenate([np.zeros(8, np.float32), X4])
    g["wide_f32x4_reduce_add"] = bits([O.compute_distance_simd(a12, np.zeros(12, np.float32))])[0]
    g["calculate_num_clusters"] = [[n, O.calculate_num_clusters(n), O.calculate_max_iterations(n)]
                                   for n in (3, 5000, 9999, 10_000, 50_000, 99_999, 100_000, 1_000_000, 10_000_000)]
    a = np.array([((i * 37) % 101) * 0.125 - 3.0 for i in range(37)], np.float32)
    b = np.array([((i * 53) % 97) * 0.0625 for i in range(37)], np.float32)
    g["euclidean_distance_squared_37"] = bits([O.euclidean_distance_squared(a, b)])[0]
    if with_kmeans:
        for n, dim, k, iters in KMEANS_CASES:
            c, l, _ = O.kmeans_mini_batch(O.create_test_vectors(n, dim), k, iters, seed=42)
            g[f"kmeans_mini_batch_ramp_{n}x{dim}_k{k}_it{iters}"] = {"centroids": bits(c), "labels": l.tolist()}
        c, l, _ = O.kmeans_parallel(O.create_test_vectors(600, 8), 5, 10, seed=42)
        g["kmeans_parallel_ramp_600x8_k5_it10"] = {"centroids": bits(c), "labels": l.tolist()}
    return g


def product_rng_side(ffi):
    """The same RNG keys from the product's own stream (csrc/rng.hpp through vidx_stdrng_*; host only)."""
    g = {}
    g["rng_u32_seed42"] = ffi.stdrng_draw(42, "u32", 130).tolist()
    g["rng_u64_seed42"] = [str(v) for v in ffi.stdrng_draw(42, "u64", 40).tolist()]
    g["rng_u64_straddle_seed7"] = [str(v) for v in ffi.stdrng_draw(7, "u64", 3, skip_u32=63).tolist()]
    for seed in (0, 1309, 756, U64_MAX):
        g[f"rng_u32_seed{seed}"] = ffi.stdrng_draw(seed, "u32", 8).tolist()
    g["shuffle_10_seed42"] = ffi.stdrng_draw(42, "shuffle", 10, arg=10).tolist()
    g["shuffle_1000_seed7"] = ffi.stdrng_draw(7, "shuffle", 1000, arg=1000).tolist()
    g["gen_range_usize_seed42"] = {str(n): ffi.stdrng_draw(42, "gen_range", 20, arg=n).tolist()
                                   for n in (10, 1000, 50_000, 1_000_000, 3_000_000_000)}
    w = np.array([((i * 37) % 11 + 1) * 0.25 for i in range(16)], np.float32)
    g["weighted_index_16_seed42"] = ffi.stdrng_weighted(42, w, 20).tolist()
    x = np.array([float((i * 7919) % 1013) for i in range(1000)], np.float32)
    w2 = (((x * x).astype(np.float32) * x).astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    g["weighted_index_1000_seed756"] = ffi.stdrng_weighted(756, w2, 20).tolist()
    g["choose_multiple_50_7_seed756"] = ffi.stdrng_draw(756, "choose_multiple", 7, arg=50).tolist()
    g["choose_multiple_4096_64_seed756"] = ffi.stdrng_draw(756, "choose_multiple", 64, arg=4096).tolist()
    return g


def compare(golden, mine, keys=None):
    """-> list of (key, detail) for every key of `mine` that the golden file has and that differs."""
    bad = []
    for k in (keys or mine):
        if k not in golden or k not in mine:
            continue
        if golden[k] != mine[k]:
            bad.append((k, f"golden {str(golden[k])[:120]} != ours {str(mine[k])[:120]}"))
    return bad

"""GPU parity: k-means through the C ABI vs the CPU oracle.  Centroids must be
bit-identical and labels equal: both sides run the reference's arithmetic
(src/kmeans.rs) and draw from the same (independently implemented) StdRng stream."""
import numpy as np
import pytest

from conftest import bench_data

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def verify_optimal_assignment(data, cents, labels):
    """tests/test_utils/mod.rs:125-144 (Euclidean, eps 1e-5)."""
    d = np.sqrt(((data[:, None, :] - cents[None, :, :]) ** 2).sum(-1, dtype=np.float32))
    assigned = d[np.arange(len(data)), labels]
    return bool((d.min(1) >= assigned - 1e-5).all())


@pytest.mark.parametrize("n,d,k", [(3000, 32, 20), (2000, 10, 7), (1500, 13, 5), (800, 3, 4), (500, 128, 50)])
def test_pp_init_matches_oracle(oracle, ffi, n, d, k):
    xb, _ = bench_data(n, d, 1, seed=n)
    co, _ = oracle.kmeans_pp_init(xb, k, 42)
    cg = ffi.kmeans_pp_init(xb, k, 42)
    assert same_bits(co, cg)


def test_pp_init_sampled_path(oracle, ffi):
    # n > 50 000 -> sampled variant incl. its row-indexing quirk (kmeans.rs:268, :431-436)
    xb, _ = bench_data(60000, 16, 1)
    co, _ = oracle.kmeans_pp_init(xb, 40, 42)
    cg = ffi.kmeans_pp_init(xb, 40, 42)
    assert same_bits(co, cg)


@pytest.mark.parametrize("n,d,k", [(4000, 32, 16), (3000, 10, 100), (2500, 13, 9), (1200, 129, 30)])
def test_assign_brute_matches_oracle(oracle, ffi, n, d, k):
    xb, _ = bench_data(n, d, 1, seed=k)
    cents = xb[:k].copy()
    assert np.array_equal(oracle.assign_points(xb, cents, 42), ffi.assign_points(xb, cents, 42))


@pytest.mark.parametrize("n,d,k", [(6000, 32, 150), (5000, 16, 400), (3000, 20, 101)])
def test_assign_hierarchical_matches_oracle(oracle, ffi, n, d, k):
    # k > 100 -> centroid hierarchy, top-3 metas, argmin in candidate order (kmeans.rs:474-581)
    xb, _ = bench_data(n, d, 1, seed=k)
    cents = xb[np.random.default_rng(1).choice(n, k, replace=False)].copy()
    lo = oracle.assign_points(xb, cents, 42)
    lg = ffi.assign_points(xb, cents, 42)
    assert np.array_equal(lo, lg), f"{(lo != lg).sum()} label mismatches"


def test_mini_batch_small_k(oracle, ffi):
    data = oracle.create_test_vectors(5000, 32)  # tests/kmeans_tests.rs:481-492
    co, lo, io = oracle.kmeans_mini_batch(data, 5, 10, None, 42)
    cg, lg, ig = ffi.kmeans_mini_batch(data, 5, 10, None, 42)
    assert io == ig and same_bits(co, cg) and np.array_equal(lo, lg)
    assert verify_optimal_assignment(data, cg, lg)


def test_mini_batch_hierarchical_k150(oracle, ffi):
    data = oracle.create_test_vectors(5000, 32)  # tests/kmeans_tests.rs:628-649
    co, lo, io = oracle.kmeans_mini_batch(data, 150, 20, None, 42)
    cg, lg, ig = ffi.kmeans_mini_batch(data, 150, 20, None, 42)
    assert io == ig and same_bits(co, cg) and np.array_equal(lo, lg)
    assert verify_optimal_assignment(data, cg, lg)


def test_mini_batch_gaussian(oracle, ffi):
    xb, _ = bench_data(20000, 64, 1)
    co, lo, io = oracle.kmeans_mini_batch(xb, 282, 30, None, 42)
    cg, lg, ig = ffi.kmeans_mini_batch(xb, 282, 30, None, 42)
    assert io == ig and same_bits(co, cg)
    assert np.array_equal(lo, lg), f"{(lo != lg).sum()} label mismatches"


def test_lloyd_matches_oracle(oracle, ffi):
    data = oracle.create_test_vectors(1000, 16)  # tests/kmeans_tests.rs:38-49
    co, lo, io = oracle.kmeans_parallel(data, 5, 30, None, 42)
    cg, lg, ig = ffi.kmeans_parallel(data, 5, 30, None, 42)
    assert io == ig and same_bits(co, cg) and np.array_equal(lo, lg)
    xb, _ = bench_data(4000, 24, 1)
    co, lo, io = oracle.kmeans_parallel(xb, 120, 6, None, 7)  # hierarchical assignment inside Lloyd
    cg, lg, ig = ffi.kmeans_parallel(xb, 120, 6, None, 7)
    assert io == ig and same_bits(co, cg) and np.array_equal(lo, lg)


def test_single_cluster_is_mean(ffi):
    # tests/kmeans_tests.rs:56-78, :596-621
    xb, _ = bench_data(500, 8, 1)
    c, l, _ = ffi.kmeans_parallel(xb, 1, 20, None, 42)
    assert (l == 0).all() and np.allclose(c[0], xb.mean(0), atol=1e-4)


def test_k_equals_n_and_k_larger_than_n(oracle, ffi):
    # tests/kmeans_tests.rs:81-95, :744-773
    xb, _ = bench_data(10, 4, 1)
    for k in (10, 15):
        co, lo, _ = oracle.kmeans_mini_batch(xb, k, 5, None, 42)
        cg, lg, _ = ffi.kmeans_mini_batch(xb, k, 5, None, 42)
        assert same_bits(co, cg) and np.array_equal(lo, lg)
        assert lg.min() >= 0 and lg.max() < k


def test_identical_points_share_a_label(ffi):
    # tests/kmeans_tests.rs:118-144
    xb = np.tile(np.array([[1.0, 2.0, 3.0, 4.0]], np.float32), (50, 1))
    c, l, _ = ffi.kmeans_mini_batch(xb, 3, 10, None, 42)
    assert len(set(l.tolist())) == 1


def test_empty_input_is_invalid_input(ffi):
    # tests/kmeans_tests.rs:735-741
    with pytest.raises(ffi.InvalidInput):
        ffi.kmeans_mini_batch(np.zeros((0, 4), np.float32), 3, 10)
    with pytest.raises(ffi.InvalidInput):
        ffi.kmeans_parallel(np.zeros((0, 4), np.float32), 3, 10)

"""The reference's k-means test properties (tests/kmeans_tests.rs) re-run against the CPU
oracle, plus the arithmetic details of src/kmeans.rs the oracle must reproduce."""
import numpy as np
import pytest


def verify_optimal_assignment(data, cents, labels):
    """tests/test_utils/mod.rs:125-144"""
    d = np.sqrt(((data[:, None, :] - cents[None, :, :]) ** 2).sum(-1, dtype=np.float32))
    return bool((d.min(1) >= d[np.arange(len(data)), labels] - 1e-5).all())


def inertia(data, cents, labels):
    return float(((data - cents[labels]) ** 2).sum())


def test_ramp_fixture(oracle):
    # tests/test_utils/mod.rs:10-16: (x as f32 * 0.1) % 50.0
    v = oracle.create_test_vectors(5, 4)
    assert v.dtype == np.float32 and v.shape == (5, 4)
    x = np.float32(7) * np.float32(0.1)
    assert v.ravel()[7] == np.fmod(x, np.float32(50.0))
    assert oracle.create_test_vectors(1000, 32).max() < 50.0


def test_heuristics(oracle):
    # utils.rs:9-26
    assert [oracle.calculate_num_clusters(n) for n in (100, 5000, 9999, 10000, 50000, 99999, 100000, 10 ** 6)] == \
        [10, 70, 99, 200, 448, 634, 1268, 4000]
    assert [oracle.calculate_max_iterations(n) for n in (1, 9999, 10000, 99999, 100000, 999999, 10 ** 6)] == \
        [300, 300, 100, 100, 50, 50, 20]


def test_distance_sum_orders(oracle):
    """utils.rs:28-30 is a sequential fold; kmeans.rs:377-419 is 8 strided lanes, then one
    4-wide chunk, then the tail, combined left to right.  Checked against explicit float32
    replays."""
    rng = np.random.default_rng(1)
    for d in (1, 3, 4, 7, 8, 12, 13, 31, 64, 100, 128):
        a = rng.standard_normal(d).astype(np.float32) * 3
        b = rng.standard_normal(d).astype(np.float32) * 3
        e = ((a - b) * (a - b)).astype(np.float32)
        s = np.float32(0)
        for x in e:
            s = np.float32(s + x)
        assert oracle.euclidean_distance_squared(a, b) == s
        lanes = np.zeros(8, np.float32)
        j = 0
        while j + 8 <= d:
            lanes = (lanes + e[j:j + 8]).astype(np.float32)
            j += 8
        a4 = np.zeros(4, np.float32)
        if j + 4 <= d:
            a4 = e[j:j + 4].copy()
            j += 4
        tail = np.float32(0)
        for x in e[j:]:
            tail = np.float32(tail + x)
        f = np.float32
        lo = f(f(f(lanes[0] + lanes[1]) + lanes[2]) + lanes[3])
        hi = f(f(f(lanes[4] + lanes[5]) + lanes[6]) + lanes[7])
        r4 = f(f(f(a4[0] + a4[1]) + a4[2]) + a4[3])
        assert oracle.compute_distance_simd(a, b) == f(f(f(lo + hi) + r4) + tail)


def test_basic_shapes_and_label_range(oracle):
    # tests/kmeans_tests.rs:12-34
    data = oracle.create_test_vectors(1000, 10)
    c, l, _ = oracle.kmeans_parallel(data, 5, 50)
    assert c.shape == (5, 10) and l.shape == (1000,) and l.min() >= 0 and l.max() < 5


def test_full_batch_labels_are_optimal(oracle):
    # tests/kmeans_tests.rs:38-49
    data = oracle.create_test_vectors(500, 8)
    c, l, _ = oracle.kmeans_parallel(data, 4, 100)
    assert verify_optimal_assignment(data, c, l)


def test_single_cluster_is_the_mean(oracle):
    # tests/kmeans_tests.rs:56-78, :596-621
    data = oracle.create_test_vectors(100, 5)
    c, l, _ = oracle.kmeans_parallel(data, 1, 50)
    assert (l == 0).all() and np.allclose(c[0], data.mean(0), atol=1e-3)


def test_k_equals_n_and_k_larger_than_n(oracle):
    # tests/kmeans_tests.rs:81-95, :744-773
    data = oracle.create_test_vectors(10, 3)
    for k in (10, 15):
        c, l, _ = oracle.kmeans_mini_batch(data, k, 10)
        assert c.shape == (k, 3) and l.min() >= 0 and l.max() < k


def test_high_dimensional(oracle):
    # tests/kmeans_tests.rs:98-115 (dim 1536)
    data = oracle.create_test_vectors(100, 1536, 0.01, 100.0)
    c, l, _ = oracle.kmeans_parallel(data, 3, 10)
    assert c.shape == (3, 1536)


def test_identical_points(oracle):
    # tests/kmeans_tests.rs:118-144
    data = np.tile(np.array([[1, 2, 3]], np.float32), (20, 1))
    c, l, _ = oracle.kmeans_parallel(data, 3, 10)
    assert len(set(l.tolist())) == 1


def test_mini_batch_optimal_assignment(oracle):
    # tests/kmeans_tests.rs:481-492
    data = oracle.create_test_vectors(500, 8)
    c, l, _ = oracle.kmeans_mini_batch(data, 4, 100)
    assert verify_optimal_assignment(data, c, l)


@pytest.mark.parametrize("k", [150, 200])
def test_hierarchical_assignment_is_optimal_on_ramp(oracle, k):
    # tests/kmeans_tests.rs:628-649, :652-698: n=5000, dim=32, k > 100 -> hierarchical
    data = oracle.create_test_vectors(5000, 32)
    c, l, _ = oracle.kmeans_mini_batch(data, k, 20)
    assert c.shape == (k, 32) and verify_optimal_assignment(data, c, l)


def test_well_separated_clusters_recovered(oracle):
    # tests/kmeans_tests.rs:330-373 (generator: tests/test_utils/mod.rs:34-66, rng unseeded there)
    rng = np.random.default_rng(0)
    k, per, dim = 3, 100, 4
    centers = np.array([[c * 10.0 + d * 0.1 for d in range(dim)] for c in range(k)], np.float32)
    data = np.concatenate([centers[c] + rng.uniform(-0.5, 0.5, (per, dim)).astype(np.float32) for c in range(k)])
    truth = np.repeat(np.arange(k), per)
    for fn in (oracle.kmeans_parallel, oracle.kmeans_mini_batch):
        _, l, _ = fn(data, k, 100)
        # every true cluster maps to exactly one label
        assert all(len(set(l[truth == c].tolist())) == 1 for c in range(k))
        assert len(set(l.tolist())) == k


def test_mini_batch_inertia_within_factor_of_full_batch(oracle):
    # tests/kmeans_tests.rs:541-579 (< 1.5x)
    data = oracle.create_test_vectors(2000, 16)
    cf, lf, _ = oracle.kmeans_parallel(data, 8, 100)
    cm, lm, _ = oracle.kmeans_mini_batch(data, 8, 100)
    lf = oracle.assign_brute_force(data, cf)
    assert inertia(data, cm, lm) < 1.5 * inertia(data, cf, lf)


def test_deterministic_runs(oracle):
    # tests/kmeans_tests.rs:201-323
    data = oracle.create_test_vectors(800, 12)
    a = oracle.kmeans_mini_batch(data, 10, 50, None, 42)
    b = oracle.kmeans_mini_batch(data, 10, 50, None, 42)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = oracle.kmeans_mini_batch(data, 10, 50, None, 43)
    assert not np.array_equal(a[0], c[0])


def test_empty_input_is_an_error(oracle):
    # tests/kmeans_tests.rs:735-741
    with pytest.raises(ValueError):
        oracle.kmeans_parallel(np.zeros((0, 4), np.float32), 3, 10)
    with pytest.raises(ValueError):
        oracle.kmeans_mini_batch(np.zeros((0, 4), np.float32), 3, 10)


def test_kmeans_pp_picks_distinct_rows_and_first_is_gen_range(oracle):
    # kmeans.rs:167-228
    data = oracle.create_deterministic_vectors(300, 6, 11)
    c, chosen = oracle.kmeans_pp_init(data, 12, 42)
    assert chosen[0] == oracle.Rng(42).gen_range(300)
    assert len(set(chosen.tolist())) == 12
    assert np.array_equal(c, data[chosen])


def test_mini_batch_counts_batches_not_points(oracle):
    """kmeans.rs:757: per-cluster count +1 per touching batch => with k=1 the learning rate
    after t iterations is 1/t; replay the update by hand in float32."""
    data = oracle.create_deterministic_vectors(64, 3, 5)
    n, k = 64, 1
    iters = 3
    c, _, ran = oracle.kmeans_mini_batch(data, k, iters, 0.0, 42)
    assert ran == iters
    # replay: rng stream = seed 42: k-means++ consumes its own stream; main loop shuffles
    cent = data[oracle.Rng(42).gen_range(n)].copy()
    r = oracle.Rng(42)
    f = np.float32
    for t in range(1, iters + 1):
        idx = r.shuffle(n)[:10]  # batch = min(256, max(10, floor(sqrt(64)) = 8)) = 10
        s = np.zeros(3, np.float32)
        for i in idx:
            s = (s + data[i]).astype(np.float32)
        mean = (s / f(len(idx))).astype(np.float32)
        eta = f(1.0) / f(t)
        cent = ((f(1.0) - eta) * cent).astype(np.float32) + (eta * mean).astype(np.float32)
        cent = cent.astype(np.float32)
    assert np.array_equal(c[0], cent)


def test_hierarchy_shape(oracle):
    # kmeans.rs:483: meta_k = clamp(floor(sqrt(k)), 2, k/2); 5 Lloyd iterations over the centroids
    cents = oracle.create_deterministic_vectors(150, 8, 3)
    meta, c2m = oracle.build_hierarchy(cents, 42)
    assert meta.shape == (12, 8) and c2m.min() >= 0 and c2m.max() < 12
    # hierarchical labels are the argmin over the top-3 metas' centroids, in candidate order
    data = oracle.create_deterministic_vectors(400, 8, 4)
    labels = oracle.assign_points(data, cents, 42)
    for i in range(0, 400, 37):
        dm = np.array([oracle.compute_distance_simd(data[i], m) for m in meta], np.float32)
        top = np.argsort(dm, kind="stable")[:3]
        cand = [c for m in top for c in np.nonzero(c2m == m)[0]]
        dc = np.array([oracle.compute_distance_simd(data[i], cents[c]) for c in cand], np.float32)
        assert labels[i] == cand[int(np.argmin(dc))]

"""The reference's IVF / integration test properties (tests/ivf_index_tests.rs,
tests/integration_tests.rs, tests/api_tests.rs) re-run against the CPU oracle."""
import numpy as np
import pytest

from conftest import bench_data


@pytest.fixture(scope="module")
def small(oracle):
    data = oracle.create_test_vectors(500, 16)
    return data, oracle.Ivf.fit(data, seed=42)


def test_fit_shapes(oracle, small):
    data, ix = small
    # n < 10 000: k = floor(sqrt(500)) = 22; num_shards = ceil(sqrt(22)) = 5 (ivf_index.rs:59, :104)
    assert ix.k_trained == 22 and ix.num_shards == 5
    assert 1 <= ix.nlist <= 22
    sizes = ix.list_sizes()
    assert sizes.sum() == 500 and (sizes > 0).all()  # empty lists filtered (ivf_index.rs:122-126)
    assert ix.centroids_to_shard().max() < 5


def test_every_vector_in_exactly_one_list_ascending(oracle, small):
    # tests/ivf_index_tests.rs:550-653; ivf_index.rs:94-101 pushes in index order
    _, ix = small
    seen = []
    for l in range(ix.nlist):
        m = ix.list_members(l)
        assert (np.diff(m) > 0).all()
        seen.append(m)
    seen = np.concatenate(seen)
    assert np.array_equal(np.sort(seen), np.arange(500))


def test_self_query_is_first_with_zero_distance(oracle, small):
    # tests/ivf_index_tests.rs:122-159 (distance < 0.1, first result)
    data, ix = small
    for i in (0, 7, 123, 499):
        ids, d = ix.search(data[i], 5, ix.nlist)
        assert d[0] == 0.0 and np.array_equal(data[ids[0]], data[i])


def test_exactly_k_results_sorted(oracle, small):
    # tests/ivf_index_tests.rs:163-224
    data, ix = small
    ids, d = ix.search(data[3] + 0.01, 10, 5)
    assert len(ids) == 10 and (np.diff(d) >= 0).all() and (d >= 0).all()


def test_k_larger_than_n_returns_all(oracle):
    # tests/ivf_index_tests.rs:278-306
    data = oracle.create_test_vectors(50, 8)
    ix = oracle.Ivf.fit(data)
    ids, _ = ix.search(data[0], 100, 1000)
    assert len(ids) == 50 and len(set(ids.tolist())) == 50


def test_single_vector_index(oracle):
    # tests/ivf_index_tests.rs:369-392
    data = np.array([[1, 2, 3, 4]], np.float32)
    ix = oracle.Ivf.fit(data)
    ids, d = ix.search(data[0], 3, 2)
    assert ids.tolist() == [0] and d[0] == 0


def test_zero_k_or_nprobe_invalid(oracle, small):
    # tests/ivf_index_tests.rs:396-457
    data, ix = small
    with pytest.raises(ValueError):
        ix.search(data[0], 0, 3)
    with pytest.raises(ValueError):
        ix.search(data[0], 3, 0)


def test_recall_monotone_in_nprobe(oracle):
    # tests/integration_tests.rs:310-391
    data = oracle.create_test_vectors(2000, 32)
    ix = oracle.Ivf.fit(data)
    rng = np.random.default_rng(3)
    xq = data[rng.choice(2000, 60, replace=False)] + rng.standard_normal((60, 32)).astype(np.float32) * 0.01
    gt = oracle.brute_force_topk(data, xq, 10)

    def recall(nprobe):
        _, I = ix.search_batch(xq, 10, nprobe, nthreads=0)
        return np.mean([len(set(I[i]) & set(gt[i])) / 10 for i in range(len(xq))])

    r5, r15, rall = recall(5), recall(15), recall(ix.nlist)
    assert rall == 1.0 and r15 >= r5 - 1e-9 and r15 >= 0.7


def test_external_ids_are_returned(oracle):
    # tests/api_tests.rs:40-92 (self-query returns external_id 42); ivf_index.rs:258
    data = oracle.create_test_vectors(100, 8)
    ext = np.arange(100, dtype=np.uint64) + 42
    ix = oracle.Ivf.fit(data, ext_ids=ext)
    ids, _ = ix.search(data[0], 1, ix.nlist)
    assert ids[0] == 42


def test_batch_binding_padding(oracle):
    # bindings/python/src/lib.rs:179-187: D init +inf, I init -1
    data = oracle.create_test_vectors(30, 4)
    ix = oracle.Ivf.fit(data)
    D, I = ix.search_batch(data[:3], 40, 100, nthreads=1)
    assert (I[:, :30] >= 0).all() and (I[:, 30:] == -1).all() and np.isinf(D[:, 30:]).all()
    D1, I1 = ix.search_batch(data[:3], 40, 100, nthreads=0)
    assert np.array_equal(D, D1) and np.array_equal(I, I1)


def test_repeated_fit_is_deterministic(oracle):
    # tests/integration_tests.rs:131-188
    xb, xq = bench_data(3000, 16, 20)
    a, b = oracle.Ivf.fit(xb), oracle.Ivf.fit(xb)
    assert np.array_equal(a.centroids(), b.centroids())
    ra, rb = a.search_batch(xq, 5, 4), b.search_batch(xq, 5, 4)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])


def test_nlist_override_and_super_centroid_shards(oracle):
    xb, _ = bench_data(4000, 8, 1)
    ix = oracle.Ivf.fit(xb, nlist=100, max_iters=5)
    assert ix.k_trained == 100 and ix.num_shards == 10 and ix.iters_run == 5
    assert set(ix.centroids_to_shard().tolist()) <= set(range(10))

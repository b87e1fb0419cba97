"""GPU parity: IVF search through the C ABI vs the CPU oracle (bit-exact distances,
identical ids and probe lists).  Reference path: src/ivf_index.rs:190-267."""
import numpy as np
import pytest

from conftest import bench_data

pytestmark = pytest.mark.gpu

# Every test of this module runs with both flavours of the tensor-core filter: seeding pass + main pass (scan mode 2, what
# indexes with giant lists use) and bounds pass + main pass (scan mode 3, what a query that visits few tiles gets); auto
# mode would pick the second for all of these sizes.
_FLAVOUR = {"mode": 0}


@pytest.fixture(autouse=True, params=["seeded", "bounds"])
def flavour(request):
    _FLAVOUR["mode"] = {"seeded": 2, "bounds": 3}[request.param]
    yield request.param
    _FLAVOUR["mode"] = 0


def make_pair(O, ffi, xb, nlist, seed=7, ext=None):
    """Same lists on both sides: random centroids from the data + oracle brute labels."""
    rng = np.random.default_rng(seed)
    cents = xb[rng.choice(len(xb), nlist, replace=False)].copy()
    labels = O.assign_brute_force(xb, cents)
    oix = O.Ivf.from_labels(xb, cents, labels, ext_ids=ext)
    gix = ffi.Index(xb.shape[1]).build_from_labels(xb, cents, labels, ext_ids=ext)
    gix.set_scan_mode(_FLAVOUR["mode"])
    return oix, gix


def check_search(oix, gix, xq, k, nprobe):
    Dg, Ig = gix.search(xq, k, nprobe)
    Do, Io = oix.search_batch(xq, k, nprobe, nthreads=0)
    # bit-exact distances (same sequential fp32 arithmetic as utils.rs:28-30)
    assert np.array_equal(Dg.view(np.uint32), Do.view(np.uint32)), \
        f"distance mismatch: {np.argwhere(Dg.view(np.uint32) != Do.view(np.uint32))[:5]}"
    assert np.array_equal(Ig, Io), f"id mismatch at {np.argwhere(Ig != Io)[:5]}"
    return Dg, Ig


@pytest.mark.parametrize("nprobe", [1, 3, 16])
def test_search_matches_oracle_sparse_regime(oracle, ffi, nprobe):
    xb, xq = bench_data(20000, 64, 300)
    oix, gix = make_pair(oracle, ffi, xb, 200)
    assert gix.nlist == oix.nlist
    check_search(oix, gix, xq, 10, nprobe)


def test_search_matches_oracle_dense_regime(oracle, ffi):
    # few lists, many queries per list -> dense (block-tiled) scan kernel, multi-segment lists
    xb, xq = bench_data(30000, 32, 3000)
    oix, gix = make_pair(oracle, ffi, xb, 12)
    check_search(oix, gix, xq, 10, 4)
    check_search(oix, gix, xq, 1, 12)
    check_search(oix, gix, xq, 32, 2)


@pytest.mark.parametrize("d,k", [(32, 17), (32, 32), (64, 32), (200, 10), (256, 10)])
def test_search_tensor_core_pipeline_shapes(oracle, ffi, d, k):
    # one / two / four K-slices per tile and the 32-entry top-k sets: every ring geometry of scan_tc_kernel
    xb, xq = bench_data(30000, d, 1500)
    oix, gix = make_pair(oracle, ffi, xb, 12)
    check_search(oix, gix, xq, k, 2)


@pytest.mark.parametrize("d,k,tc", [(320, 10, True), (384, 32, True), (512, 10, True), (520, 10, True), (768, 10, True),
                                    (1000, 32, True), (1536, 5, True), (2064, 5, False)])
def test_search_large_dimensions(oracle, ffi, d, k, tc):
    """The reference's own grids and tests use D = 512, 768, 1536 (bench.yaml:1-15, shards_tests.rs:239-270,
    ivf_index_tests.rs:661-686).  Up to D = 512 the query tile stays in shared memory next to a two-stage ring; from there to
    D = 2048 it is streamed through the ring with the list tiles (scan_tc_kernel<.., SA>, query tiles prepared by
    tc_atile_kernel); beyond that the exact FP32 kernels answer.  Same bits every way."""
    xb, xq = bench_data(12000 if d <= 1536 else 4000, d, 700, seed=d)
    oix, gix = make_pair(oracle, ffi, xb, 10)
    gix.set_profiling(True)
    check_search(oix, gix, xq, k, 3)
    assert (gix.stats()["n_tc_items"] > 0) == tc
    gix.set_profiling(False)


@pytest.mark.parametrize("d,nq,nlist,nprobe", [(768, 1, 6, 6), (768, 130, 3, 2), (640, 1031, 40, 7), (1536, 257, 1, 1)])
def test_streamed_query_tiles_ragged(oracle, ffi, d, nq, nlist, nprobe):
    """Streamed query tiles (D > 512) whose row counts are not multiples of 8, one-row tiles, lists with several tiles and a
    short last one, a one-list index."""
    xb, xq = bench_data(6000, d, nq, seed=d + nq)
    oix, gix = make_pair(oracle, ffi, xb, nlist)
    gix.set_profiling(True)
    check_search(oix, gix, xq, 10, nprobe)
    assert gix.stats()["n_tc_items"] > 0
    gix.set_profiling(False)
    Dg, Ig = gix.search(xq, 10, nprobe)  # a second call reuses every buffer
    gix.set_scan_mode(1)
    De, Ie = gix.search(xq, 10, nprobe)
    assert np.array_equal(Dg.view(np.uint32), De.view(np.uint32)) and np.array_equal(Ig, Ie)


@pytest.mark.parametrize("flag", ["1"])
def test_cta_pair_kernel_is_bit_exact(oracle, ffi, flag, monkeypatch):
    """The experimental CTA-pair scan (tcgen05 cta_group::2, VIDX_TC_PAIR=1): same answers as the oracle."""
    monkeypatch.setenv("VIDX_TC_PAIR", flag)
    for d, k, nq in ((64, 10, 900), (128, 17, 1500), (200, 32, 700), (384, 10, 600)):
        xb, xq = bench_data(20000, d, nq, seed=d + 1)
        oix, gix = make_pair(oracle, ffi, xb, 12)
        check_search(oix, gix, xq, k, 3)


@pytest.mark.parametrize("d", [1, 3, 30, 129, 200])
def test_search_odd_dimensions(oracle, ffi, d):
    xb, xq = bench_data(3000, d, 150, seed=d)
    oix, gix = make_pair(oracle, ffi, xb, 20)
    check_search(oix, gix, xq, 5, 4)


def test_coarse_probes_match_oracle(oracle, ffi):
    xb, xq = bench_data(20000, 64, 200)
    oix, gix = make_pair(oracle, ffi, xb, 300)
    lists, dists = gix.coarse_probes(xq, 20)
    for i in range(len(xq)):
        lo, do = oix.probes(xq[i], 20)
        assert np.array_equal(lists[i], lo)
        assert np.array_equal(dists[i].view(np.uint32), do.view(np.uint32))


def test_coarse_large_nlist_radix_select(oracle, ffi):
    # nlist > 4096 exercises the multi-pass radix select
    xb, xq = bench_data(12000, 16, 64)
    oix, gix = make_pair(oracle, ffi, xb, 6000)
    lists, _ = gix.coarse_probes(xq, 50)
    for i in range(len(xq)):
        lo, _ = oix.probes(xq[i], 50)
        assert np.array_equal(lists[i], lo)
    check_search(oix, gix, xq, 10, 50)


def test_large_k_path(oracle, ffi):
    # k > 32: all-distance rows + exact radix select (reference bench default K=100)
    xb, xq = bench_data(20000, 32, 200)
    oix, gix = make_pair(oracle, ffi, xb, 50)
    check_search(oix, gix, xq, 100, 5)
    check_search(oix, gix, xq, 33, 1)


def test_k_larger_than_candidates_pads(oracle, ffi):
    # tests/ivf_index_tests.rs:278-306: k > n returns everything; binding pads with +inf / -1
    xb, xq = bench_data(50, 8, 10)
    oix, gix = make_pair(oracle, ffi, xb, 3)
    D, I = check_search(oix, gix, xq, 64, 3)
    assert (I[:, :50] >= 0).all() and (I[:, 50:] == -1).all() and np.isinf(D[:, 50:]).all()
    D, I = check_search(oix, gix, xq, 20, 1)


def test_nprobe_larger_than_nlist(oracle, ffi):
    # tests/ivf_index_tests.rs:310-336
    xb, xq = bench_data(2000, 16, 20)
    oix, gix = make_pair(oracle, ffi, xb, 10)
    check_search(oix, gix, xq, 10, 1000)


def test_zero_k_or_nprobe_is_invalid_input(oracle, ffi):
    # src/ivf_index.rs:197-202
    xb, xq = bench_data(500, 8, 4)
    _, gix = make_pair(oracle, ffi, xb, 4)
    with pytest.raises(ffi.InvalidInput):
        gix.search(xq, 0, 5)
    with pytest.raises(ffi.InvalidInput):
        gix.search(xq, 5, 0)


def test_self_query_returns_itself_with_payload(oracle, ffi):
    # tests/ivf_index_tests.rs:122-159, tests/api_tests.rs:90: rank 0, distance 0, bit-equal vector
    xb = oracle.create_test_vectors(1000, 16)
    ext = np.arange(1000, dtype=np.uint64) * 3 + 42
    oix, gix = make_pair(oracle, ffi, xb, 8, ext=ext)
    D, I, V = gix.search(xb[:100], 3, 8, include_vectors=True)
    assert (D[:, 0] == 0).all()
    # the ramp fixture repeats every 500 elements, so equal vectors exist: ids may be any duplicate
    for i in range(100):
        assert np.array_equal(V[i, 0], xb[i])
        src = (I[i, 0] - 42) // 3
        assert np.array_equal(xb[src], xb[i])
    Do, Io = oix.search_batch(xb[:100], 3, 8, nthreads=0)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)


def test_duplicate_vectors_tie_order(oracle, ffi):
    # equal distances keep probe-rank / list order (stable sort, ivf_index.rs:265)
    base, xq = bench_data(200, 8, 30)
    xb = np.concatenate([base, base, base])
    oix, gix = make_pair(oracle, ffi, xb, 5)
    check_search(oix, gix, xq, 12, 5)


def test_repeated_search_identical(oracle, ffi):
    # tests/integration_tests.rs:131-188
    xb, xq = bench_data(5000, 32, 100)
    _, gix = make_pair(oracle, ffi, xb, 40)
    a = gix.search(xq, 10, 8)
    for _ in range(4):
        b = gix.search(xq, 10, 8)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_repeated_searches_are_replayed_as_graphs(oracle, ffi):
    """The second search with the same arguments is captured into a CUDA graph, later ones replay it (index.cu, run_cached).
    Interleaved shapes on one handle: a larger batch moves the context's buffers, which must drop the captured graph;
    changing the scan mode (handle epoch) must do the same.  Every answer is compared with the oracle's."""
    xb, xq = bench_data(20000, 64, 900)
    oix, gix = make_pair(oracle, ffi, xb, 30)
    before = ffi.kernel_launch_count()
    small = [check_search(oix, gix, xq[:100], 10, 4) for _ in range(4)]
    per_call = (ffi.kernel_launch_count() - before) // 4
    assert per_call > 5 and ffi.kernel_launch_count() - before == 4 * per_call  # replays count their kernels too
    big = [check_search(oix, gix, xq, 10, 6) for _ in range(3)]      # grows the buffers
    again = [check_search(oix, gix, xq[:100], 10, 4) for _ in range(3)]
    for D, I in small[1:] + again:
        assert np.array_equal(D.view(np.uint32), small[0][0].view(np.uint32)) and np.array_equal(I, small[0][1])
    for D, I in big[1:]:
        assert np.array_equal(D.view(np.uint32), big[0][0].view(np.uint32)) and np.array_equal(I, big[0][1])
    gix.set_scan_mode(1)  # exact kernels: same answer, other launches
    for _ in range(3):
        check_search(oix, gix, xq[:100], 10, 4)
    xq2 = xq[100:200].copy()  # same shape, other queries, same staging buffers: the replay reads the new contents
    for _ in range(2):
        check_search(oix, gix, xq2, 10, 4)


def test_search_stats_and_launch_counter(oracle, ffi):
    xb, xq = bench_data(20000, 64, 500)
    _, gix = make_pair(oracle, ffi, xb, 100)
    gix.set_profiling(True)
    before = ffi.kernel_launch_count()
    gix.search(xq, 10, 4)
    st = gix.stats()
    assert ffi.kernel_launch_count() - before == st["kernel_launches"] > 5
    assert st["n_pairs"] >= 500 * 4
    assert st["scan_bytes_logical"] >= st["scan_bytes_algorithmic"] > 0
    assert st["ms_scan"] > 0 and st["ms_total"] >= st["ms_scan"] >= st["ms_scan_tc"] > 0


# ---- tensor-core pre-filter path (tcgen05 TF32 + exact re-check) ------------------------------
@pytest.mark.parametrize("d,nlist,nq,nprobe,k", [(64, 8, 700, 3, 10), (128, 20, 300, 20, 10), (32, 4, 150, 4, 32),
                                                  (8, 6, 90, 2, 1), (96, 50, 1000, 7, 5)])
def test_tensor_core_scan_equals_exact_scan_and_oracle(oracle, ffi, d, nlist, nq, nprobe, k):
    xb, xq = bench_data(40000, d, nq, seed=d + nq)
    oix, gix = make_pair(oracle, ffi, xb, nlist)
    gix.set_profiling(True)
    Dt, It = check_search(oix, gix, xq, k, nprobe)
    st = gix.stats()
    assert st["n_tc_items"] > 0 and st["n_tc_survivors"] >= nq * min(k, 1) and st["n_tc_overflow"] == 0, st
    assert st["n_dense_items"] == 0 and st["n_sparse_items"] == 0, st
    gix.set_scan_mode(1)
    De, Ie = gix.search(xq, k, nprobe)
    st = gix.stats()
    assert st["n_tc_items"] == 0 and st["n_dense_items"] + st["n_sparse_items"] > 0
    assert np.array_equal(Dt.view(np.uint32), De.view(np.uint32)) and np.array_equal(It, Ie)


def test_tensor_core_survivor_overflow_is_redone_exactly(oracle, ffi):
    # thousands of identical vectors: every one ties with the k-th best, the survivor buffer of
    # those queries overflows and the exact kernels redo them -- results still match the oracle
    rng = np.random.default_rng(5)
    base = rng.standard_normal((3000, 16)).astype(np.float32)
    dup = np.tile(base[:1], (6000, 1))
    xb = np.concatenate([dup, base])
    xq = np.concatenate([base[:1] + 0.01, base[5:40]])
    oix, gix = make_pair(oracle, ffi, xb, 4)
    gix.set_profiling(True)
    check_search(oix, gix, xq, 10, 4)
    st = gix.stats()
    assert st["n_tc_overflow"] >= 1 and st["n_dense_items"] + st["n_sparse_items"] > 0, st


def test_tensor_core_filter_never_drops_a_true_neighbour_on_hard_data(oracle, ffi):
    # large norms + tiny distances: the regime where |q|^2+|v|^2-2q.v cancels catastrophically
    rng = np.random.default_rng(11)
    centers = rng.standard_normal((50, 64)).astype(np.float32) * 100
    xb = (centers[rng.integers(0, 50, 30000)] + rng.standard_normal((30000, 64)).astype(np.float32) * 0.01).astype(np.float32)
    xq = (xb[rng.choice(30000, 400, replace=False)] + rng.standard_normal((400, 64)).astype(np.float32) * 0.001).astype(np.float32)
    oix, gix = make_pair(oracle, ffi, xb, 16)
    check_search(oix, gix, xq, 10, 8)


@pytest.mark.parametrize("d,nlist,nprobe", [(64, 300, 20), (128, 2500, 32), (30, 700, 1), (200, 4096, 8), (768, 900, 16)])
def test_coarse_tensor_core_filter_equals_exact_coarse(oracle, ffi, d, nlist, nprobe):
    # ivf_index.rs:205-220 through the fp16 filter + exact re-check: identical probe lists and bit-identical distances
    xb, xq = bench_data(20000, d, 600, seed=nlist)
    oix, gix = make_pair(oracle, ffi, xb, nlist)
    gix.set_coarse_mode(2)
    lists, dists = gix.coarse_probes(xq, nprobe)
    gix.set_coarse_mode(1)
    lists_e, dists_e = gix.coarse_probes(xq, nprobe)
    assert np.array_equal(lists, lists_e) and np.array_equal(dists.view(np.uint32), dists_e.view(np.uint32))
    for i in range(0, len(xq), 37):
        lo, do = oix.probes(xq[i], nprobe)
        assert np.array_equal(lists[i], lo) and np.array_equal(dists[i].view(np.uint32), do.view(np.uint32))
    gix.set_coarse_mode(2)
    D, I = gix.search(xq, 10, nprobe)
    Do, Io = oix.search_batch(xq, 10, nprobe, nthreads=0)
    assert np.array_equal(D.view(np.uint32), Do.view(np.uint32)) and np.array_equal(I, Io)

"""Multi-GPU pieces on whatever GPUs the box has.
One GPU: N handles, each built as rank r of N (so each holds only its part), searched locally with merge keys and merged --
the answer must equal the single-GPU answer bit for bit, ties included; the library's own collective path is run with a
one-rank NCCL communicator.  Two or more GPUs: real ranks in separate processes over NCCL (vidx_search_multi)."""
import os
import sys

import numpy as np
import pytest

from conftest import bench_data

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same_bits(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def local_then_merge(ffi, parts, xq, k, nprobe):
    import torch
    nq, world = len(xq), len(parts)
    d_xq = torch.from_numpy(xq).cuda()
    D = torch.empty((world, nq, k), dtype=torch.float32, device="cuda")
    I = torch.empty((world, nq, k), dtype=torch.int64, device="cuda")
    K = torch.empty((world, nq, k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for r, p in enumerate(parts):
        p.search_local_device(d_xq.data_ptr(), nq, k, nprobe, D[r].data_ptr(), I[r].data_ptr(), K[r].data_ptr(), 0)
    oD = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    oI = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ffi.merge_topk_keyed_device(0, D.data_ptr(), I.data_ptr(), K.data_ptr(), world, nq, k, oD.data_ptr(), oI.data_ptr(), 0)
    torch.cuda.synchronize()
    return oD.cpu().numpy(), oI.cpu().numpy()


@pytest.mark.parametrize("mode", ["shards", "ranges"])
@pytest.mark.parametrize("k", [10, 100])
def test_resident_partition_merges_to_single_gpu_answer(ffi, mode, k):
    xb, xq = bench_data(30000, 32, 300)
    xb[5000:5040] = xb[100]                       # duplicates: equal distances across lists / ranks
    full = ffi.Index(32).build(xb)
    D0, I0 = full.search(xq, k, 16)
    world = 4
    parts = []
    for r in range(world):
        p = ffi.Index(32)
        p.set_partition_mode(mode)
        p.set_partition(r, world)                 # BEFORE the build: only this rank's part reaches HBM
        p.build(xb)
        assert p.partition_kind == mode and p.nlist == full.nlist and p.ntotal == full.ntotal
        parts.append(p)
    res = [p.resident_vectors for p in parts]
    assert sum(res) == 30000 and max(res) < 30000
    assert sum(p.resident_bytes for p in parts) < 1.5 * full.resident_bytes
    if mode == "ranges":                          # shards balance only as well as the reference's shard sizes allow
        assert max(res) <= 0.6 * 30000
    D, I = local_then_merge(ffi, parts, xq, k, 16)
    assert same_bits(D, D0)
    assert np.array_equal(I, I0), "keyed merge must reproduce the single-GPU order, ties included"
    # a resident part cannot be re-partitioned
    with pytest.raises(ffi.VidxError):
        parts[0].set_partition(1, world)
    parts[0].set_partition(0, world)              # the same partition is fine


def test_partitioned_load_reads_only_the_owned_part(ffi, tmp_path):
    xb, xq = bench_data(20000, 24, 128)
    full = ffi.Index(24).build(xb)
    full.save(str(tmp_path / "index"), str(tmp_path / "shards"))
    D0, I0 = full.search(xq, 10, 12)
    for mode in ("shards", "ranges"):
        world, parts = 3, []
        for r in range(world):
            p = ffi.Index(24)
            p.set_partition_mode(mode)
            p.set_partition(r, world)
            p.load(str(tmp_path / "index"), str(tmp_path / "shards"))
            assert p.ntotal == 20000 and p.partition_kind == mode and not p.load_warnings()
            parts.append(p)
        assert sum(p.resident_vectors for p in parts) == 20000
        D, I = local_then_merge(ffi, parts, xq, 10, 12)
        assert same_bits(D, D0) and np.array_equal(I, I0)
    # shard-partitioned ranks saving into one directory reproduce the single-GPU files
    parts = []
    for r in range(2):
        p = ffi.Index(24)
        p.set_partition_mode("shards")
        p.set_partition(r, 2)
        p.build(xb)
        p.save(str(tmp_path / "index2"), str(tmp_path / "shards2"))
        parts.append(p)
    for name in sorted(os.listdir(tmp_path / "shards")):
        a = open(tmp_path / "shards" / name, "rb").read()
        b = open(tmp_path / "shards2" / name, "rb").read()
        # timestamps are wall-clock seconds of each build (vector_store.rs:36-40): compare sizes and the id columns
        assert len(a) == len(b), name
    assert open(tmp_path / "index" / "index.bin", "rb").read() == open(tmp_path / "index2" / "index.bin", "rb").read()
    back = ffi.Index(24).load(str(tmp_path / "index2"), str(tmp_path / "shards2"))
    D, I = back.search(xq, 10, 12)
    assert same_bits(D, D0) and np.array_equal(I, I0)


@pytest.mark.parametrize("world,nparts", [(4, 2), (4, 1), (6, 3), (8, 2)])
@pytest.mark.parametrize("nq,k", [(301, 10), (7, 100), (1, 10)])
def test_rank_grid_emulated_on_one_gpu(ffi, world, nparts, nq, k):
    """vidx_search_multi's grid of index parts x query groups, with this process playing every rank: rank r searches part
    r % P for the queries vidx_grid_plan gives group r / P, the packed runs are laid out as the all-gather would leave them,
    and the grouped merge must return the single-GPU answer for the whole batch (ragged last group, groups without queries)."""
    import torch
    xb, xq = bench_data(30000, 32, 301)
    xq = xq[:nq]
    xb[5000:5040] = xb[100]
    full = ffi.Index(32).build(xb)
    D0, I0 = full.search(xq, k, 16)
    parts = []
    for p in range(nparts):
        ix = ffi.Index(32)
        if nparts > 1:
            ix.set_partition(p, nparts)
        parts.append(ix.build(xb))
    plans = [ffi.grid_plan(nq, world, nparts, r) for r in range(world)]
    per_group = plans[0]["per_group"]
    # the groups tile the batch, and the ranks of a group tile its coarse work
    covered = np.zeros(nq, np.int32)
    for r, pl in enumerate(plans):
        assert pl["group"] == r // nparts and pl["per_group"] == per_group
        assert pl["q_lo"] <= pl["coarse_lo"] <= pl["coarse_hi"] <= pl["q_hi"] or pl["coarse_lo"] == pl["coarse_hi"]
        covered[pl["coarse_lo"]:pl["coarse_hi"]] += 1
    assert (covered == 1).all()
    d_xq = torch.from_numpy(xq).cuda()
    D = torch.full((world, per_group, k), float("inf"), dtype=torch.float32, device="cuda")
    I = torch.full((world, per_group, k), -1, dtype=torch.int64, device="cuda")
    K = torch.full((world, per_group, k), -1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for r, pl in enumerate(plans):
        n = pl["q_hi"] - pl["q_lo"]
        if n:
            parts[r % nparts].search_local_device(d_xq[pl["q_lo"]:].data_ptr(), n, k, 16, D[r].data_ptr(), I[r].data_ptr(),
                                                  K[r].data_ptr(), 0)
    oD = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    oI = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ffi.merge_topk_grid_device(0, D.data_ptr(), I.data_ptr(), K.data_ptr(), nparts, per_group, nq, k, oD.data_ptr(), oI.data_ptr(), 0)
    torch.cuda.synchronize()
    assert same_bits(oD.cpu().numpy(), D0) and np.array_equal(oI.cpu().numpy(), I0)


@pytest.mark.parametrize("exchange", ["", "1"])
def test_collective_path_with_a_one_rank_communicator(ffi, exchange, monkeypatch):
    """vidx_comm_init + vidx_search_multi on world = 1: the probe all-gather, the bound exchange (forced on with
    VIDX_BOUNDS_EXCHANGE: a one-rank world would skip it), the packed all-gather and the keyed merge all run (over one rank)
    and must return the plain answer; k > 32 takes the large-k path."""
    if exchange:
        monkeypatch.setenv("VIDX_BOUNDS_EXCHANGE", exchange)
    xb, xq = bench_data(20000, 32, 333)
    ix = ffi.Index(32).build(xb)
    ix.comm_init(0, 1, ffi.comm_unique_id())
    assert ix.comm_version.startswith("NCCL ")
    for k, nprobe in ((10, 8), (1, 1), (100, 20), (10, 10 ** 6)):
        D0, I0 = ix.search(xq, k, nprobe)
        D, I = ix.search_multi(xq, k, nprobe)
        assert same_bits(D, D0) and np.array_equal(I, I0), (k, nprobe)
    ix.comm_destroy()
    with pytest.raises(ffi.VidxError):
        ix.search_multi(xq, 10, 8)


def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "vector-indexer_b200"))
    import torch
    import torch.distributed as dist
    from vector_indexer_py import _ffi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # host-side plumbing only
    xb, xq = bench_data(40000, 48, 500)
    xb[7000:7030] = xb[3]
    ok = True
    full = _ffi.Index(48, rank).build(xb) if rank == 0 else None
    # parts x groups grids: parts = world is the plain sharded index, parts = 1 a replica per rank with the batch split
    # by query, anything between both at once (4 GPUs: 2 parts x 2 query groups)
    for parts in sorted({p for p in (world, 1, 2) if world % p == 0}, reverse=True):
        for mode in ("shards", "ranges") if parts > 1 else ("auto",):
            ix = _ffi.Index(48, rank)
            ix.set_partition_mode(mode)
            ix.set_partition(rank % parts, parts)
            ix.build(xb)
            obj = [_ffi.comm_unique_id() if rank == 0 else None]   # rank 0's NCCL id, handed round by the host program
            dist.broadcast_object_list(obj, src=0)
            ix.comm_init(rank, world, obj[0])
            for nq, k, nprobe in ((500, 10, 12), (333, 100, 7), (1, 10, 12)):
                D, I = ix.search_multi(xq[:nq], k, nprobe)
                if rank == 0:
                    D0, I0 = full.search(xq[:nq], k, nprobe)
                    ok = ok and same_bits(D, D0) and np.array_equal(I, I0)
            if rank == 0:
                ok = ok and (ix.resident_vectors < 40000 if parts > 1 else ix.resident_vectors == 40000)
            ix.comm_destroy()
    # a communicator that does not fit the partition is refused on every rank (no collective is entered)
    bad = _ffi.Index(48, rank)
    bad.set_partition(0, 3)
    bad.build(xb[:5000])
    obj = [_ffi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    bad.comm_init(rank, world, obj[0])
    try:
        bad.search_multi(xq[:8], 10, 4)
        ok = False
    except _ffi.VidxError:
        pass
    bad.comm_destroy()
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_two_ranks_over_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 2000
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"rank{r}.txt").read() for r in range(2)] == ["1", "1"]


def test_concurrent_searching_threads(ffi):
    """tests/ivf_index_tests.rs:768-807: several OS threads search one shared index at once; every thread must get the
    answer a lone search gets.  Each call runs on its own context (stream + scratch) of the handle."""
    import threading
    xb, xq = bench_data(30000, 32, 2000)
    ix = ffi.Index(32).build(xb)
    want = [ix.search(xq[i * 500:(i + 1) * 500], 10, 4 + 4 * i) for i in range(4)]
    got, errs = [None] * 4, []

    def work(i):
        try:
            for _ in range(8):
                got[i] = ix.search(xq[i * 500:(i + 1) * 500], 10, 4 + 4 * i)
                if not (same_bits(got[i][0], want[i][0]) and np.array_equal(got[i][1], want[i][1])):
                    errs.append(i)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def test_device_searches_on_two_streams_do_not_share_scratch(ffi):
    """vidx_search_device on two caller streams back to back: each enqueue takes its own context, and a reused
    context waits for the search that used it last."""
    import torch
    xb, xq = bench_data(30000, 32, 4000)
    ix = ffi.Index(32).build(xb)
    d_xq = torch.from_numpy(xq).cuda()
    outs = [(torch.empty((2000, 10), dtype=torch.float32, device="cuda"), torch.empty((2000, 10), dtype=torch.int64, device="cuda"))
            for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(6):
        for s in range(2):
            ix.search_device(d_xq[s * 2000:].data_ptr(), 2000, 10, 8, outs[s][0].data_ptr(), outs[s][1].data_ptr(), streams[s].cuda_stream)
    torch.cuda.synchronize()
    for s in range(2):
        D0, I0 = ix.search(xq[s * 2000:(s + 1) * 2000], 10, 8)
        assert same_bits(outs[s][0].cpu().numpy(), D0) and np.array_equal(outs[s][1].cpu().numpy(), I0)


def test_build_device_equals_build(ffi):
    import torch
    xb, xq = bench_data(12000, 40, 100)
    a = ffi.Index(40).build(xb)
    d = torch.from_numpy(xb).cuda()
    torch.cuda.synchronize()
    b = ffi.Index(40).build_device(d.data_ptr(), len(xb))
    assert same_bits(a.centroids(), b.centroids()) and np.array_equal(a.list_sizes(), b.list_sizes())
    ra, rb = a.search(xq, 10, 8), b.search(xq, 10, 8)
    assert same_bits(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])


def test_nan_query_is_answered_like_the_exact_kernels(ffi):
    xb, xq = bench_data(8000, 16, 64)
    xq = xq.copy()
    xq[5, 3] = np.nan
    ix = ffi.Index(16).build(xb)
    D, I = ix.search(xq, 10, 8)
    ix.set_scan_mode(1)
    De, Ie = ix.search(xq, 10, 8)
    assert np.array_equal(D.view(np.uint32), De.view(np.uint32)) and np.array_equal(I, Ie)

"""N > 1 host logic on CPU: world_size-2 gloo processes.  Each rank owns the shards the
library's partition rule gives it, scans only those lists (the oracle stands in for the
device-local scan here -- it is the checker, not the product), the per-rank top-k are
exchanged with all_gather in the [run][q][k] layout the device merge kernel consumes, and
the merged answer must equal the unpartitioned search."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def merge_runs(D, I):
    """Restatement of merge_runs_kernel: stable by (distance, run, position)."""
    nruns, nq, k = D.shape
    oD = np.full((nq, k), np.inf, np.float32)
    oI = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        d = D[:, q, :].reshape(-1)
        i = I[:, q, :].reshape(-1)
        keep = i >= 0
        order = np.argsort(d[keep], kind="stable")[:k]
        oD[q, :len(order)] = d[keep][order]
        oI[q, :len(order)] = i[keep][order]
    return oD, oI


def worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "vector-indexer_b200"))
    import oracle as O
    from vector_indexer_py import _ffi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)
    xb = rng.standard_normal((3000, 16)).astype(np.float32)
    xq = rng.standard_normal((40, 16)).astype(np.float32)
    ix = O.Ivf.fit(xb, seed=42)  # deterministic: every rank builds the same index
    c2s, sizes = ix.centroids_to_shard(), ix.list_sizes()
    shard_sizes = np.bincount(c2s, weights=sizes, minlength=ix.num_shards).astype(np.uint64)
    owner = _ffi.partition_shards(shard_sizes, world)
    kind, owner2 = _ffi.partition_plan(sizes, c2s, ix.num_shards, world, mode="shards")
    assert kind == "shards" and np.array_equal(owner, owner2)
    k, nprobe = 10, 12
    Dfull, Ifull = ix.search_batch(xq, k, nprobe)
    ix.set_list_mask(owner[c2s] == rank)
    Dl, Il = ix.search_batch(xq, k, nprobe)
    gD = [torch.empty(40, k) for _ in range(world)]
    gI = [torch.empty(40, k, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gD, torch.from_numpy(Dl))
    dist.all_gather(gI, torch.from_numpy(Il))
    D, I = merge_runs(torch.stack(gD).numpy(), torch.stack(gI).numpy())
    ok = np.array_equal(D, Dfull) and np.array_equal(I, Ifull)
    loads = np.bincount(owner, weights=shard_sizes.astype(np.float64), minlength=world)
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{int(ok)} {loads.tolist()} {owner.tolist()}")
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [open(tmp_path / f"rank{r}.txt").read() for r in range(2)]
    assert all(o.startswith("1 ") for o in outs), outs
    assert outs[0][2:] == outs[1][2:]  # both ranks derived the same partition


def test_split_decision_balances_vectors_and_scan_work(ffi):
    """vidx_partition_plan: shards when they balance, ranges when one rank would hold too many vectors OR too much expected scan
    work (sum of len^2: a query probes a list about as often as a vector falls into it)."""
    rng = np.random.default_rng(1)
    # 64 lists of similar size in 8 shards: shards balance in both measures
    sizes = rng.integers(900, 1100, 64)
    shard = np.arange(64) % 8
    for world in (2, 4, 8):
        kind, owner = ffi.partition_plan(sizes, shard, 8, world)
        assert kind == "shards" and set(owner.tolist()) == set(range(world))
    # one shard holds nearly everything (the SIFT-1M-shaped bench index): vectors cannot be balanced
    sizes2 = np.array([100_000] * 10 + [1] * 54)
    shard2 = np.array([0] * 10 + list(np.arange(54) % 7 + 1))
    assert ffi.partition_plan(sizes2, shard2, 8, 2)[0] == "ranges"
    # equal vector counts per shard, but shard 0 is ONE giant list and the others are many small ones (configs[4]-like):
    # vectors balance, the expected scan work does not
    sizes3 = np.array([8000] + [100] * 80 * 7)
    shard3 = np.array([0] + [1 + i // 80 for i in range(560)])
    loads = np.bincount(shard3, weights=sizes3, minlength=8)
    assert loads.max() == loads.min() == 8000
    assert ffi.partition_plan(sizes3, shard3, 8, 8)[0] == "ranges"
    assert ffi.partition_plan(sizes3, shard3, 8, 8, mode="shards")[0] == "shards"
    assert ffi.partition_plan(sizes, shard, 8, 4, mode="ranges")[0] == "ranges"
    assert ffi.partition_plan(sizes3, shard3, 8, 1)[0] == "shards"


def test_partition_rule_balances(ffi):
    owner = ffi.partition_shards([100, 90, 50, 40, 10, 10], 2)
    assert owner.tolist() == [0, 1, 1, 0, 0, 1]  # 100|90, 50->r1, 40->r0, tie 140/140 -> lower rank
    sizes = np.random.default_rng(0).integers(1, 1000, 64)
    for world in (2, 4, 8):
        o = ffi.partition_shards(sizes, world)
        loads = np.bincount(o, weights=sizes, minlength=world)
        assert loads.max() - loads.min() <= sizes.max() and set(o.tolist()) == set(range(world))
    assert ffi.partition_shards([5, 5, 5], 1).tolist() == [0, 0, 0]


@pytest.mark.parametrize("nq", [0, 1, 7, 10000, 10007])
@pytest.mark.parametrize("world,parts", [(1, 1), (2, 1), (2, 2), (4, 2), (8, 1), (8, 4), (8, 8), (6, 3)])
def test_grid_plan_tiles_the_batch(nq, world, parts):
    """vidx_grid_plan (host only): world = parts x groups.  The groups' query ranges tile [0, nq) in rank order, every rank of a
    group reports the same range, the coarse slices of all ranks tile the batch once, and a rank's slice lies in its group."""
    sys.path.insert(0, os.path.join(ROOT, "vector-indexer_b200"))
    from vector_indexer_py import _ffi
    plans = [_ffi.grid_plan(nq, world, parts, r) for r in range(world)]
    per_group = plans[0]["per_group"]
    seen = np.zeros(nq, np.int32)
    coarse = np.zeros(nq, np.int32)
    for r, p in enumerate(plans):
        g = r // parts
        assert p["group"] == g and p["per_group"] == per_group
        assert p["q_lo"] == min(nq, g * per_group) and p["q_hi"] == min(nq, (g + 1) * per_group)
        assert p["q_hi"] - p["q_lo"] <= per_group
        if r % parts == 0:
            seen[p["q_lo"]:p["q_hi"]] += 1
        coarse[p["coarse_lo"]:p["coarse_hi"]] += 1
        if p["coarse_hi"] > p["coarse_lo"]:
            assert p["q_lo"] <= p["coarse_lo"] and p["coarse_hi"] <= p["q_hi"]
    assert (seen == 1).all() and (coarse == 1).all()
    with pytest.raises(_ffi.VidxError):
        _ffi.grid_plan(nq, 4, 3, 0)

"""Size-independent properties at BASELINE configs[1] full size (1M x 128 fp32, nlist = 1024, nq = 10 000, k = 10), where the
oracle is too slow to check everything: ordering, idempotence, the tensor-core filter against the exact kernels, probing every
list against brute force, the partitioned answer against the single-GPU one, and the oracle itself on a sample."""
import numpy as np
import pytest

from conftest import bench_data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(ffi):
    xb, xq = bench_data(1_000_000, 128, 10_000)
    ix = ffi.Index(128).build(xb, seed=42, nlist=1024)
    return xb, xq, ix


def test_ordering_counts_and_idempotence(big):
    xb, xq, ix = big
    D, I = ix.search(xq, 10, 8)
    assert D.shape == (10_000, 10) and np.all(np.isfinite(D)) and np.all(I >= 0) and np.all(I < len(xb))
    assert np.all(np.diff(D, axis=1) >= 0), "distances must ascend (ivf_index_tests.rs:211-220)"
    assert all(len(set(r.tolist())) == 10 for r in I[::97]), "no vector twice in one answer"
    D2, I2 = ix.search(xq, 10, 8)
    assert np.array_equal(D.view(np.uint32), D2.view(np.uint32)) and np.array_equal(I, I2), "repeated search identical"
    # returned distances are the reference arithmetic of the returned ids (utils.rs:28-30: sequential fp32 sum)
    for q in (0, 4999, 9999):
        acc = np.float32(0)
        for t in (xq[q] - xb[I[q, 0]]).astype(np.float32):
            acc = np.float32(acc + np.float32(t * t))
        assert acc.view(np.uint32) == D[q, 0].view(np.uint32)


def test_filter_equals_exact_kernels_on_a_sample(big):
    xb, xq, ix = big
    s = xq[1234:1234 + 384]
    D, I = ix.search(s, 10, 8)
    ix.set_scan_mode(1)
    try:
        De, Ie = ix.search(s, 10, 8)
    finally:
        ix.set_scan_mode(0)
    assert np.array_equal(D.view(np.uint32), De.view(np.uint32)) and np.array_equal(I, Ie)


def test_probing_every_list_is_brute_force(big):
    xb, xq, ix = big
    s = xq[:64]
    D, I = ix.search(s, 10, ix.nlist)
    # float64 brute force in blocks
    best = np.full((64, 10), np.inf)
    besti = np.full((64, 10), -1, np.int64)
    for b0 in range(0, len(xb), 100_000):
        blk = xb[b0:b0 + 100_000].astype(np.float64)
        dd = (s.astype(np.float64) ** 2).sum(1)[:, None] - 2 * s.astype(np.float64) @ blk.T + (blk ** 2).sum(1)[None, :]
        cat = np.concatenate([best, dd], 1)
        cati = np.concatenate([besti, np.arange(b0, b0 + len(blk))[None, :].repeat(64, 0)], 1)
        o = np.argsort(cat, 1)[:, :10]
        best, besti = np.take_along_axis(cat, o, 1), np.take_along_axis(cati, o, 1)
    assert np.allclose(D, best, rtol=1e-5), "within 1e-5 relative of float64 brute force"
    assert (I == besti).mean() > 0.999  # identical modulo fp32 near-ties


def test_partitioned_answer_equals_single_gpu(big, ffi):
    import torch
    xb, xq, ix = big
    s = xq[:512]
    D0, I0 = ix.search(s, 10, 8)
    world = 4
    Ds, Is = [], []
    try:
        for r in range(world):
            ix.set_partition(r, world)
            assert ix.partition_kind == "ranges"  # the reference's shards cannot be balanced on this index
            D, I = ix.search(s, 10, 8)
            Ds.append(D)
            Is.append(I)
    finally:
        ix.set_partition(0, 1)
    dD, dI = torch.tensor(np.stack(Ds)).cuda(), torch.tensor(np.stack(Is)).cuda()
    oD = torch.empty((512, 10), dtype=torch.float32, device="cuda")
    oI = torch.empty((512, 10), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ffi.merge_topk_device(0, dD.data_ptr(), dI.data_ptr(), world, 512, 10, oD.data_ptr(), oI.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(oD.cpu().numpy().view(np.uint32), D0.view(np.uint32))
    assert (oI.cpu().numpy() == I0).mean() > 0.999


def test_oracle_parity_on_a_sample(big, oracle):
    xb, xq, ix = big
    oix = oracle.Ivf.from_labels(xb, ix.train_centroids(), ix.train_labels())
    s = xq[7000:7016]
    for nprobe in (1, 8):
        D, I = ix.search(s, 10, nprobe)
        Do, Io = oix.search_batch(s, 10, nprobe, nthreads=0)
        assert np.array_equal(D.view(np.uint32), Do.view(np.uint32)) and np.array_equal(I, Io)

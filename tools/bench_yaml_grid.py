"""The reference's own (orphan) Rust bench grid, /root/reference/bench.yaml:1-15, on one GPU: D in {128, 256, 768} x
N in {100 000, 500 000} x n_probe in {8, 16, 32}, 10 000 queries, k = 10, seed 42, index built by the library's own
k-means (vidx_build, the reference's heuristics).  Per setting: device time per 10 000-query batch (CUDA events on the
stream the kernels run on, 10 reps after 3 warm-ups), QPS, set recall@10 against a float64 brute force, and -- for
D = 768, where the query tile is streamed through the ring -- the same search on the exact FP32 kernels (scan mode 1),
whose answer must be bit-identical.  One JSON line per setting; `--dims 768 --counts 100000` cuts the grid."""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi

ap = argparse.ArgumentParser()
ap.add_argument('--dims', default='128,256,768')
ap.add_argument('--counts', default='100000,500000')
ap.add_argument('--nprobes', default='8,16,32')
ap.add_argument('--nq', type=int, default=10_000)
ap.add_argument('--reps', type=int, default=10)
a = ap.parse_args()
k = 10
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)


def ground_truth(xb, xq):  # float64 brute force, blocks of rows (independent of the library)
    q = torch.from_numpy(xq).cuda().double()
    best_d = torch.full((len(xq), k), float('inf'), dtype=torch.float64, device='cuda')
    best_i = torch.full((len(xq), k), -1, dtype=torch.int64, device='cuda')
    for r0 in range(0, len(xb), 32768):
        b = torch.from_numpy(xb[r0:r0 + 32768]).cuda().double()
        d = (q * q).sum(1)[:, None] - 2.0 * q @ b.T + (b * b).sum(1)[None, :]
        dd = torch.cat([best_d, d], 1)
        ii = torch.cat([best_i, torch.arange(r0, r0 + len(b), device='cuda')[None, :].expand(len(xq), -1)], 1)
        best_d, sel = torch.topk(dd, k, dim=1, largest=False)
        best_i = torch.gather(ii, 1, sel)
    return best_i.cpu().numpy()


def timed(ix, d_xq, nq, npb, d_D, d_I, reps):
    run = lambda: ix.search_device(d_xq.data_ptr(), nq, k, npb, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(ts)
    for _ in range(reps): run()
    e1.record(ts); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for d in [int(v) for v in a.dims.split(',')]:
    for n in [int(v) for v in a.counts.split(',')]:
        rng = np.random.default_rng(42)
        xb = rng.standard_normal((n, d), dtype=np.float32); xq = rng.standard_normal((a.nq, d), dtype=np.float32)
        ix = _ffi.Index(d, 0).build(xb, seed=42)
        gt = ground_truth(xb, xq)
        d_xq = torch.from_numpy(xq).cuda()
        d_D = torch.empty((a.nq, k), device='cuda'); d_I = torch.empty((a.nq, k), dtype=torch.int64, device='cuda')
        for npb in [int(v) for v in a.nprobes.split(',')]:
            ix.set_scan_mode(0)
            ms = timed(ix, d_xq, a.nq, npb, d_D, d_I, a.reps)
            D, I = d_D.cpu().numpy().copy(), d_I.cpu().numpy().copy()
            ix.set_profiling(True)
            ix.search_device(d_xq.data_ptr(), a.nq, k, npb, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
            st = ix.stats(); ix.set_profiling(False)
            rec = float(np.mean([len(set(I[i]) & set(gt[i])) / k for i in range(a.nq)]))
            line = dict(d=d, n=n, nlist=ix.nlist, nprobe=npb, nq=a.nq, k=k, ms_per_batch=ms, qps=a.nq / ms * 1e3, recall_at_10=rec,
                        tensor_core_items=st['n_tc_items'], ms_scan=st['ms_scan'], ms_coarse=st['ms_coarse'],
                        survivors_per_query=st['n_tc_survivors'] / a.nq, queries_redone_exactly=st['n_tc_overflow'])
            if d > 512:
                ix.set_scan_mode(1)
                ms_e = timed(ix, d_xq, a.nq, npb, d_D, d_I, max(1, a.reps // 5))
                same = np.array_equal(D.view(np.uint32), d_D.cpu().numpy().view(np.uint32)) and np.array_equal(I, d_I.cpu().numpy())
                line.update(ms_per_batch_exact_fp32=ms_e, speedup_vs_exact_fp32=ms_e / ms, bit_identical_to_exact=bool(same))
            print(json.dumps(line), flush=True)
        del ix
os._exit(0)

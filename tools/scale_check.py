"""BASELINE configs[3] shape on ONE GPU: 10M x 128, reference nlist heuristic (12 652), n_probe = 32 -- build + search at scale,
filter vs exact kernels and vs float64 brute force on a sample (no oracle: too slow at this size)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi
n, d, nq, k, nprobe = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, 128, 10_000, 10, 32
rng = np.random.default_rng(42)
t0 = time.perf_counter()
xb = np.empty((n, d), np.float32)
for b0 in range(0, n, 1_000_000):
    xb[b0:b0 + 1_000_000] = rng.standard_normal((min(1_000_000, n - b0), d)).astype(np.float32)
xq = rng.standard_normal((nq, d)).astype(np.float32)
print(f"data {time.perf_counter() - t0:.1f} s", flush=True)
t0 = time.perf_counter(); ix = _ffi.Index(d, 0).build(xb, seed=42); print(f"build {time.perf_counter() - t0:.2f} s nlist {ix.nlist} shards {ix.num_shards}", flush=True)
sizes = np.sort(ix.list_sizes())[::-1]; print("largest lists", sizes[:8], "median", int(np.median(sizes)))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
d_xq = torch.from_numpy(xq).cuda(); d_D = torch.empty((nq, k), device='cuda'); d_I = torch.empty((nq, k), dtype=torch.int64, device='cuda')
ix.set_profiling(True)
if len(sys.argv) > 2: ix.set_coarse_mode(int(sys.argv[2]))
for _ in range(3):
    ix.search_device(d_xq.data_ptr(), nq, k, nprobe, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
s = ix.stats(); print({kk: round(s[kk], 3) for kk in s if kk.startswith('ms_')}, 'pairs/query', s['tc_mma_flops'] // (2 * 128 * nq), 'survivors/query', s['n_tc_survivors'] / nq, 'overflow', s['n_tc_overflow'], 'QPS', round(nq / (s['ms_total'] / 1e3)))
ix.set_profiling(False)
D = d_D.cpu().numpy(); I = d_I.cpu().numpy()
smp = xq[:64]
ix.set_scan_mode(1); De, Ie = ix.search(smp, k, nprobe); ix.set_scan_mode(0)
print("filter == exact kernels on 64 queries:", np.array_equal(D[:64].view(np.uint32), De.view(np.uint32)) and np.array_equal(I[:64], Ie))
Da, Ia = ix.search(smp[:16], k, ix.nlist)
best = np.full((16, k), np.inf); besti = np.full((16, k), -1, np.int64); s64 = smp[:16].astype(np.float64)
for b0 in range(0, n, 500_000):
    blk = xb[b0:b0 + 500_000].astype(np.float64)
    dd = (s64 ** 2).sum(1)[:, None] - 2 * s64 @ blk.T + (blk ** 2).sum(1)[None, :]
    cat = np.concatenate([best, dd], 1); cati = np.concatenate([besti, np.arange(b0, b0 + len(blk))[None, :].repeat(16, 0)], 1)
    o = np.argsort(cat, 1)[:, :k]; best, besti = np.take_along_axis(cat, o, 1), np.take_along_axis(cati, o, 1)
print("all lists probed == float64 brute force on 16 queries:", np.allclose(Da, best, rtol=1e-5), (Ia == besti).mean(),
      "recall@10 at nprobe", nprobe, np.mean([len(set(I[i]) & set(besti[i])) / k for i in range(16)]))
os._exit(0)

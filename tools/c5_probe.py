"""BASELINE configs[4] shape on ONE GPU at reduced n (default 8M x 96, nlist = 65 536, n_probe = 32, nq = 10 000): build time,
per-stage search times, and coarse quantization both ways (exact FP32 kernel vs tensor-core filter + exact re-check)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
d, nq, k, nlist, npb = 96, 10_000, 10, 65536, 32
g = torch.Generator(device='cuda'); g.manual_seed(42)
xb = torch.empty((n, d), dtype=torch.float32, device='cuda')
for i0 in range(0, n, 4_000_000):
    i1 = min(n, i0 + 4_000_000); xb[i0:i1] = torch.randn((i1 - i0, d), generator=g, device='cuda')
xq = torch.randn((nq, d), generator=g, device='cuda'); torch.cuda.synchronize()
t0 = time.time(); ix = _ffi.Index(d, 0).build_device(xb.data_ptr(), n, seed=42, nlist=nlist); print(f'build {time.time() - t0:.1f} s nlist {ix.nlist} shards {ix.num_shards} resident {ix.resident_bytes / 1e9:.2f} GB', flush=True)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
d_D = torch.empty((nq, k), device='cuda'); d_I = torch.empty((nq, k), dtype=torch.int64, device='cuda')
base = None
for cm in (1, 2, 0):
    ix.set_coarse_mode(cm); ix.set_profiling(True)
    acc = {}
    for it in range(5):
        ix.search_device(xq.data_ptr(), nq, k, npb, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
        if it >= 2:
            s = ix.stats()
            for kk in s: acc[kk] = acc.get(kk, 0) + s[kk] / 3 if kk.startswith('ms_') else s[kk]
    ix.set_profiling(False)
    D, I = d_D.cpu().numpy().copy(), d_I.cpu().numpy().copy()
    if base is None: base = (D, I)
    ok = np.array_equal(D.view(np.uint32), base[0].view(np.uint32)) and np.array_equal(I, base[1])
    alg = 2.0 * d * nq * ix.nlist
    print(f"coarse_mode {cm}: " + ' '.join(f"{kk[3:]} {acc[kk]:.3f}" for kk in acc if kk.startswith('ms_')) +
          f" | coarse+select {acc['ms_coarse'] + acc['ms_select']:.3f} ms = {alg / (acc['ms_coarse'] + acc['ms_select']) / 1e9:.1f} TFLOP/s algorithmic (2*D*nq*nlist)"
          f" items {acc['n_tc_items']} surv/q {acc['n_tc_survivors'] / nq:.1f} hbm alg {acc['scan_bytes_algorithmic'] / 1e9:.2f} GB -> {acc['scan_bytes_algorithmic'] / acc['ms_scan_tc'] / 1e6:.0f} GB/s {'OK' if ok else 'MISMATCH'}", flush=True)
os._exit(0)

"""configs[1] (1M x 128, nlist 1024, n_probe 8) on ONE GPU for batches of 10 000 / 2^i queries: what one rank of a
replica grid (vidx_search_multi with an unpartitioned index on N GPUs) has to do at N = 1, 2, 4, 8, 16.  Per batch size:
device time per search (CUDA events, 20 reps) for scan modes auto / seeded / bounds-first, and the stage split."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi
n, d, k, npb = 1_000_000, 128, 10, int(os.environ.get('NPROBE', 8))
rng = np.random.default_rng(42)
xb = rng.standard_normal((n, d), dtype=np.float32); xq = rng.standard_normal((10_000, d), dtype=np.float32)
ix = _ffi.Index(d, 0).build(xb, seed=42, nlist=1024)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
d_xq = torch.from_numpy(xq).cuda()
d_D = torch.empty((10_000, k), device='cuda'); d_I = torch.empty((10_000, k), dtype=torch.int64, device='cuda')
for nq in [int(v) for v in os.environ.get('NQS', '10000,5000,2500,1250,625').split(',')]:
    base = None
    for sm in [int(v) for v in os.environ.get('SCAN_MODES', '0,2,3').split(',')]:
        ix.set_scan_mode(sm)
        run = lambda: ix.search_device(d_xq.data_ptr(), nq, k, npb, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream)
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(ts)
        for _ in range(20): run()
        e1.record(ts); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ix.set_profiling(True); acc = {}
        for it in range(3):
            run(); torch.cuda.synchronize(); s = ix.stats()
            for kk in s: acc[kk] = acc.get(kk, 0) + s[kk] / 3 if kk.startswith('ms_') else s[kk]
        ix.set_profiling(False)
        D, I = d_D[:nq].cpu().numpy().copy(), d_I[:nq].cpu().numpy().copy()
        if base is None: base = (D, I)
        ok = np.array_equal(D.view(np.uint32), base[0].view(np.uint32)) and np.array_equal(I, base[1])
        print(f"nq {nq:6d} scan_mode {sm}: {ms:.3f} ms/search = {nq / ms / 1e3:.2f} M QPS | " +
              ' '.join(f"{kk[3:]} {acc[kk]:.3f}" for kk in acc if kk.startswith('ms_')) +
              f" | items {acc['n_tc_items']} surv/q {acc['n_tc_survivors'] / nq:.1f} {'OK' if ok else 'MISMATCH'}", flush=True)
os._exit(0)

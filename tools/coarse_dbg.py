import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi
n, d, nq, k = 1_000_000, 128, 10_000, 10
rng = np.random.default_rng(42)
xb = rng.standard_normal((n, d)).astype(np.float32); xq = rng.standard_normal((nq, d)).astype(np.float32)
ix = _ffi.Index(d, 0).build(xb, seed=42, nlist=1024)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
d_xq = torch.from_numpy(xq).cuda(); d_D = torch.empty((nq, k), device='cuda'); d_I = torch.empty((nq, k), dtype=torch.int64, device='cuda')
for cm in (1, 2, 2, 1, 2):
    ix.set_coarse_mode(cm); ix.set_profiling(True)
    for it in range(4):
        ix.search_device(d_xq.data_ptr(), nq, k, 8, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
        s = ix.stats(); print(cm, it, {kk: round(s[kk], 3) for kk in s if kk.startswith('ms_')}, s['kernel_launches'], flush=True)
    ix.set_profiling(False)
os._exit(0)

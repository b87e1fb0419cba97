"""Launch list helper: searches with the bounds pass (scan mode 3) and with the seeding pass (2) on the hbm_regime index (run under ncu)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi
n, d, nq, k, nlist = 1_000_000, 128, 10_000, 10, 1024
rng = np.random.default_rng(42)
xb = rng.standard_normal((n, d)).astype(np.float32); xq = rng.standard_normal((nq, d)).astype(np.float32)
cents = xb[np.random.default_rng(7).choice(n, nlist, replace=False)].copy()
labels = _ffi.assign_points(xb, cents)
ix = _ffi.Index(d, 0).build_from_labels(xb, cents, labels)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
d_xq = torch.from_numpy(xq).cuda(); d_D = torch.empty((nq, k), device='cuda'); d_I = torch.empty((nq, k), dtype=torch.int64, device='cuda')
rt = torch.cuda.cudart()
for mode, npb in ((3, 1), (3, 4), (2, 1)):
    ix.set_scan_mode(mode)
    for it in range(3):
        if it == 2: rt.cudaProfilerStart()
        ix.search_device(d_xq.data_ptr(), nq, k, npb, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
        if it == 2: rt.cudaProfilerStop()
os._exit(0)

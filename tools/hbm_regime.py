"""The list scan in its HBM-bound regime: balanced lists, n_probe = 1, ~10 queries per list."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi
n, d, nq, k, nlist = 1_000_000, 128, 10_000, 10, 1024
rng = np.random.default_rng(42)
xb = rng.standard_normal((n, d)).astype(np.float32); xq = rng.standard_normal((nq, d)).astype(np.float32)
cents = xb[np.random.default_rng(7).choice(n, nlist, replace=False)].copy()
t0 = time.time(); labels = _ffi.assign_points(xb, cents); print('assign', time.time() - t0, 'list sizes min/med/max', np.bincount(labels, minlength=nlist).min(), np.median(np.bincount(labels, minlength=nlist)), np.bincount(labels, minlength=nlist).max())
ix = _ffi.Index(d, 0).build_from_labels(xb, cents, labels)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
d_xq = torch.from_numpy(xq).cuda(); d_D = torch.empty((nq, k), device='cuda'); d_I = torch.empty((nq, k), dtype=torch.int64, device='cuda')
ix.set_profiling(True)
for mode, npb in ((2, 1), (3, 1), (2, 4), (3, 4), (2, 16), (3, 16), (2, 64), (3, 64)):
    ix.set_scan_mode(mode)
    for _ in range(3):
        ix.search_device(d_xq.data_ptr(), nq, k, npb, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
    s = ix.stats()
    t = s['ms_scan_tc'] / 1e3
    print('mode', mode, 'nprobe', npb, 'submin slots', s['n_tc_submin_slots'], {kk: round(s[kk], 4) for kk in s if kk.startswith('ms_')}, 'alg GB', s['scan_bytes_algorithmic'] / 1e9, 'GB/s', s['scan_bytes_algorithmic'] / t / 1e9, 'pairs', s['n_pairs'], 'surv', s['n_tc_survivors'], 'ovf', s['n_tc_overflow'], 'items', s['n_tc_items'])
ix.set_scan_mode(1)
for npb in (1,):
    for _ in range(3):
        ix.search_device(d_xq.data_ptr(), nq, k, npb, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
    s = ix.stats()
    t = s['ms_scan'] / 1e3
    print('exact', npb, {kk: round(s[kk], 4) for kk in s if kk.startswith('ms_')}, 'alg GB', s['scan_bytes_algorithmic'] / 1e9, 'GB/s', s['scan_bytes_algorithmic'] / t / 1e9, 'pairs', s['n_pairs'], 'dense', s['n_dense_items'], 'sparse', s['n_sparse_items'])
os._exit(0)

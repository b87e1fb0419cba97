"""A/B sweep of the tensor-core scan's tuning knobs (environment variables read per search) on one GPU:
work-item sizing (VIDX_ITEMS_PER_SM, VIDX_MIN_CHUNK_TILES), CTA-local top-k sets (VIDX_TC_FLAGS bit 0), the bounds-pass
threshold (VIDX_BOUNDS_MAX_TILES).  Cases: the bench workload at nq = 10 000 and nq = 128, the per-rank share of a 4- and
8-way partition (mask mode), and a balanced-list index (the HBM-bound regime of SURVEY 8d).  Every setting's answer is
compared with the default's."""
import itertools, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi

n, d, nq, k = 1_000_000, 128, 10_000, 10
rng = np.random.default_rng(42)
xb = rng.standard_normal((n, d)).astype(np.float32); xq = rng.standard_normal((nq, d)).astype(np.float32)
ix = _ffi.Index(d, 0).build(xb, seed=42, nlist=1024)
cents = xb[np.random.default_rng(7).choice(n, 1024, replace=False)].copy()
bal = _ffi.Index(d, 0).build_from_labels(xb, cents, _ffi.assign_points(xb, cents))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
d_xq = torch.from_numpy(xq).cuda(); d_D = torch.empty((nq, k), device='cuda'); d_I = torch.empty((nq, k), dtype=torch.int64, device='cuda')

def run(index, nqq, npb, reps=5):
    index.set_profiling(True)
    acc = {}
    for it in range(2 + reps):
        index.search_device(d_xq.data_ptr(), nqq, k, npb, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
        if it >= 2:
            s = index.stats()
            for kk in s: acc[kk] = acc.get(kk, 0) + (s[kk] / reps if kk.startswith('ms_') else 0) if kk.startswith('ms_') else s[kk]
    index.set_profiling(False)
    return acc, d_D[:nqq].cpu().numpy().copy(), d_I[:nqq].cpu().numpy().copy()

cases = [("c2 nq=10000 np=8", ix, 10000, 8, None), ("c2 nq=128 np=8", ix, 128, 8, None), ("c2 nq=512 np=8", ix, 512, 8, None),
         ("c2 nq=2048 np=8", ix, 2048, 8, None), ("c2 nq=5000 np=8", ix, 5000, 8, None),
         ("c2 nq=10000 np=8 rank0/2", ix, 10000, 8, (0, 2)), ("c2 nq=10000 np=8 rank0/4", ix, 10000, 8, (0, 4)),
         ("c2 nq=10000 np=8 rank0/8", ix, 10000, 8, (0, 8)),
         ("bal nq=10000 np=1", bal, 10000, 1, None), ("bal nq=10000 np=8", bal, 10000, 8, None), ("bal nq=10000 np=32", bal, 10000, 32, None),
         ("bal nq=1024 np=8", bal, 1024, 8, None), ("bal nq=128 np=8", bal, 128, 8, None)]
settings = [dict(), dict(VIDX_TC_FLAGS=0), dict(VIDX_TC_FLAGS=1), dict(VIDX_ITEMS_PER_SM=4), dict(VIDX_ITEMS_PER_SM=2),
            dict(VIDX_BOUNDS_MAX_TILES=0), dict(VIDX_BOUNDS_MAX_TILES=0, VIDX_ITEMS_PER_SM=4), dict(VIDX_BOUNDS_MAX_TILES=1 << 30),
            dict(scan_mode=3), dict(scan_mode=1)]
KEYS = ("VIDX_TC_FLAGS", "VIDX_ITEMS_PER_SM", "VIDX_MIN_CHUNK_TILES", "VIDX_BOUNDS_MAX_TILES")
for name, index, nqq, npb, part in cases:
    if part: index.set_partition(*part)
    base = None
    print(f"== {name}")
    for st in settings:
        for kk in KEYS: os.environ.pop(kk, None)
        for kk, v in st.items():
            if kk != "scan_mode": os.environ[kk] = str(v)
        index.set_scan_mode(st.get("scan_mode", 0))
        s, D, I = run(index, nqq, npb)
        index.set_scan_mode(0)
        if base is None: base = (D, I)
        ok = np.array_equal(D.view(np.uint32), base[0].view(np.uint32)) and np.array_equal(I, base[1])
        t = (s['ms_scan_tc'] if s['n_tc_items'] else s['ms_scan']) / 1e3
        print(f"  {str(st):70s} scan {t * 1e3:.4f} total {s['ms_total']:.4f} group {s['ms_group']:.3f} merge {s['ms_merge']:.3f} items {s['n_tc_items']:6d} "
              f"surv/q {s['n_tc_survivors'] / nqq:7.1f} ovf {s['n_tc_overflow']} hbmfrac {s['scan_bytes_algorithmic'] / t / 1e9 / 6555.2:.3f} "
              f"tcfrac {s['tc_mma_flops'] / t / 1e12 / 1642.6:.3f} {'OK' if ok else 'MISMATCH'}", flush=True)
    for kk in KEYS: os.environ.pop(kk, None)
    if part: index.set_partition(0, 1)
os._exit(0)

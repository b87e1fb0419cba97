"""A variant of the scan kernel (CTA pair: VIDX_TC_PAIR, default; query tile in TMEM: VIDX_TC_TSA as argv[2]) against the
default kernel: identical answers, timings.  usage: pair_check.py [small|full] [VIDX_TC_PAIR|VIDX_TC_TSA]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi

def run(ix, xq, k, npb, reps=5):
    ix.set_profiling(True)
    acc = {}
    D = I = None
    for it in range(2 + reps):
        D, I = ix.search(xq, k, npb)
        if it >= 2:
            s = ix.stats()
            for kk in s: acc[kk] = acc.get(kk, 0) + s[kk] / reps if kk.startswith('ms_') else s[kk]
    ix.set_profiling(False)
    return acc, D, I

small = len(sys.argv) > 1 and sys.argv[1] == "small"
VAR = sys.argv[2] if len(sys.argv) > 2 else "VIDX_TC_PAIR"   # or VIDX_TC_TSA: the query tile in tensor memory
cases = [(20000, 64, 600, 10, 6, 24), (30000, 128, 2000, 10, 8, 16), (30000, 200, 700, 17, 8, 16), (20000, 30, 500, 32, 5, 12), (40000, 256, 900, 8, 4, 8)]
rng = np.random.default_rng(1)
for n, d, nq, k, npb, nl in cases:
    xb = rng.standard_normal((n, d)).astype(np.float32); xq = rng.standard_normal((nq, d)).astype(np.float32)
    cents = xb[rng.choice(n, nl, replace=False)].copy()
    dd = (xb * xb).sum(1)[:, None] - 2.0 * xb @ cents.T + (cents * cents).sum(1)[None, :]
    ix = _ffi.Index(d, 0).build_from_labels(xb, cents, dd.argmin(1).astype(np.uint64))
    os.environ[VAR] = "0"; s0, D0, I0 = run(ix, xq, k, npb, 2)
    os.environ[VAR] = "1"; s1, D1, I1 = run(ix, xq, k, npb, 2)
    ok = np.array_equal(D0.view(np.uint32), D1.view(np.uint32)) and np.array_equal(I0, I1)
    print(f"n={n} d={d} nq={nq} k={k} np={npb}: single {s0['ms_scan_tc']:.4f} ms ({s0['n_tc_items']} items, {s0['n_tc_survivors']} surv)  "
          f"pair {s1['ms_scan_tc']:.4f} ms ({s1['n_tc_items']} items, {s1['n_tc_survivors']} surv)  {'OK' if ok else 'MISMATCH'}", flush=True)
if not small:
    n, d, nq, k = 1_000_000, 128, 10_000, 10
    rng = np.random.default_rng(42)
    xb = rng.standard_normal((n, d)).astype(np.float32); xq = rng.standard_normal((nq, d)).astype(np.float32)
    ix = _ffi.Index(d, 0).build(xb, seed=42, nlist=1024)
    for part in (None, (0, 2), (0, 8)):
        if part: ix.set_partition(*part)
        for nqq in (10000, 2048):
            os.environ[VAR] = "0"; s0, D0, I0 = run(ix, xq[:nqq], k, 8)
            os.environ[VAR] = "1"; s1, D1, I1 = run(ix, xq[:nqq], k, 8)
            ok = np.array_equal(D0.view(np.uint32), D1.view(np.uint32)) and np.array_equal(I0, I1)
            print(f"c2 part={part} nq={nqq}: single scan {s0['ms_scan_tc']:.4f} total {s0['ms_total']:.4f} surv/q {s0['n_tc_survivors'] / nqq:.1f} | "
                  f"pair scan {s1['ms_scan_tc']:.4f} total {s1['ms_total']:.4f} surv/q {s1['n_tc_survivors'] / nqq:.1f} tcfrac {s1['tc_mma_flops'] / s1['ms_scan_tc'] / 1e9 / 1642.6:.3f}  {'OK' if ok else 'MISMATCH'}", flush=True)
        if part: ix.set_partition(0, 1)
os._exit(0)

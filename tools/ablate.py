"""Timing ablations of scan_tc_kernel on the bench workload (answers are wrong by construction): VIDX_TC_FLAGS bit 1 = the
epilogue drains nothing, bit 2 = no MMAs are issued.  Tells which role bounds the tile rate.  With a -DVIDX_TC_ABLATE build
(tools/build_variant.sh ablate -DVIDX_TC_ABLATE; VIDX_B200_LIB selects it) bits 12-18 take the rare path apart: 4096 never taken,
8192 loads only, 16384 hits found but not queued, 32768 the selector drops everything; 65536 / 131072 / 262144 keep the answers
valid (no item-end union, idle selector sleeps, adopts every eighth poll).  VIDX_TC_NB=1 runs the eight-MMA main pass."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi
n, d, nq, k = 1_000_000, 128, int(os.environ.get('NQ', 10_000)), 10
variants = [int(v) for v in os.environ.get('VARIANTS', '0,1,2').split(',')]
rng = np.random.default_rng(42)
xb = rng.standard_normal((n, d)).astype(np.float32); xq = rng.standard_normal((nq, d)).astype(np.float32)
ix = _ffi.Index(d, 0).build(xb, seed=42, nlist=1024)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
d_xq = torch.from_numpy(xq).cuda(); d_D = torch.empty((nq, k), device='cuda'); d_I = torch.empty((nq, k), dtype=torch.int64, device='cuda')
for pair in variants:   # 0 = default kernel, 1 = CTA pair, 2 = query tile in tensor memory
    for fl in [int(v) for v in os.environ.get('FLAGS', '0,8,2,6').split(',')]:
        os.environ["VIDX_TC_PAIR"] = str(int(pair == 1)); os.environ["VIDX_TC_TSA"] = str(int(pair == 2)); os.environ["VIDX_TC_FLAGS"] = str(fl)
        ix.set_profiling(True)
        acc = 0
        for it in range(5):
            ix.search_device(d_xq.data_ptr(), nq, k, 8, d_D.data_ptr(), d_I.data_ptr(), ts.cuda_stream); torch.cuda.synchronize()
            if it >= 2: acc += ix.stats()['ms_scan_tc'] / 3
        ix.set_profiling(False)
        print(f"variant={('default', 'cta pair', 'A in TMEM')[pair]} flags={fl} ({ {0: 'full', 1: 'CTA-local sets', 2: 'no epilogue', 4: 'no MMA', 6: 'neither', 8: 'independent producer', 4096: 'rare path never taken', 8192: 'tcgen05.ld + wait only', 16384: 'hits found, not queued', 32768: 'selector drops everything', 65536: 'no item-end union', 131072: 'idle selector sleeps', 262144: 'adopts every 8th poll'}.get(fl, '?') }): scan_tc {acc:.4f} ms", flush=True)
os._exit(0)

"""Coarse quantization at the BASELINE configs[4] shape (nlist = 65 536, D = 96, nq = 10 000, n_probe = 32) inside a
cudaProfilerStart/Stop window: one coarse-only call with the exact FP32 kernel, one with the tensor-core filter."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np, torch
from vector_indexer_py import _ffi
n, d, nq, nlist, npb = 2_000_000, 96, 10_000, 65536, 32
rng = np.random.default_rng(42)
xb = rng.standard_normal((n, d)).astype(np.float32); xq = rng.standard_normal((nq, d)).astype(np.float32)
cents = xb[:nlist].copy()
labels = np.arange(n, dtype=np.uint64) % nlist   # every list non-empty; the coarse stage does not look at the lists
ix = _ffi.Index(d, 0).build_from_labels(xb, cents, labels)
rt = torch.cuda.cudart()
for cm in (1, 2):
    ix.set_coarse_mode(cm)
    for it in range(3):
        if it == 2: rt.cudaProfilerStart()
        lists, dists = ix.coarse_probes(xq, npb)
        torch.cuda.synchronize()
        if it == 2: rt.cudaProfilerStop()
    print(cm, lists[:2, :4], flush=True)
os._exit(0)

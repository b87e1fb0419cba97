"""BASELINE configs[2]: mini-batch k-means 1M x 128, k = 4096 with hierarchical assignment -- GPU build (k-means++ sampled, 20 mini-batch
iterations, final hierarchical assignment, list build) vs the oracle port on the host cores, with bit-equality of centroids and labels."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np
import oracle as O
from vector_indexer_py import _ffi
n, d, k = 1_000_000, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
xb = np.random.default_rng(42).standard_normal((n, d)).astype(np.float32)
_ffi.Index(d, 0).build(xb[:20000], seed=42)  # warm up the context
t0 = time.perf_counter(); ix = _ffi.Index(d, 0).build(xb, seed=42, nlist=k); t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); res = _ffi.kmeans_mini_batch(xb, k, 20, seed=42); t_km = time.perf_counter() - t0
t0 = time.perf_counter(); oix = O.Ivf.fit(xb, seed=42, nlist=k); t_cpu = time.perf_counter() - t0
same_c = np.array_equal(ix.train_centroids().view(np.uint32), oix.centroids_all().view(np.uint32))
same_l = np.array_equal(np.asarray(ix.train_labels()), np.asarray(oix.labels_all(n)))
print(f"k={k}: GPU build (train + add + upload) {t_gpu:.3f} s; k-means alone {t_km:.3f} s; oracle port fit on {O.num_threads()} host threads {t_cpu:.2f} s; "
      f"centroids bit-equal {same_c}; labels equal {same_l}; nlist non-empty {ix.nlist}")
os._exit(0)

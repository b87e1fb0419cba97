"""configs[2] k-means (1M x 128, k = 4096, 20 mini-batch iterations) five times in a row on one box: wall time and the
host-serial / blocked-on-device / other split of vidx_kmeans_last_profile (the first run pays context and allocator set-up)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
import numpy as np
from vector_indexer_py import _ffi
xb = np.random.default_rng(42).standard_normal((1_000_000, 128)).astype(np.float32)
for rep in range(5):
    _ffi.kmeans_last_profile()
    t0 = time.perf_counter(); c, labels, iters = _ffi.kmeans_mini_batch(xb, 4096, 20, seed=42); t = time.perf_counter() - t0
    rng_s, wait_s = _ffi.kmeans_last_profile()
    print(f"run {rep}: {t:.3f} s = host-serial random stream {rng_s:.3f} + blocked on device {wait_s:.3f} + other {t - rng_s - wait_s:.3f}; iterations {iters}", flush=True)
os._exit(0)

"""
vector_indexer_py -- drop-in for the reference's Python package of the same name
(bindings/python/python/vector_indexer_py/__init__.py:23-133), backed by the B200
library libvidx_b200.so instead of the Rust extension.

Same names, argument meaning and error behaviour:
- build(xb, work_dir=None) -> VectorIndex      one-shot index build from a numpy array
- load(index_dir, shards_dir, dimension)       load index.bin + shard files into HBM
- suggest_nlist(n)                              the nlist the build would pick for n vectors
- VectorIndex.search (async) / search_sync      (D, I) = squared-L2 distances f32[nq,k] padded
                                                with +inf, external ids i64[nq,k] padded with -1
Errors surface as RuntimeError, as the PyO3 layer raises (bindings/python/src/lib.rs:134-137).

Differences, all additive: build() accepts nlist / max_iters / seed / device keywords
(the BASELINE configurations fix nlist; the reference hard-codes seed 42, src/api.rs:143);
with work_dir=None nothing is written to disk -- the index lives in HBM -- while a given
work_dir receives the reference's own file formats (index/index.bin, shards/shard_*.bin).
"""
import asyncio
import os
from typing import Optional, Tuple

import numpy as np
from numpy.typing import NDArray

from . import _ffi
from ._ffi import VidxError

__all__ = ["build", "load", "suggest_nlist", "VectorIndex"]


class PyVectorIndex:
    """The native handle (PyVectorIndex in bindings/python/src/lib.rs:48-52)."""

    def __init__(self, index: _ffi.Index, dimension: int):
        self._index = index
        self._dimension = dimension

    @property
    def dimension(self) -> int:
        return self._dimension

    def search_blocking(self, xq, k: int, n_probe: int):
        # bindings/python/src/lib.rs:123-203
        xq = np.asarray(xq)
        if xq.ndim != 2:
            raise RuntimeError("Query array must be 2-dimensional")
        if xq.shape[1] != self._dimension:
            raise RuntimeError(f"Query dimension {xq.shape[1]} doesn't match index dimension {self._dimension}")
        if xq.dtype != np.float32 or not xq.flags["C_CONTIGUOUS"]:
            raise RuntimeError("Query array must be contiguous")
        try:
            return self._index.search(xq, k, n_probe)
        except VidxError as e:
            raise RuntimeError(str(e)) from e


class VectorIndex:
    """What build() and load() return: the index resident in HBM, searchable from asyncio code (`await index.search`)
    or directly (`index.search_sync`).  Mirrors the reference class of the same name
    (bindings/python/python/vector_indexer_py/__init__.py:26-94)."""

    def __init__(self, native_index: PyVectorIndex):
        self._native = native_index

    @property
    def dimension(self) -> int:
        """Vector dimension the index was built / loaded with."""
        return self._native.dimension

    async def search(self, xq: NDArray[np.float32], k: int, n_probe: int) -> Tuple[NDArray[np.float32], NDArray[np.int64]]:
        """(D, I) for the whole batch: squared L2 distances f32[nq, k] and external ids i64[nq, k]; the blocking GPU
        call runs in the loop's default executor so the event loop stays free."""
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        loop = asyncio.get_event_loop()
        D, I = await loop.run_in_executor(None, self._native.search_blocking, xq, k, n_probe)
        return D, I

    def search_sync(self, xq: NDArray[np.float32], k: int, n_probe: int) -> Tuple[NDArray[np.float32], NDArray[np.int64]]:
        """The same search without an event loop: blocks until (D, I) are back on the host."""
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        return self._native.search_blocking(xq, k, n_probe)


def build(xb: NDArray[np.float32], work_dir: Optional[str] = None, *, nlist: int = 0, max_iters: int = 0, seed: int = 42,
          device: int = 0) -> VectorIndex:
    """
    Build an index from a numpy array of vectors (external_id = row index, timestamp = now;
    bindings/python/src/lib.rs:220-280).
    """
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    if xb.ndim != 2 or xb.shape[0] == 0:
        raise RuntimeError("Cannot build index from empty array")
    try:
        ix = _ffi.Index(xb.shape[1], device)
        ix.build(xb, None, None, seed=seed, nlist=nlist, max_iters=max_iters)
        if work_dir is not None:
            ix.save(os.path.join(work_dir, "index"), os.path.join(work_dir, "shards"))
    except VidxError as e:
        raise RuntimeError(f"Failed to build index: {e}") from e
    return VectorIndex(PyVectorIndex(ix, xb.shape[1]))


def load(index_dir: str, shards_dir: str, dimension: int, *, device: int = 0) -> VectorIndex:
    """Load an existing index from disk (bindings/python/src/lib.rs:291-304)."""
    try:
        ix = _ffi.Index(dimension, device).load(index_dir, shards_dir)
    except VidxError as e:
        raise RuntimeError(f"Failed to load index: {e}") from e
    # the dimension stored in index.bin wins (the reference never compares it with the argument, src/api.rs:109-112;
    # validating queries against the caller's number would let a wider file dimension read past the query buffer)
    return VectorIndex(PyVectorIndex(ix, ix.dimension))


def suggest_nlist(n: int) -> int:
    """calculate_num_clusters (src/utils.rs:9-16; bindings/python/src/lib.rs:308-315)."""
    return _ffi.calculate_num_clusters(n)

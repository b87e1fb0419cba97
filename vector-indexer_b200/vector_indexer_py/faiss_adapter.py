"""Faiss-style adapter surface of the reference's benchmark harness
(bench/faiss_bench_official/vector_indexer_adapter.py:75-140): `.d`, an `nprobe` attribute and a synchronous
`.search(xq, k) -> (D, I)`, so sweeps written against the reference's harness (bench_all_ivf.py:283-363:
`index.nprobe = p; D, I = index.search(xq, k)`) run unchanged on the B200 library.  The reference needs a dedicated
asyncio thread because its search is an async Rust future; here the batch goes to the GPU in one call."""
from typing import Tuple

import numpy as np


class VectorIndexerFaissAdapter:
    def __init__(self, vector_index, k: int = 100):
        self._idx = vector_index
        self._k = k
        self._nprobe = 1

    @property
    def d(self) -> int:
        return self._idx.dimension

    @property
    def nprobe(self) -> int:
        return self._nprobe

    @nprobe.setter
    def nprobe(self, value: int):
        self._nprobe = int(value)

    def search(self, xq: np.ndarray, k: int = None) -> Tuple[np.ndarray, np.ndarray]:
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        return self._idx.search_sync(xq, self._k if k is None else k, self._nprobe)

    def __repr__(self):
        return f"VectorIndexerFaissAdapter(d={self.d}, nprobe={self.nprobe})"


def recall_at_ranks(I: np.ndarray, gt: np.ndarray, ranks=(1, 10, 100)):
    """The harness's R@r (bench_all_ivf.py:336-350): fraction of queries whose true nearest neighbour is among the
    first r results."""
    nq = I.shape[0]
    return {r: float((I[:, :r] == gt[:, :1]).sum()) / nq for r in ranks if r <= I.shape[1]}

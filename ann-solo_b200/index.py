"""The slice of the Faiss API the reference uses (SURVEY.md §8b "Faiss surface used"), backed by
the device IVF index behind the C-ABI: IndexFlatIP(d), IndexIVFFlat(quantizer, d, nlist,
METRIC_INNER_PRODUCT) with train / add / search / reset / nprobe / ntotal, plus set_centroids
so parity runs can share centroids with the oracle.
"""
from __future__ import annotations

import itertools

import numpy as np

from .spectrum import default_engine

METRIC_INNER_PRODUCT = 0
_slots = itertools.count(1000)  # private "charge" keys of the engine's index table


class IndexFlatIP:
    def __init__(self, d: int):
        self.d = d


class IndexIVFFlat:
    """reference spectral_library.py:167-181, :443-444, :497."""

    def __init__(self, quantizer, d: int, nlist: int, metric=METRIC_INNER_PRODUCT, engine=None, slot=None):
        if metric != METRIC_INNER_PRODUCT:
            raise ValueError("only METRIC_INNER_PRODUCT is implemented (the reference uses no other)")
        self.d, self.nlist = int(d), int(nlist)
        self.nprobe = 1  # Faiss default
        self.is_trained = False
        self._eng = engine or default_engine()
        self._slot = next(_slots) if slot is None else slot
        self.train_iters = 10

    @property
    def ntotal(self) -> int:
        return self._eng.ivf_info(self._slot)[0] if self.is_trained else 0

    def set_centroids(self, centroids: np.ndarray):
        centroids = np.ascontiguousarray(centroids, np.float32)
        if centroids.shape != (self.nlist, self.d):
            raise ValueError("centroids must be (nlist, d)")
        self._eng.ivf_set_centroids(self._slot, centroids)
        self.is_trained = True

    def train(self, x: np.ndarray):
        self._eng.ivf_train(self._slot, x, self.nlist, self.train_iters)
        self.is_trained = True

    def add(self, x: np.ndarray):
        if not self.is_trained:
            raise RuntimeError("index is not trained")
        self._eng.ivf_add(self._slot, x)

    def reset(self):
        if self.is_trained:
            self._eng.ivf_reset(self._slot)

    def search(self, x: np.ndarray, k: int):
        return self._eng.ivf_search(self._slot, x, k, self.nprobe)

    def setNumProbes(self, nprobe: int):  # GpuIndexIVF spelling used at spectral_library.py:495
        self.nprobe = int(nprobe)

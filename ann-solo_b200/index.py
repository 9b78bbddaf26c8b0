"""The slice of the Faiss API the reference uses (SURVEY.md §8b "Faiss surface used"), backed by
the device IVF index behind the C-ABI: IndexFlatIP(d), IndexIVFFlat(quantizer, d, nlist,
METRIC_INNER_PRODUCT) with train / add / search / reset / nprobe / ntotal, plus set_centroids
so parity runs can share centroids with the oracle.
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import _lib
from .spectrum import default_engine

METRIC_INNER_PRODUCT = 0
_slots = itertools.count(1000)  # private "charge" keys of the engine's index table


class IndexFlatIP:
    def __init__(self, d: int):
        self.d = d


class IndexIVF:
    """Base name the reference annotates with (spectral_library.py:457)."""


class IndexIVFFlat(IndexIVF):
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

    def reconstruct_n(self, i0: int = 0, ni: int = None) -> np.ndarray:
        return self._eng.ivf_reconstruct(self._slot, i0, ni)


def get_num_gpus() -> int:
    """faiss.get_num_gpus (reference spectral_library.py:71). Answers 0 on purpose: every index of this module
    already lives on the device, so the reference takes its plain branch (`index.nprobe = n`, :497) instead of
    cloning to a GpuIndex and clamping num_probe / num_candidates to Faiss-GPU's 1024 (:76-86)."""
    return 0


class StandardGpuResources:   # accepted for callers that force the reference's GPU branch; nothing to hold
    pass


class GpuClonerOptions:
    useFloat16 = False


def index_cpu_to_gpu(res, device: int, index: "IndexIVFFlat", options=None) -> "IndexIVFFlat":
    """faiss.index_cpu_to_gpu (reference spectral_library.py:494): the index is on the device already."""
    return index


def write_index(index: IndexIVFFlat, fname: str) -> None:
    """faiss.write_index (reference spectral_library.py:181): Faiss' IndexIVFFlat byte layout."""
    if not index.is_trained:
        raise RuntimeError("index is not trained")
    index._eng.ivf_write_index(index._slot, fname, index.nprobe)


def inspect_index(fname: str) -> dict:
    """Header and list table of a Faiss IndexIVFFlat file, parsed on the host (no GPU needed)."""
    info = _lib.IdxannInfo()
    err = C.create_string_buffer(512)
    rc = _lib.load().solo_idxann_inspect(str(fname).encode(), C.byref(info), err, len(err))
    if rc != 0:
        if rc == _lib.SOLO_EINVAL:
            raise ValueError(err.value.decode())
        raise _lib.SoloError(rc, err.value.decode())
    out = {name: getattr(info, name) for name, _ in info._fields_ if name != "reserved"}
    out["fourcc"] = info.fourcc.decode()
    out["quantizer_fourcc"] = info.quantizer_fourcc.decode()
    return out


def read_index(fname: str, engine=None, slot=None) -> IndexIVFFlat:
    """faiss.read_index (reference spectral_library.py:490): the file's centroids and inverted lists
    are taken as stored (no re-training, no re-assignment)."""
    info = inspect_index(fname)
    index = IndexIVFFlat(IndexFlatIP(info["d"]), info["d"], info["nlist"], info["metric"], engine=engine, slot=slot)
    index.nprobe = max(1, index._eng.ivf_read_index(index._slot, fname))
    index.is_trained = True
    return index

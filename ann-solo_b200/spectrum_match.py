"""Drop-in for the reference's Cython module src/ann_solo/spectrum_match.pyx: same function
name, arguments and return value; the scoring runs in the K5 CUDA kernel through the C-ABI.
"""
from __future__ import annotations

import numpy as np

from .spectrum import default_engine, spectra_to_store

_SCRATCH_CHARGE = -1  # slot of the engine's library table used for ad-hoc candidate lists


def get_best_match(query, candidates, fragment_mz_tolerance, allow_shift, engine=None):
    """Reference spectrum_match.pyx:28-108: returns ``(candidates[best], score, [(query_peak,
    candidate_peak), ...])``. An empty candidate list is a caller error in the reference (it
    dereferences NULL, pyx:97); here it raises ValueError."""
    if not candidates:
        raise ValueError("get_best_match needs at least one candidate (reference callers guard this, "
                         "spectral_library.py:359)")
    eng = engine or default_engine()
    lib = spectra_to_store(candidates, with_charge=True)
    lib["valid"][:] = 1
    eng.load_library(_SCRATCH_CHARGE, lib)
    q = spectra_to_store([query], with_charge=False)
    query.charge = np.zeros(len(query.mz), np.uint8)  # pyx:88
    ids = np.arange(len(candidates), dtype=np.int32)
    off = np.array([0, len(candidates)], np.int64)
    max_pairs = max(1, len(query.mz))
    bp, bs, npairs, pairs = eng.best_match_batch(_SCRATCH_CHARGE, q, ids, off, fragment_mz_tolerance,
                                                 bool(allow_shift), max_pairs)
    n = int(npairs[0])
    return candidates[int(bp[0])], float(bs[0]), [(int(a), int(b)) for a, b in pairs[0, :n]]

"""Host-side mirror of the reference's src/ann_solo/spectrum.py for the hot path: same names,
argument meaning and return types; the arithmetic of ``spectrum_to_vector`` runs in the K1
CUDA kernel. ``process_spectrum`` (reference :57-119) stays on the host for now — its five
peak operations come from spectrum_utils, which is absent here and are
restated from its documented behaviour.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from .config import config


class MsmsSpectrum:
    """Minimal stand-in for spectrum_utils.spectrum.MsmsSpectrum with the attributes the hot
    path reads (SURVEY.md §8b "Spectrum object contract")."""

    def __init__(self, identifier, precursor_mz, precursor_charge, mz, intensity, annotation=None,
                 retention_time=None, peptide=None, is_decoy=False):
        self.identifier = identifier
        self.precursor_mz = float(precursor_mz)
        self.precursor_charge = precursor_charge
        order = None
        mz = np.asarray(mz)
        if len(mz) > 1 and np.any(np.diff(mz) < 0):
            order = np.argsort(mz, kind="stable")
            mz = mz[order]
        self._mz = mz
        inten = np.asarray(intensity, np.float32)
        self._intensity = inten[order] if order is not None else inten
        if annotation is not None and order is not None:
            annotation = [annotation[i] for i in order]
        self._annotation = annotation
        self.retention_time = retention_time
        self.peptide = peptide
        self.is_decoy = is_decoy
        self.is_valid = True
        self.is_processed = False
        self.index = None

    @property
    def mz(self):
        return self._mz

    @property
    def intensity(self):
        return self._intensity

    @property
    def annotation(self):
        if self._annotation is None:
            return [None] * len(self._mz)
        return self._annotation


def _check_spectrum_valid(spectrum_mz, min_peaks, min_mz_range) -> bool:
    # reference spectrum.py:14-36
    return len(spectrum_mz) >= min_peaks and spectrum_mz[-1] - spectrum_mz[0] >= min_mz_range


def _select(spectrum, keep):
    ann = spectrum._annotation
    spectrum._mz = spectrum._mz[keep]
    spectrum._intensity = spectrum._intensity[keep]
    if ann is not None:
        idx = np.flatnonzero(keep) if keep.dtype == bool else keep
        spectrum._annotation = [ann[i] for i in idx]


# The five spectrum_utils (<0.4) peak operations the reference's process_spectrum calls (spectrum.py:78-107), as
# methods with spectrum_utils' names and argument order, each changing the spectrum in place and returning it.
# spectrum_utils is absent here: the bodies are restated from its documented behaviour and are
# what oracle/solo_oracle.py:process_spectrum_np and the K0 kernel compute bit for bit.
def _set_mz_range(self, min_mz, max_mz):
    _select(self, (self._mz >= min_mz) & (self._mz <= max_mz))
    return self


def _round(self, decimals: int = 0, combine: str = "sum"):
    """Round m/z to `decimals` and merge peaks that fall on the same value: intensities are combined ('sum' or
    'max'), the annotation of the most intense peak of the group is kept (first one on ties)."""
    if len(self._mz) == 0:
        return self
    r = np.round(self._mz, decimals)
    head = np.ones(len(r), bool)
    head[1:] = r[1:] != r[:-1]          # m/z is ascending: equal rounded values are adjacent
    group = np.cumsum(head) - 1
    n = int(group[-1]) + 1
    if n < len(r):
        if combine == "sum":
            merged = np.zeros(n, np.float32)
            np.add.at(merged, group, self._intensity)     # float32, in peak order
        elif combine == "max":
            merged = np.full(n, -np.inf, np.float32)
            np.maximum.at(merged, group, self._intensity)
        else:
            raise ValueError("Unknown method to combine peak intensities")
        ann = self._annotation
        if ann is not None:
            best = np.zeros(n, np.int64)
            top = np.full(n, -np.inf)
            for i, g in enumerate(group):                   # first most intense peak of each group
                if self._intensity[i] > top[g]:
                    top[g], best[g] = self._intensity[i], i
            self._annotation = [ann[i] for i in best]
        self._mz, self._intensity = r[head], merged
    else:
        self._mz = r
    return self


def _remove_precursor_peak(self, fragment_tol_mass, fragment_tol_mode, isotope: int = 0):
    z = self.precursor_charge
    neutral = (self.precursor_mz - 1.0072766) * z
    rm = np.zeros(len(self._mz), bool)
    for c in range(z, 0, -1):
        for iso in range(isotope + 1):
            target = (neutral + iso) / c + 1.0072766
            if fragment_tol_mode == "Da":
                rm |= np.abs(self._mz - target) <= fragment_tol_mass
            elif fragment_tol_mode == "ppm":
                rm |= np.abs(self._mz - target) / target * 1e6 <= fragment_tol_mass
            else:
                raise ValueError("Unknown fragment tolerance mode")
    _select(self, ~rm)
    return self


def _filter_intensity(self, min_intensity: float = 0.0, max_num_peaks=None):
    inten = self._intensity
    if max_num_peaks is None:
        max_num_peaks = len(inten)
    order = np.argsort(inten, kind="stable")
    thr = min_intensity * (inten[order[-1]] if len(order) else 0.0)
    start = int(np.searchsorted(inten[order], thr, side="right"))
    sel = order[max(start, len(order) - max_num_peaks):]
    keep = np.zeros(len(inten), bool)
    keep[sel] = True
    _select(self, keep)
    return self


def _scale_intensity(self, scaling=None, max_intensity=None, degree: int = 2, base: int = 2, max_rank=None):
    if scaling == "root":
        self._intensity = np.power(self._intensity, 1 / degree).astype(np.float32) if degree != 2 else \
            np.sqrt(self._intensity).astype(np.float32)
    elif scaling == "log":
        self._intensity = (np.log1p(self._intensity) / np.log(base)).astype(np.float32)
    elif scaling == "rank":
        if max_rank is None:
            max_rank = len(self._intensity)
        if max_rank < len(self._intensity):
            raise ValueError("`max_rank` should be greater than or equal to the number of peaks in the spectrum")
        inten = self._intensity
        self._intensity = (max_rank - np.argsort(np.argsort(inten, kind="stable")[::-1], kind="stable")
                           ).astype(np.float32)
    elif scaling is not None:
        raise ValueError("Unknown intensity scaling")
    if max_intensity is not None:
        self._intensity = (self._intensity * max_intensity / self._intensity.max()).astype(np.float32)
    return self


MsmsSpectrum.set_mz_range = _set_mz_range
MsmsSpectrum.round = _round
MsmsSpectrum.remove_precursor_peak = _remove_precursor_peak
MsmsSpectrum.filter_intensity = _filter_intensity
MsmsSpectrum.scale_intensity = _scale_intensity


def process_spectrum(spectrum, is_library: bool):
    """Reference spectrum.py:57-119, same order of operations and validity checks."""
    if spectrum.is_processed:
        return spectrum
    min_peaks, min_mz_range = config.min_peaks, config.min_mz_range

    def invalid():
        spectrum.is_valid = False
        spectrum.is_processed = True
        return spectrum

    spectrum.set_mz_range(config.min_mz, config.max_mz)
    if not _check_spectrum_valid(spectrum._mz, min_peaks, min_mz_range):
        return invalid()
    if config.resolution is not None:
        spectrum.round(config.resolution, "sum")
        if not _check_spectrum_valid(spectrum._mz, min_peaks, min_mz_range):
            return invalid()
    if config.remove_precursor:
        spectrum.remove_precursor_peak(config.remove_precursor_tolerance, "Da", 2)
        if not _check_spectrum_valid(spectrum._mz, min_peaks, min_mz_range):
            return invalid()
    max_peaks = config.max_peaks_used_library if is_library else config.max_peaks_used
    spectrum.filter_intensity(config.min_intensity, max_peaks)
    if not _check_spectrum_valid(spectrum._mz, min_peaks, min_mz_range):
        return invalid()
    scaling = config.scaling
    if scaling == "sqrt":
        scaling = "root"
    if scaling is not None:
        spectrum.scale_intensity(scaling, max_rank=max_peaks)
    nrm = np.float32(np.sqrt(np.sum(spectrum._intensity.astype(np.float64) ** 2)))
    spectrum._intensity = (spectrum._intensity / nrm).astype(np.float32)  # _norm_intensity, spectrum.py:39-54
    spectrum.is_valid = True
    spectrum.is_processed = True
    return spectrum


def get_dim(min_mz, max_mz, bin_size):
    """Reference spectrum.py:123-143."""
    min_mz, max_mz = float(min_mz), float(max_mz)
    start_dim = min_mz - min_mz % bin_size
    end_dim = max_mz + bin_size - max_mz % bin_size
    return round((end_dim - start_dim) / bin_size), start_dim, end_dim


_default_engine = None


def default_engine():
    """Process-wide engine on cuda:0 for the single-call API (spectrum_to_vector, get_best_match)."""
    global _default_engine
    if _default_engine is None:
        from .engine import SoloEngine
        _default_engine = SoloEngine(0)
    return _default_engine


def hash_idx(bin_idx: int, hash_len: int) -> int:
    """Reference spectrum.py:147-163, answered from the device LUT the kernels use."""
    eng = default_engine()
    if eng.hash_len != hash_len:
        eng.set_vectorizer(config.min_mz, config.max_mz, config.bin_size, hash_len)
    return eng.hash_slot(bin_idx)


def spectrum_to_vector(spectrum, min_mz: float, max_mz: float, bin_size: float, hash_len: int, norm: bool = True,
                       vector: Optional[np.ndarray] = None, engine=None) -> np.ndarray:
    """Reference spectrum.py:166-214, one spectrum through the batched K1 kernel."""
    if hash_len is None:
        raise ValueError("the CUDA vectoriser implements the hashed form only (hash_len must be set)")
    eng = engine or default_engine()
    eng.set_vectorizer(min_mz, max_mz, bin_size, hash_len)
    if vector is not None and vector.shape[0] != hash_len:
        raise ValueError("Incorrect vector dimensionality")
    mz = np.asarray(spectrum.mz)
    if mz.dtype != np.float32:
        mz = mz.astype(np.float64)
    v = eng.vectorize(mz, np.asarray(spectrum.intensity, np.float32), np.array([0, len(mz)], np.int64), norm)[0]
    if vector is not None:
        # the reference accumulates into the caller's vector before normalising; callers pass zeros
        vector[:] = v
        return vector
    return v


class SpectrumSpectrumMatch:
    """Reference spectrum.py:217-271."""

    def __init__(self, query_spectrum, library_spectrum=None, peak_matches=None, search_engine_score=math.nan,
                 q=math.nan):
        self.query_spectrum = query_spectrum
        self.library_spectrum = library_spectrum
        self.peak_matches = peak_matches
        self.search_engine_score = search_engine_score
        self.q = q

    @property
    def sequence(self):
        return self.library_spectrum.peptide if self.library_spectrum is not None else None

    @property
    def query_identifier(self):
        return self.query_spectrum.identifier

    @property
    def query_index(self):
        return self.query_spectrum.index

    @property
    def library_identifier(self):
        return self.library_spectrum.identifier if self.library_spectrum is not None else None

    @property
    def retention_time(self):
        return self.query_spectrum.retention_time

    @property
    def charge(self):
        return self.query_spectrum.precursor_charge

    @property
    def exp_mass_to_charge(self):
        return self.query_spectrum.precursor_mz

    @property
    def calc_mass_to_charge(self):
        return self.library_spectrum.precursor_mz if self.library_spectrum is not None else None

    @property
    def is_decoy(self):
        return self.library_spectrum.is_decoy if self.library_spectrum is not None else None


def spectra_to_store(spectra, with_charge: bool = True) -> dict:
    """Flatten duck-typed spectrum objects into a peak store (CSR). Peak charges come from
    ``annotation[i].charge`` (0 when the annotation is None), as spectrum_match.pyx:74-79."""
    n = len(spectra)
    counts = np.fromiter((len(s.mz) for s in spectra), np.int64, n)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    mz = np.empty(off[-1], np.float32)
    mz64 = np.empty(off[-1], np.float64)
    inten = np.empty(off[-1], np.float32)
    chg = np.zeros(off[-1], np.uint8)
    for i, s in enumerate(spectra):
        b, e = off[i], off[i + 1]
        mz64[b:e] = s.mz
        mz[b:e] = np.asarray(s.mz).astype(np.float32)  # pyx:80 astype(np.float32)
        inten[b:e] = np.asarray(s.intensity).astype(np.float32)
        if with_charge:
            cached = getattr(s, "charge", None)
            if cached is not None and len(cached) == e - b:
                chg[b:e] = cached
            else:
                ann = getattr(s, "_annotation", None) if hasattr(s, "_annotation") else s.annotation
                if ann is not None:
                    c = np.fromiter((0 if a is None else a.charge for a in ann), np.uint8, e - b)
                    chg[b:e] = c
                    try:
                        s.charge = c  # the reference caches this on the object too (pyx:75-79)
                    except AttributeError:
                        pass
    return dict(mz=mz, mz64=mz64, inten=inten, chg=chg, off=off,
                prec_mz=np.fromiter((s.precursor_mz for s in spectra), np.float64, n),
                prec_z=np.fromiter((s.precursor_charge or 0 for s in spectra), np.int32, n),
                valid=np.fromiter((1 if getattr(s, "is_valid", True) else 0 for s in spectra), np.uint8, n))

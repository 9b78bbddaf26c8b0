"""Host-side mirror of the reference's search engine, src/ann_solo/spectral_library.py, for the
hot path: ``SpectralLibrary.search`` / ``_search_cascade`` / ``_search_batch`` /
``_get_library_candidates`` keep their names, arguments and results, but one batch of
<= config.batch_size queries is ONE fused device call (vectorise -> IVF top-k -> precursor
window -> shifted dot) instead of per-query Python/Faiss/Cython work.

Out of scope here: file readers, HDF5/.spcfg caches, mokapot rescoring. ANN indexes
are cached in Faiss' own ``.idxann`` files when ``ann_basename`` is given (reference :92-130). The library is handed over as an object with the reader surface the hot path
uses (``spec_info`` and ``read_spectrum``); FDR scoring is an injectable callable.
"""
from __future__ import annotations

import collections
import copy
import hashlib
import json
import logging
import os
from typing import Callable, Dict, Iterator, List, Optional

import numpy as np

from .config import config
from .engine import SoloEngine
from .spectrum import MsmsSpectrum, SpectrumSpectrumMatch, process_spectrum, spectra_to_store


class InMemoryLibrary:
    """The slice of SpectralLibraryReader (reference reader.py:29-437) the hot path touches:
    ``spec_info['charge'][z] = {'id', 'precursor_mz' (float32)}`` (reader.py:180-189) and
    ``read_spectrum(id, process_peaks)`` (reader.py:218-246), over a peak store of already
    processed library spectra (see synth.py)."""

    def __init__(self, store: dict, identifiers: Optional[np.ndarray] = None, peptides=None):
        self.store = store
        n = len(store["prec_mz"])
        self.identifiers = np.arange(n) if identifiers is None else np.asarray(identifiers)
        self.peptides = peptides
        self._row_of = {ident: i for i, ident in enumerate(self.identifiers.tolist())}
        self.spec_info = {"charge": {}}
        self.rows = {}
        for z in np.unique(store["prec_z"]):
            rows = np.flatnonzero(store["prec_z"] == z)
            self.rows[int(z)] = rows
            self.spec_info["charge"][int(z)] = {
                "id": self.identifiers[rows],
                "precursor_mz": store["prec_mz"][rows].astype(np.float32),  # reader.py:188-189
            }
        self.is_recreated = False

    def charge_store(self, charge: int) -> dict:
        from .synth import take_spectra
        s = take_spectra(self.store, self.rows[charge])
        s["prec_mz32"] = self.spec_info["charge"][charge]["precursor_mz"]
        return s

    def read_spectrum(self, spec_id, process_peaks: bool = False) -> MsmsSpectrum:
        r = self._row_of[spec_id]
        st = self.store
        b, e = st["off"][r], st["off"][r + 1]
        s = MsmsSpectrum(spec_id, st["prec_mz"][r], int(st["prec_z"][r]), st["mz"][b:e], st["inten"][b:e],
                         peptide=None if self.peptides is None else self.peptides[r],
                         is_decoy=bool(st["is_decoy"][r]) if "is_decoy" in st else False)
        s.charge = st["chg"][b:e]
        s.is_valid = bool(st["valid"][r])
        s.is_processed = True
        return s

    def close(self):
        pass


def tdc_score_ssms(ssms: List[SpectrumSpectrumMatch], fdr: float, model=None, grouped: bool = False):
    """Stand-in for reference utils.score_ssms (mokapot rescoring, OUT OF SCOPE): plain
    target-decoy q-values on the hot path's dot-product score."""
    ssms = sorted(ssms, key=lambda s: -s.search_engine_score)
    decoys = targets = 0
    fdrs = []
    for s in ssms:
        if s.is_decoy:
            decoys += 1
        else:
            targets += 1
        fdrs.append(decoys / max(targets, 1))
    q = np.minimum.accumulate(np.asarray(fdrs[::-1]))[::-1] if fdrs else []
    for s, qv in zip(ssms, q):
        s.q = float(qv)
    return [s for s in ssms if not s.is_decoy]


class SpectralLibrary:
    """Spectral library search engine (reference spectral_library.py:27-500) on one B200."""

    _hyperparameters = ["min_mz", "max_mz", "bin_size", "hash_len", "num_list"]

    def __init__(self, library, engine: Optional[SoloEngine] = None, device: int = 0,
                 centroids: Optional[Dict[int, np.ndarray]] = None,
                 score_ssms: Callable = tdc_score_ssms, train_iters: int = 10,
                 ann_basename: Optional[str] = None) -> None:
        """``ann_basename``: what the reference derives from the library file name
        (``os.path.splitext(filename)[0]``, :99-100); when given, the index of every charge is read
        from / written to ``{ann_basename}_{hash[:7]}_{charge}.idxann`` like the reference does."""
        self._engine = engine or SoloEngine(device)
        if isinstance(library, (str, os.PathLike)):   # the reference's signature: SpectralLibrary(filename), :46-71
            from .reader import SpectralLibraryReader
            filename = os.fspath(library)
            try:
                library = SpectralLibraryReader(filename, self._get_hyperparameter_hash(), engine=self._engine)
            except FileNotFoundError as e:
                logging.error(e)
                raise
            if ann_basename is None:
                ann_basename = os.path.splitext(filename)[0]
        self._library_reader = library
        self._score_ssms = score_ssms
        self._num_probe = config.num_probe
        self._num_candidates = config.num_candidates
        self._engine.set_vectorizer(config.min_mz, config.max_mz, config.bin_size, config.hash_len)
        self._ann_charges = set()
        self._ann_filenames = {}
        self._lib_ids = {}
        self._lib_row_of = {}
        for charge, info in self._library_reader.spec_info["charge"].items():
            self._engine.load_library(charge, self._charge_store(charge))
            self._lib_ids[charge] = info["id"]
            self._lib_row_of[charge] = {ident: i for i, ident in enumerate(np.asarray(info["id"]).tolist())}
        if config.mode == "ann":
            # No ANN index for infrequent precursor charges (reference :101-104).
            ann_charges = [z for z, info in self._library_reader.spec_info["charge"].items()
                           if len(info["id"]) >= config.num_list]
            create_ann_charges = []
            for charge in sorted(ann_charges):
                if ann_basename is None:
                    create_ann_charges.append(charge)
                    continue
                self._ann_filenames[charge] = (f"{ann_basename}_{self._get_hyperparameter_hash()[:7]}_"
                                               f"{charge}.idxann")
                if (getattr(self._library_reader, "is_recreated", False) or
                        not os.path.isfile(self._ann_filenames[charge])):
                    create_ann_charges.append(charge)
                    logging.warning("Missing ANN index for charge %d", charge)
                else:  # reference _get_ann_index :490 (loaded once: every charge stays resident in HBM)
                    self._engine.ivf_read_index(charge, self._ann_filenames[charge])
                    n_lib = len(self._library_reader.spec_info["charge"][charge]["id"])
                    n_idx = self._engine.ivf_info(charge)[0]
                    if n_idx != n_lib:
                        # a stale or foreign index (library replaced under the same name, index written with
                        # decoys added, ...): the reference rebuilds on a configuration change (reader.py
                        # is_recreated); row ids beyond the library must never reach the peak store
                        logging.warning("ANN index %s holds %d rows but charge %d of the library has %d: rebuilding",
                                        self._ann_filenames[charge], n_idx, charge, n_lib)
                        self._engine.ivf_reset(charge)
                        create_ann_charges.append(charge)
                    else:
                        self._ann_charges.add(charge)
            self._create_ann_indexes(create_ann_charges, centroids or {}, train_iters)

    def _get_hyperparameter_hash(self) -> str:
        """Reference :119-131."""
        hyperparameters_bytes = json.dumps({hp: config[hp] for hp in self._hyperparameters}).encode("utf-8")
        return hashlib.sha1(hyperparameters_bytes).hexdigest()

    def _charge_store(self, charge: int) -> dict:
        if hasattr(self._library_reader, "charge_store"):
            return self._library_reader.charge_store(charge)
        info = self._library_reader.spec_info["charge"][charge]
        spectra = [self._library_reader.read_spectrum(i, True) for i in info["id"]]
        st = spectra_to_store(spectra, with_charge=True)
        st["prec_mz32"] = np.asarray(info["precursor_mz"], np.float32)
        return st

    def _create_ann_indexes(self, charges: List[int], centroids: Dict[int, np.ndarray], train_iters: int) -> None:
        """Reference :133-183: vectorise every library spectrum of the charge, k-means, add.
        Row i of the index is position i of spec_info['charge'][charge]['id']."""
        for charge in charges:
            if charge in centroids:
                self._engine.ivf_set_centroids(charge, centroids[charge])
            else:
                st = self._charge_store(charge)
                vecs = self._engine.vectorize(st["mz"], st["inten"], st["off"], True)
                self._engine.ivf_train(charge, vecs[np.isfinite(vecs).all(axis=1)], config.num_list, train_iters)
            self._engine.ivf_add_library(charge)
            self._ann_charges.add(charge)
            if charge in self._ann_filenames:  # reference :181
                self._engine.ivf_write_index(charge, self._ann_filenames[charge])

    def compute_ssm_features(self, ssms: List[SpectrumSpectrumMatch]) -> Dict[str, List]:
        """Reference utils._compute_ssm_features (utils.py:276-457) for SSMs of this library: one K6
        launch per precursor charge against the device-resident peak stores."""
        from .utils import _compute_ssm_features
        return _compute_ssm_features(ssms, engine=self._engine, library_rows=self._lib_row_of)

    def shutdown(self) -> None:
        self._library_reader.close()
        for charge in self._ann_charges:
            self._engine.ivf_reset(charge)

    # ------------------------------------------------------------------ search
    def search(self, query_spectra) -> List[SpectrumSpectrumMatch]:
        """Reference :193-260. ``query_spectra`` is a query file name (``.mgf``, read natively) like in the
        reference, or an iterable of query spectrum objects."""
        if isinstance(query_spectra, (str, os.PathLike)):   # the reference's signature: search(query_filename), :193-215
            from .reader import read_query_file
            query_spectra = read_query_file(os.fspath(query_spectra))
        by_charge = collections.defaultdict(list)
        for query_spectrum in query_spectra:
            if query_spectrum.precursor_charge is not None:
                qsc = [query_spectrum]
            else:  # unknown charge: try 2 and 3 (:219-223)
                qsc = []
                for charge in (2, 3):
                    qsc.append(copy.copy(query_spectrum))
                    qsc[-1].precursor_charge = charge
            for qs in qsc:
                if process_spectrum(qs, False).is_valid:
                    by_charge[qs.precursor_charge].append(qs)
        identifications = {}
        do_cascade_open = (config.precursor_tolerance_mass_open is not None and
                           config.precursor_tolerance_mode_open is not None)
        n_identified = 0
        for ssm in self._search_cascade(by_charge, "std"):
            if not do_cascade_open or ssm.q < config.fdr:
                identifications[ssm.query_identifier] = ssm
                n_identified += ssm.q < config.fdr
        logging.info("%d spectra identified after the standard search", n_identified)
        if do_cascade_open:
            for charge, qs in by_charge.items():
                by_charge[charge] = [s for s in qs if s.identifier not in identifications]
            for ssm in self._search_cascade(by_charge, "open"):
                identifications[ssm.query_identifier] = ssm
                n_identified += ssm.q < config.fdr
            logging.info("%d spectra identified after the open search", n_identified)
        return list(identifications.values())

    def _search_cascade(self, query_spectra: Dict[int, List], mode: str) -> Iterator[SpectrumSpectrumMatch]:
        """Reference :262-326."""
        if mode not in ("std", "open"):
            raise ValueError("Unknown search mode")
        ssms = {}
        batch_size = config.batch_size
        for charge, qs in query_spectra.items():
            for b in range(0, len(qs), batch_size):
                for ssm in self._search_batch(qs[b:b + batch_size], charge, mode):
                    # first SSM per identifier wins (scores are NaN at this point in the
                    # reference); ours carry the dot product, keep the rule
                    if ssm is not None and ssm.query_identifier not in ssms:
                        ssms[ssm.query_identifier] = ssm
        return self._score_ssms(list(ssms.values()), config.fdr, None, mode == "open")

    def _mode_tolerance(self, mode: str):
        if mode == "std":
            return config.precursor_tolerance_mass, config.precursor_tolerance_mode
        if mode == "open":
            return config.precursor_tolerance_mass_open, config.precursor_tolerance_mode_open
        raise ValueError("Unknown search mode")

    def _use_ann(self, charge: int, mode: str) -> bool:
        return config.mode == "ann" and mode == "open" and charge in self._ann_charges  # reference :432-433

    def _search_batch(self, query_spectra: List, charge: int, mode: str) -> Iterator[SpectrumSpectrumMatch]:
        """Reference :328-370 as one fused device call."""
        tol_val, tol_mode = self._mode_tolerance(mode)
        if charge not in self._library_reader.spec_info["charge"] or not query_spectra:
            return
        q = spectra_to_store(query_spectra, with_charge=False)
        mz_vec = q["mz64"] if any(np.asarray(s.mz).dtype != np.float32 for s in query_spectra) else None
        max_pairs = max(1, int(np.diff(q["off"]).max()))
        params = SoloEngine.make_params(self._use_ann(charge, mode), self._num_candidates, self._num_probe, tol_val,
                                        tol_mode, config.fragment_mz_tolerance, config.allow_peak_shifts, max_pairs,
                                        mz_vec is not None)
        res = self._engine.search_batch(charge, params, q, mz_vec)
        ids = self._lib_ids[charge]
        for i, query_spectrum in enumerate(query_spectra):
            row = int(res["best_row"][i])
            if row < 0:  # no candidates: the reference yields nothing (:359)
                continue
            library_match = self._library_reader.read_spectrum(ids[row], True)
            n = int(res["n_pairs"][i])
            yield SpectrumSpectrumMatch(query_spectrum, library_match,
                                        peak_matches=res["pairs"][i, :n].astype(np.int64),
                                        search_engine_score=float(res["score"][i]))

    def _get_library_candidates(self, query_spectra: List, charge: int, mode: str) -> Iterator[List]:
        """Reference :372-455 — the explicit candidate lists (window mask AND ANN ids, ascending
        library position, invalid spectra dropped). The fused ``_search_batch`` never
        materialises these; this generator exists for callers and parity tests that want them."""
        tol_val, tol_mode = self._mode_tolerance(mode)
        if charge not in self._library_reader.spec_info["charge"]:
            return
        info = self._library_reader.spec_info["charge"][charge]
        query_mzs = np.array([s.precursor_mz for s in query_spectra], float).reshape(-1, 1)
        library_mzs = np.asarray(info["precursor_mz"], np.float32).astype(np.float64).reshape(1, -1)
        if tol_mode == "Da":
            candidate_filters = np.abs(query_mzs - library_mzs) * charge <= tol_val
        elif tol_mode == "ppm":
            candidate_filters = np.abs(query_mzs - library_mzs) / library_mzs * 10 ** 6 <= tol_val
        else:
            raise ValueError("Unknown precursor tolerance mode")
        if self._use_ann(charge, mode):
            q = spectra_to_store(query_spectra, with_charge=False)
            mz = q["mz64"] if any(np.asarray(s.mz).dtype != np.float32 for s in query_spectra) else q["mz"]
            vecs = self._engine.vectorize(mz, q["inten"], q["off"], True)
            _, I = self._engine.ivf_search(charge, vecs, self._num_candidates, self._num_probe, want_d=False)
            mask = np.zeros_like(candidate_filters)
            for mask_i, ann_filter in zip(mask, I):
                mask_i[ann_filter[ann_filter != -1]] = True
            candidate_filters = np.logical_and(candidate_filters, mask)
        for candidate_filter in candidate_filters:
            query_candidates = []
            for idx in info["id"][candidate_filter]:
                candidate = self._library_reader.read_spectrum(idx, True)
                if candidate.is_valid:
                    query_candidates.append(candidate)
            yield query_candidates

"""Host-side mirror of the feature half of the reference's src/ann_solo/utils.py for rescoring:
``_compute_ssm_features`` (utils.py:276-457) keeps its name, argument and the dictionary it returns,
but the per-SSM Python loop over two ``SpectrumSimilarityCalculator`` objects is ONE device launch
(K6, csrc/k6_ssm_features.cu). Model fitting (mokapot / scikit-learn, utils.py:24-273) is out of
scope and stays with the reference.
"""
from __future__ import annotations

import collections
from typing import Dict, Iterable, List, Optional

import numpy as np

from .spectrum import SpectrumSpectrumMatch, default_engine, spectra_to_store


def _compute_ssm_features(ssms: Iterable[SpectrumSpectrumMatch], engine=None,
                          library_rows: Optional[Dict[int, Dict]] = None) -> Dict[str, List]:
    """Reference utils.py:276-457. Returns the same dictionary: "index", "sequence", the 44 feature
    columns and "is_target", with one entry per SSM that has peak matches (others are skipped,
    :332-333).

    The library side of every SSM must be resident in the engine's peak store of the query's
    precursor charge: ``library_rows[charge][library_identifier]`` is its row there (the mapping
    ``SpectralLibrary`` keeps); without it ``ssm.library_spectrum.index`` is taken as the row."""
    engine = engine or default_engine()
    names = engine.feature_names()
    ssms = list(ssms)
    features = {"index": [], "sequence": []}
    features.update({n: [] for n in names})
    features["is_target"] = []
    by_charge = collections.defaultdict(list)
    for i, ssm in enumerate(ssms):
        if len(ssm.peak_matches) == 0:
            continue
        by_charge[ssm.library_spectrum.precursor_charge].append(i)
    rows_out = {}
    for charge, idx in by_charge.items():
        sel = [ssms[i] for i in idx]
        queries = [s.query_spectrum for s in sel]
        q = spectra_to_store(queries, with_charge=False)
        mz_vec = q["mz64"] if any(np.asarray(s.mz).dtype != np.float32 for s in queries) else None
        max_pairs = max(len(s.peak_matches) for s in sel)
        pairs = np.zeros((len(sel), max_pairs, 2), np.uint32)
        for j, s in enumerate(sel):
            pairs[j, :len(s.peak_matches)] = np.asarray(s.peak_matches, np.int64)
        n_pairs = np.array([len(s.peak_matches) for s in sel], np.int32)
        if library_rows is not None:
            lib_row = np.array([library_rows[charge][s.library_identifier] for s in sel], np.int32)
        else:
            lib_row = np.array([s.library_spectrum.index for s in sel], np.int32)
        q_charge = np.array([s.query_spectrum.precursor_charge for s in sel], np.int32)
        seq_len = np.array([len(s.sequence) if s.sequence is not None else 0 for s in sel], np.int32)
        table = engine.ssm_features(charge, q, lib_row, pairs, n_pairs, q_charge, seq_len, mz_vec)
        for j, i in enumerate(idx):
            rows_out[i] = table[j]
    int_cols = {"sequence_len", "precursor_charge_2", "precursor_charge_3", "precursor_charge_4", "precursor_charge_5",
                "n_matched_peaks"}
    for i in sorted(rows_out):  # the reference appends in SSM order
        ssm = ssms[i]
        features["index"].append(i)
        features["sequence"].append(ssm.sequence)
        for n, v in zip(names, rows_out[i]):
            features[n].append(int(v) if n in int_cols else float(v))
        features["is_target"].append(not ssm.is_decoy)
    return features

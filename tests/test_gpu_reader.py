"""End to end from a ``.splib`` file: native parser -> K0 process_spectrum -> device peak stores ->
ANN indexes cached in .idxann files -> fused search, through ``SpectralLibrary(filename)`` (the
reference's constructor signature, spectral_library.py:46-117)."""
import glob
import os

import numpy as np
import pytest

from oracle import splib_io

pytestmark = pytest.mark.gpu
ANN = {0: "?", 1: "b3/0.01", 2: "y5^2/-0.01"}


def _write_raw_library(path, lib, rng):
    """Every processed synthetic spectrum becomes a raw one: ranks scaled to counts, plus weak noise
    peaks below min_intensity * max that process_spectrum must drop again."""
    specs = []
    for r in range(len(lib["prec_mz"])):
        b, e = lib["off"][r], lib["off"][r + 1]
        inten = (50 - np.argsort(np.argsort(-lib["inten"][b:e], kind="stable"), kind="stable")) * 100.0  # rank * 100
        n_noise = int(rng.integers(10, 40))
        nmz = rng.uniform(60.0, 1990.0, n_noise)
        mz = np.concatenate([lib["mz"][b:e].astype(np.float64), nmz])
        it = np.concatenate([inten, rng.uniform(0.1, 0.009 * inten.max(), n_noise)])
        ann = [ANN[int(c)] for c in lib["chg"][b:e]] + ["?"] * n_noise
        order = np.argsort(mz, kind="stable")
        specs.append(dict(id=5000 + r, peptide="PEPTIDE" + "ACDEFGHIK"[r % 9] * (r % 5 + 1), charge=int(lib["prec_z"][r]),
                          prec_mz=float(lib["prec_mz"][r]), mz=mz[order], intensity=it[order],
                          annotations=[ann[i] for i in order], decoy=bool(lib["is_decoy"][r])))
    splib_io.write_splib(path, specs)
    return specs


def test_spectral_library_from_splib_file(engine, synth, tmp_path):
    from ann_solo_b200.config import config
    from ann_solo_b200.spectral_library import InMemoryLibrary, SpectralLibrary
    from ann_solo_b200.spectrum import process_spectrum
    rng = np.random.default_rng(131)
    lib = synth.make_library(2400, seed=131, decoy_seed=132)
    queries = synth.make_queries(lib, 150, seed=133)
    path = str(tmp_path / "human.splib")
    _write_raw_library(path, lib, rng)
    config.update(dict(num_list=16, num_probe=8, num_candidates=64, precursor_tolerance_mass_open=300.0,
                       precursor_tolerance_mode_open="Da", mode="ann"))
    try:
        sl = SpectralLibrary(path, engine=engine, train_iters=3)
        reader = sl._library_reader
        # K0 gave back exactly the processed spectra the raw ones were made from
        for z, (store, rows) in synth.split_by_charge(lib).items():
            ps = reader.charge_store(z)
            assert np.array_equal(ps["off"], store["off"]) and ps["valid"].all()
            assert np.array_equal(ps["mz"], store["mz"]) and np.array_equal(ps["chg"], store["chg"])
            assert np.array_equal(ps["inten"].view(np.uint32), store["inten"].view(np.uint32))
            assert np.array_equal(reader.spec_info["charge"][z]["id"], np.array([str(5000 + r) for r in rows]))
            assert np.array_equal(reader.spec_info["charge"][z]["precursor_mz"], store["prec_mz"].astype(np.float32))
        # ... and the per-spectrum host process_spectrum of the raw object agrees with the K0 row
        for spec_id in ("5000", "5017", "6203"):
            raw = reader.read_spectrum(spec_id)
            assert not raw.is_processed and len(raw.mz) > 30
            host = process_spectrum(raw, True)
            dev = reader.read_spectrum(spec_id, True)
            assert dev.is_processed and dev.is_valid == host.is_valid
            assert np.array_equal(dev.mz, host.mz) and np.array_equal(dev.intensity.view(np.uint32),
                                                                      host.intensity.view(np.uint32))
            assert [None if a is None else a.charge for a in dev.annotation] == \
                   [None if a is None else a.charge for a in host.annotation]
        # index files next to the library, named like the reference names them
        files = sorted(glob.glob(str(tmp_path / "human_*.idxann")))
        assert [os.path.basename(f) for f in files] == [f"human_{sl._get_hyperparameter_hash()[:7]}_{z}.idxann"
                                                        for z in sorted(sl._ann_charges)]
        # the search equals the one over the in-memory synthetic library with the same centroids
        cents = {z: engine.ivf_get_centroids(z) for z in sl._ann_charges}
        qreader = InMemoryLibrary(queries)
        qs = [qreader.read_spectrum(i) for i in range(150)]
        for s in qs:
            s.is_processed = True
        got = {}
        for z in (2, 3):
            for ssm in sl._search_batch([s for s in qs if s.precursor_charge == z], z, "open"):
                got[ssm.query_identifier] = (int(ssm.library_identifier) - 5000, ssm.search_engine_score,
                                             ssm.peak_matches.tolist(), ssm.sequence)
        ref = SpectralLibrary(InMemoryLibrary(lib), engine=engine, centroids=cents)
        want = {}
        for z in (2, 3):
            for ssm in ref._search_batch([s for s in qs if s.precursor_charge == z], z, "open"):
                want[ssm.query_identifier] = (int(ssm.library_identifier), ssm.search_engine_score,
                                              ssm.peak_matches.tolist())
        assert got.keys() == want.keys() and len(got) > 80
        for k in want:
            assert got[k][:3] == want[k] and got[k][3].startswith("PEPTIDE")
        # second construction: the .idxann files are read, not rebuilt
        stamp = [os.path.getmtime(f) for f in files]
        sl2 = SpectralLibrary(path, engine=engine, train_iters=1)
        assert [os.path.getmtime(f) for f in files] == stamp
        assert all(np.array_equal(engine.ivf_get_centroids(z), cents[z]) for z in sl2._ann_charges)
        with pytest.raises(FileNotFoundError):
            SpectralLibrary(str(tmp_path / "missing.splib"), engine=engine)
        with pytest.raises(FileNotFoundError):
            SpectralLibrary(str(tmp_path / "human.mgf"), engine=engine)
    finally:
        config.update(dict(num_list=256, num_probe=128, num_candidates=1024))


def test_search_takes_a_query_file_name(engine, synth, tmp_path):
    """SpectralLibrary.search(query_filename) (reference spectral_library.py:193-215): raw MGF queries
    are read natively, processed, and searched through the cascade; the identifications equal those
    of the same raw spectra handed over as objects."""
    from oracle import mgf_io
    from ann_solo_b200.config import config
    from ann_solo_b200.spectral_library import SpectralLibrary
    from ann_solo_b200.spectrum import MsmsSpectrum
    rng = np.random.default_rng(141)
    lib = synth.make_library(2400, seed=141, decoy_seed=142)
    queries = synth.make_queries(lib, 120, seed=143)
    lib_path, mgf_path = str(tmp_path / "lib.splib"), str(tmp_path / "queries.mgf")
    _write_raw_library(lib_path, lib, rng)
    entries, objects = [], []
    for i in range(120):
        b, e = queries["off"][i], queries["off"][i + 1]
        inten = (50 - np.argsort(np.argsort(-queries["inten"][b:e], kind="stable"), kind="stable")) * 100.0
        n_noise = int(rng.integers(5, 30))
        mz = np.concatenate([queries["mz"][b:e].astype(np.float64), rng.uniform(60.0, 1990.0, n_noise)])
        it = np.concatenate([inten, rng.uniform(0.1, 0.009 * inten.max(), n_noise)])
        order = np.argsort(mz, kind="stable")
        z = int(queries["prec_z"][i])
        entry = dict(title=f"q{i}", prec_mz=float(queries["prec_mz"][i]), mz=mz[order], intensity=it[order])
        if i % 10:                      # every tenth query has no CHARGE line: searched as 2+ and 3+
            entry["charge"] = f"{z}+"
        entries.append(entry)
        objects.append(MsmsSpectrum(f"q{i}", entry["prec_mz"], z if i % 10 else None, mz[order],
                                    it[order].astype(np.float32)))
    mgf_io.write_mgf(mgf_path, entries)
    config.update(dict(num_list=16, num_probe=8, num_candidates=64, precursor_tolerance_mass=20.0,
                       precursor_tolerance_mode="ppm", precursor_tolerance_mass_open=300.0,
                       precursor_tolerance_mode_open="Da", fdr=0.05, mode="ann"))
    try:
        sl = SpectralLibrary(lib_path, engine=engine, train_iters=3)
        from_file = {s.query_identifier: (s.library_identifier, s.search_engine_score) for s in sl.search(mgf_path)}
        from_objects = {s.query_identifier: (s.library_identifier, s.search_engine_score) for s in sl.search(objects)}
        assert from_file == from_objects and len(from_file) > 50
        truth = queries["truth"]
        correct = sum(1 for q, (lid, _) in from_file.items() if truth[int(q[1:])] >= 0 and
                      int(lid) - 5000 == truth[int(q[1:])])
        assert correct > 35
    finally:
        config.update(dict(num_list=256, num_probe=128, num_candidates=1024))

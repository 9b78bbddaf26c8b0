"""The drop-in claim, tested on the reference's own code: /root/reference/src/ann_solo/spectral_library.py is
loaded UNMODIFIED (tests/ref_harness.py) with the modules it imports replaced, and its own
`SpectralLibrary.__init__` / `_create_ann_indexes` / `search` / `_search_cascade` / `_search_batch` /
`_get_library_candidates` / `_get_ann_index` (reference :46-500) run on a synthetic library.

* CPU (`-m "not gpu"`): faiss -> an oracle-backed stand-in, spectrum_match -> the oracle scorer. The SSMs the
  reference module produces must equal the oracle pipeline's (oracle.candidates + scorer) — this pins the oracle's
  glue (window mask, ANN mask, cascade, first-SSM rule) on the reference's code.
* GPU (`-m gpu`): faiss -> ann_solo_b200.index, spectrum_match -> ann_solo_b200.spectrum_match, spectrum ->
  ann_solo_b200.spectrum (K1 vectoriser) or the reference's own spectrum.py. The SSMs must equal those of the fused
  device path (ann_solo_b200.SpectralLibrary) over the very same .idxann files.

/root/reference does not exist on the GPU box: the GPU tests skip there unless $SOLO_REFERENCE_SRC points at a copy
(they run in the build container's CPU suite through the oracle stand-ins instead).
"""
import os
import types

import numpy as np
import pytest

import ref_harness
from conftest import canon_pairs

needs_reference = pytest.mark.skipif(not ref_harness.have_reference(), reason="reference sources not present")

CFG = dict(precursor_tolerance_mass=20, precursor_tolerance_mode="ppm", precursor_tolerance_mass_open=300,
           precursor_tolerance_mode_open="Da", fragment_mz_tolerance=0.02, allow_peak_shifts=True, num_list=16,
           num_probe=6, num_candidates=96, batch_size=50, mode="ann", no_gpu=True, model="none")


def _world(synth, n_targets=1800, n_queries=150):
    from ann_solo_b200.spectral_library import InMemoryLibrary
    from ann_solo_b200.spectrum import MsmsSpectrum
    lib = synth.make_library(n_targets, seed=21, decoy_seed=22, charges=(2, 3), charge_p=(0.6, 0.4))
    ident = np.arange(len(lib["prec_mz"])) * 3 + 7     # identifiers are not row numbers
    inmem = InMemoryLibrary(lib, ident, peptides=[f"PEP{i}" for i in range(len(ident))])
    qs = synth.make_queries(lib, n_queries, seed=23, charges=(2, 3), charge_p=(0.6, 0.4))
    # exact copies of some library spectra so that the first cascade level (20 ppm) identifies something
    for j, r in enumerate(range(0, 40, 2)):
        b, e = lib["off"][r], lib["off"][r + 1]
        qb, qe = qs["off"][j], qs["off"][j + 1]
        if e - b == qe - qb:
            qs["mz"][qb:qe] = lib["mz"][b:e]
            qs["inten"][qb:qe] = lib["inten"][b:e]
        qs["prec_mz"][j] = lib["prec_mz"][r] * (1 + 2e-6)
        qs["prec_z"][j] = lib["prec_z"][r]

    def query_objects():
        out = []
        for i in range(n_queries):
            b, e = qs["off"][i], qs["off"][i + 1]
            s = MsmsSpectrum(f"q{i}", qs["prec_mz"][i], int(qs["prec_z"][i]), qs["mz"][b:e].copy(), qs["inten"][b:e].copy())
            s.is_processed, s.is_valid = True, True      # synthetic queries are processed spectra already
            out.append(s)
        return out

    return lib, inmem, qs, query_objects


def _key(ssms):
    return {s.query_identifier: (s.library_identifier, canon_pairs(np.asarray(s.peak_matches), len(s.peak_matches)).tolist())
            for s in ssms}


# ------------------------------------------------------------------ CPU: oracle-backed stand-ins
def _oracle_faiss(oracle):
    from oracle import faiss_io
    m = types.ModuleType("faiss")
    m.METRIC_INNER_PRODUCT = 0
    m.get_num_gpus = lambda: 0
    m.built = {}

    class IndexFlatIP:
        def __init__(self, d):
            self.d = d

    class IndexIVF:
        pass

    class IndexIVFFlat(IndexIVF):
        def __init__(self, quantizer, d, nlist, metric=0):
            self.d, self.nlist, self.nprobe = d, nlist, 1
            self.cent = self.lists = None

        def train(self, x):
            self.cent = oracle.kmeans(np.ascontiguousarray(x, np.float32), self.nlist, seed=4, iters=3)

        def add(self, x):
            x = np.ascontiguousarray(x, np.float32)
            self.lists = oracle.build_lists(x, oracle.ivf_assign(x, self.cent), self.nlist)

        def search(self, q, k):
            off, ids, vecs = self.lists
            return oracle.ivf_search(np.ascontiguousarray(q, np.float32), self.cent, off, ids, vecs,
                                     min(self.nprobe, self.nlist), k)

        def reset(self):
            self.lists = None

    def write_index(index, fname):
        off, ids, vecs = index.lists
        faiss_io.write_ivf_flat(fname, index.cent, [ids[off[i]:off[i + 1]] for i in range(index.nlist)],
                                [vecs[off[i]:off[i + 1]] for i in range(index.nlist)])
        m.built[fname] = index

    def read_index(fname):
        f = faiss_io.read_ivf_flat(fname)
        index = IndexIVFFlat(None, f["d"], f["nlist"])
        index.cent = f["centroids"]
        sizes = np.array([len(i) for i in f["list_ids"]], np.int64)
        off = np.zeros(f["nlist"] + 1, np.int64)
        np.cumsum(sizes, out=off[1:])
        index.lists = (off, np.concatenate(f["list_ids"]).astype(np.int32), np.concatenate(f["list_vecs"]))
        return index

    m.IndexFlatIP, m.IndexIVF, m.IndexIVFFlat, m.write_index, m.read_index = IndexFlatIP, IndexIVF, IndexIVFFlat, write_index, read_index
    return m


def _oracle_spectrum_match(oracle):
    from ann_solo_b200.spectrum import spectra_to_store
    m = types.ModuleType("ann_solo.spectrum_match")

    def get_best_match(query, candidates, fragment_mz_tolerance, allow_shift):
        lib = spectra_to_store(candidates, with_charge=True)
        q = spectra_to_store([query], with_charge=False)
        ids, off = np.arange(len(candidates), dtype=np.int32), np.array([0, len(candidates)], np.int64)
        fn = oracle.ref_best_match_batch if oracle.have_ref() else oracle.best_match_batch
        bp, bs, npairs, pairs = fn(q, lib, ids, off, fragment_mz_tolerance, bool(allow_shift), max_pairs=max(1, len(query.mz)))
        return candidates[int(bp[0])], float(bs[0]), [(int(a), int(b)) for a, b in pairs[0, :int(npairs[0])]]

    m.get_best_match = get_best_match
    return m


def _expected_psms(oracle, lib, inmem, qs, indexes, ref_spectrum_mod):
    """The cascade as the oracle pipeline runs it: level 1 = window only, level 2 (queries without a level-1 SSM)
    = ANN ids AND window; reference spectral_library.py:229-258 with the test's pass-through FDR (q = 0)."""
    from ann_solo_b200.spectrum import MsmsSpectrum
    out = {}
    for z in sorted(inmem.rows):
        rows = inmem.rows[z]
        store = inmem.charge_store(z)
        qi = np.flatnonzero(qs["prec_z"] == z)
        if len(qi) == 0:
            continue
        from ann_solo_b200.synth import take_spectra
        q = take_spectra(qs, qi)
        done = np.zeros(len(qi), bool)
        for level, (tol, mode) in enumerate([(CFG["precursor_tolerance_mass"], CFG["precursor_tolerance_mode"]),
                                             (CFG["precursor_tolerance_mass_open"], CFG["precursor_tolerance_mode_open"])]):
            todo = np.flatnonzero(~done)
            if len(todo) == 0:
                continue
            sub = take_spectra(q, todo)
            ann = None
            if level == 1 and z in indexes:
                qv = np.zeros((len(todo), 800), np.float32)
                for i, t in enumerate(todo):
                    b, e = q["off"][t], q["off"][t + 1]
                    ref_spectrum_mod.spectrum_to_vector(MsmsSpectrum(0, 500.0, z, q["mz"][b:e], q["inten"][b:e]), 11, 2010,
                                                        0.04, 800, True, qv[i])
                ix = indexes[z]
                ix.nprobe = CFG["num_probe"]
                ann = ix.search(qv, CFG["num_candidates"])[1]
            cand, coff = oracle.candidates(sub["prec_mz"], store["prec_mz32"], store["valid"], z, float(tol), mode, ann)
            fn = oracle.ref_best_match_batch if oracle.have_ref() else oracle.best_match_batch
            bp, bs, npairs, pairs = fn(sub, store, cand, coff, CFG["fragment_mz_tolerance"], True, max_pairs=50)
            for i, t in enumerate(todo):
                if bp[i] < 0:
                    continue
                row = int(cand[coff[i] + bp[i]])
                out[f"q{qi[t]}"] = (int(inmem.identifiers[rows[row]]), canon_pairs(pairs[i], int(npairs[i])).tolist())
                done[t] = True
    return out


@needs_reference
def test_reference_module_runs_on_oracle_standins_and_equals_oracle_pipeline(oracle, synth, tmp_path):
    lib, inmem, qs, query_objects = _world(synth)
    faiss_mod = _oracle_faiss(oracle)
    surface = ref_harness.ReaderSurface(inmem)
    sl, cfg, restore = ref_harness.load_reference(faiss_mod, _oracle_spectrum_match(oracle),
                                                  ref_harness.reader_module(surface, query_objects()),
                                                  ref_harness.utils_module(), config_values=CFG)
    try:
        import sys
        ref_spectrum = sys.modules["ann_solo.spectrum"]
        assert ref_spectrum.__file__.startswith(ref_harness.REF_SRC)        # the reference's own spectrum.py
        sl.SpectralLibrary._ann_filenames = {}                               # class attribute in the reference (:41)
        library = sl.SpectralLibrary(str(tmp_path / "lib.splib"))            # reference :46-117, builds + writes the indexes
        assert sorted(library._ann_filenames) == [2, 3]
        assert all(os.path.isfile(f) for f in library._ann_filenames.values())
        got = _key(library.search("queries.mgf"))                            # reference :193-260
        want = _expected_psms(oracle, lib, inmem, qs, {z: faiss_mod.built[f] for z, f in library._ann_filenames.items()},
                              ref_spectrum)
        assert len(want) > 100
        assert got.keys() == want.keys()
        for k in want:
            assert got[k] == want[k], k
        # both cascade levels produced SSMs
        n_exact = sum(1 for k, v in got.items() if int(k[1:]) < 20)
        assert n_exact >= 15
        # the reference's _get_library_candidates (:372-455) against the oracle's candidate lists, open level
        z = 2
        qobjs = [s for s in query_objects() if s.precursor_charge == z][:40]
        cands = list(library._get_library_candidates(qobjs, z, "open"))
        store = inmem.charge_store(z)
        qv = np.zeros((len(qobjs), 800), np.float32)
        for i, s in enumerate(qobjs):
            ref_spectrum.spectrum_to_vector(s, 11, 2010, 0.04, 800, True, qv[i])
        ix = faiss_mod.built[library._ann_filenames[z]]
        ix.nprobe = CFG["num_probe"]
        ann = ix.search(qv, CFG["num_candidates"])[1]
        cand, coff = oracle.candidates(np.array([s.precursor_mz for s in qobjs]), store["prec_mz32"], store["valid"], z, 300.0,
                                       "Da", ann)
        ids = inmem.spec_info["charge"][z]["id"]
        for i, c in enumerate(cands):
            assert [s.identifier for s in c] == ids[cand[coff[i]:coff[i + 1]]].tolist()
        library.shutdown()
    finally:
        restore()


@needs_reference
def test_reference_process_spectrum_runs_on_the_spectrum_standin(oracle):
    """The reference's own process_spectrum (spectrum.py:57-119) drives the MsmsSpectrum stand-in's spectrum_utils
    methods; the result equals the oracle restatement (round, precursor removal, filter, scale, norm)."""
    from ann_solo_b200.spectrum import MsmsSpectrum
    cfgv = dict(CFG, resolution=1, remove_precursor=True, remove_precursor_tolerance=0.5, scaling="sqrt")
    sl, cfg, restore = ref_harness.load_reference(_oracle_faiss(oracle), _oracle_spectrum_match(oracle),
                                                  ref_harness.reader_module(None, []), ref_harness.utils_module(),
                                                  config_values=cfgv)
    try:
        import sys
        ref_spectrum = sys.modules["ann_solo.spectrum"]
        rng = np.random.default_rng(3)
        n_valid = 0
        for i in range(30):
            n = int(rng.integers(30, 500))
            mz = np.sort(rng.uniform(5.0, 2100.0, n))
            inten = rng.exponential(1.0, n).astype(np.float32)
            s = ref_spectrum.process_spectrum(MsmsSpectrum(i, 650.0, 2, mz.copy(), inten.copy()), False)
            omz, oint, ovalid, _ = oracle.process_spectrum_np(mz, inten, 650.0, 2, resolution=1, remove_precursor=True,
                                                              remove_precursor_tolerance=0.5, scaling="sqrt")
            assert s.is_valid == ovalid and s.is_processed
            if ovalid:
                n_valid += 1
                assert np.array_equal(s.mz, omz)
                # the reference normalises with np.linalg.norm (float32), the restatement with a float64 sum
                np.testing.assert_allclose(s.intensity, oint, rtol=2e-7)
        assert n_valid > 15
    finally:
        restore()


# ------------------------------------------------------------------ GPU: the product modules under the reference's code
@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize("own_vectoriser", [True, False])
def test_reference_module_on_device_modules_equals_fused_path(synth, tmp_path, own_vectoriser):
    """own_vectoriser: ann_solo.spectrum -> ann_solo_b200.spectrum (K1) — SSMs identical to the fused path;
    otherwise the reference's own Python spectrum_to_vector feeds the device index (its float32 norm differs from
    K1's in the last bit, which can move a candidate across the num_candidates boundary: >= 98 % identical SSMs)."""
    import ann_solo_b200.index as index_mod
    import ann_solo_b200.spectrum as spectrum_mod
    import ann_solo_b200.spectrum_match as sm_mod
    from ann_solo_b200.config import config as my_config
    from ann_solo_b200.spectral_library import SpectralLibrary
    lib, inmem, qs, query_objects = _world(synth)
    surface = ref_harness.ReaderSurface(inmem)
    sl, cfg, restore = ref_harness.load_reference(index_mod, sm_mod, ref_harness.reader_module(surface, query_objects()),
                                                  ref_harness.utils_module(),
                                                  spectrum_mod=spectrum_mod if own_vectoriser else None, config_values=CFG)
    saved = dict(my_config._ns)
    try:
        sl.SpectralLibrary._ann_filenames = {}
        library = sl.SpectralLibrary(str(tmp_path / "lib.splib"))
        got = _key(library.search("queries.mgf"))
        files = dict(library._ann_filenames)
        library.shutdown()
        # the fused device path over the same index files
        my_config.update({k: v for k, v in CFG.items() if k in my_config._ns})
        mine = SpectralLibrary(inmem, ann_basename=str(tmp_path / "lib"),
                               score_ssms=lambda ssms, fdr, model, is_open: [setattr(s, "q", 0.0) or s for s in ssms])
        assert {z: mine._ann_filenames[z] for z in files} == files           # same cache names, nothing rebuilt
        want = _key(mine.search(query_objects()))
        mine.shutdown()
        assert len(want) > 100 and got.keys() == want.keys()
        same = sum(got[k] == want[k] for k in want)
        if own_vectoriser:
            assert same == len(want)
        else:
            assert same >= 0.98 * len(want)
    finally:
        my_config._ns.clear()
        my_config._ns.update(saved)
        restore()

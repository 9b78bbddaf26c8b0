"""Host logic and the drop-in boundary (no GPU compute)."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_cabi_exports_every_declared_symbol():
    from ann_solo_b200 import _lib
    header = open(os.path.join(ROOT, "include", "solo_b200.h")).read()
    declared = set(re.findall(r"\b(solo_[a-z0-9_]+)\s*\(", header))
    declared -= {"solo_handle", "solo_search_params"}
    assert len(declared) >= 30
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()  # binds every symbol; AttributeError if one is missing from the .so
    assert lib.solo_version().startswith(b"solo_b200")
    assert lib.solo_profile_num_stages() == 10
    assert lib.solo_stage_name(4) == b"scan"


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from ann_solo_b200.engine import SoloEngine
    from ann_solo_b200 import SoloError
    with pytest.raises(SoloError, match="no CPU fallback"):
        SoloEngine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ann-solo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert not re.search(r"#include\s+[\"<][^\">]*oracle", src), f
                assert "libsolo_oracle" not in src and "libsolo_ref" not in src, f


def test_config_defaults_match_reference():
    # reference src/ann_solo/config.py:71-216
    from ann_solo_b200.config import Config
    c = Config()
    assert (c.min_mz, c.max_mz, c.bin_size, c.hash_len) == (11, 2010, 0.04, 800)
    assert (c.min_peaks, c.min_mz_range, c.max_peaks_used, c.max_peaks_used_library) == (10, 250, 50, 50)
    assert (c.num_candidates, c.batch_size, c.num_list, c.num_probe) == (1024, 16384, 256, 128)
    assert c.scaling == "rank" and c.min_intensity == 0.01 and c.mode == "ann"
    with pytest.raises(AttributeError):
        c.not_a_key


def test_get_dim_host(oracle):
    from ann_solo_b200.spectrum import get_dim
    assert get_dim(11, 2010, 0.04) == oracle.get_dim(11, 2010, 0.04) == (49976, 10.96, 2010.0)


@pytest.mark.parametrize("scaling", ["rank", "sqrt", None])
def test_process_spectrum_matches_oracle_restatement(oracle, scaling):
    from ann_solo_b200.config import config
    from ann_solo_b200.spectrum import MsmsSpectrum, process_spectrum
    rng = np.random.default_rng(5)
    config.update(dict(scaling=scaling, remove_precursor=True, remove_precursor_tolerance=0.5))
    try:
        n_valid = 0
        for i in range(60):
            n = int(rng.integers(5, 400))
            mz = np.sort(rng.uniform(5.0, 2100.0, n))
            inten = rng.exponential(1.0, n).astype(np.float32)
            pm, z = float(rng.uniform(300, 1200)), int(rng.integers(2, 5))
            s = process_spectrum(MsmsSpectrum(i, pm, z, mz.copy(), inten.copy()), is_library=bool(i % 2))
            omz, oint, ovalid, _ = oracle.process_spectrum_np(mz, inten, pm, z, remove_precursor=True,
                                                              remove_precursor_tolerance=0.5, scaling=scaling)
            assert s.is_valid == ovalid and s.is_processed
            if ovalid:
                n_valid += 1
                assert np.array_equal(s.mz, omz) and np.array_equal(s.intensity, oint)
                assert len(s.mz) <= 50 and abs(float(np.sum(s.intensity.astype(np.float64) ** 2)) - 1) < 1e-5
        assert n_valid > 20
    finally:
        config.update(dict(scaling="rank", remove_precursor=False, remove_precursor_tolerance=0))


@pytest.mark.parametrize("resolution", [0, 1, 2])
def test_process_spectrum_with_resolution(oracle, resolution):
    """round(resolution, 'sum') (reference spectrum.py:84-89): the host method equals the oracle restatement, the
    merged peak keeps the annotation of its group's most intense member."""
    from ann_solo_b200.config import config
    from ann_solo_b200.spectrum import MsmsSpectrum, process_spectrum
    rng = np.random.default_rng(9 + resolution)
    config.update(dict(resolution=resolution))
    try:
        n_valid = 0
        for i in range(40):
            n = int(rng.integers(50, 600))
            mz = np.sort(rng.uniform(5.0, 2100.0, n))
            if i % 2:
                mz = mz.astype(np.float32)
            inten = rng.exponential(1.0, n).astype(np.float32)
            ann = list(range(n))
            s = process_spectrum(MsmsSpectrum(i, 600.0, 2, mz.copy(), inten.copy(), annotation=ann), is_library=True)
            omz, oint, ovalid, oidx = oracle.process_spectrum_np(mz, inten, 600.0, 2, resolution=resolution)
            assert s.is_valid == ovalid
            if ovalid:
                n_valid += 1
                assert np.array_equal(s.mz, omz) and np.array_equal(s.intensity, oint)
                assert s.annotation == list(oidx)
        assert n_valid > 20
    finally:
        config.update(dict(resolution=None))
    sp = MsmsSpectrum(0, 500.0, 2, np.array([100.01, 100.04, 100.2, 250.0]), np.array([1.0, 3.0, 2.0, 5.0], np.float32),
                      annotation=["a", "b", "c", "d"])
    sp.round(1, "sum")
    assert np.array_equal(sp.mz, [100.0, 100.2, 250.0]) and np.array_equal(sp.intensity, [4.0, 2.0, 5.0])
    assert sp.annotation == ["b", "c", "d"]


def test_rank_scaling_is_what_synth_generates(synth):
    lib = synth.make_library(200, seed=3)
    for r in range(0, 200, 17):
        b, e = lib["off"][r], lib["off"][r + 1]
        raw = lib["inten"][b:e].astype(np.float64)
        ranks = np.sort(raw / raw.min() * (50 - (e - b) + 1))
        np.testing.assert_allclose(ranks, np.arange(50 - (e - b) + 1, 51), rtol=1e-5)
        assert (np.diff(lib["mz"][b:e]) > 0).all()


def test_spectra_to_store_and_inmemory_library(synth):
    from ann_solo_b200.spectral_library import InMemoryLibrary
    from ann_solo_b200.spectrum import spectra_to_store
    lib = synth.make_library(300, seed=4)
    reader = InMemoryLibrary(lib)
    assert set(reader.spec_info["charge"]) <= {2, 3, 4}
    ids = reader.spec_info["charge"][2]["id"][:20]
    spectra = [reader.read_spectrum(i, True) for i in ids]
    st = spectra_to_store(spectra)
    ref = synth.take_spectra(lib, np.asarray(ids))
    for k in ("mz", "inten", "chg", "off", "prec_mz", "prec_z"):
        assert np.array_equal(st[k], ref[k]), k
    assert reader.spec_info["charge"][2]["precursor_mz"].dtype == np.float32


def test_search_mode_errors():
    from ann_solo_b200.engine import SoloEngine
    with pytest.raises(ValueError, match="Unknown precursor tolerance mode"):
        SoloEngine.make_params(True, 10, 1, 1.0, "mmu", 0.02, True)


def test_partition_helpers(synth):
    from ann_solo_b200 import parallel
    assert [parallel.shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    lib = synth.make_library(50, decoy_fraction=0, seed=1)
    parts = [parallel.shard_store(lib, r, 3) for r in range(3)]
    assert sum(len(p["prec_mz"]) for p in parts) == 50
    assert np.array_equal(np.concatenate([p["mz"] for p in parts]), lib["mz"])
    assert all(p["off"][0] == 0 and p["off"][-1] == len(p["mz"]) for p in parts)
    owner = parallel.assign_lists(np.array([5, 1, 9, 3, 3, 7]), 2)
    loads = [np.array([5, 1, 9, 3, 3, 7])[owner == r].sum() for r in range(2)]
    assert abs(loads[0] - loads[1]) <= 2


def test_merge_topk_equals_global(oracle, synth):
    from ann_solo_b200 import parallel
    lib = synth.make_library(1200, decoy_fraction=0, seed=2)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    x[700] = x[3]  # a cross-shard exact tie
    cent = oracle.kmeans(x, 12, iters=2)
    assign = oracle.ivf_assign(x, cent)
    off, ids, vecs = oracle.build_lists(x, assign, 12)
    q = x[:25]
    D, I = oracle.ivf_search(q, cent, off, ids, vecs, nprobe=6, k=40)
    owner = parallel.assign_lists(np.diff(off), 3)
    Dp, Ip = [], []
    for r in range(3):
        # a rank scans only the probed lists it owns: emulate by emptying the others
        keep = owner[assign] == r
        a_r = np.where(keep, assign, -1)
        o_r, i_r, v_r = oracle.build_lists(x, a_r, 12)
        d, i = oracle.ivf_search(q, cent, o_r, i_r, v_r, nprobe=6, k=40)
        Dp.append(d)
        Ip.append(i)
    Dm, Im = parallel.merge_topk(Dp, Ip, 40)
    assert np.array_equal(Im, I) and np.array_equal(Dm, D)


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/solo_b200.h must compile as C (no C++-isms, no torch types)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "hdr.c"
    src.write_text('#include "solo_b200.h"\n'
                   'int main(void) { solo_process_params p; solo_idxann_info i; solo_search_params s;\n'
                   '  (void)p; (void)i; (void)s; return SOLO_N_SSM_FEATURES == 44 ? 0 : 1; }\n')
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "hdr.o")])
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "solo_b200.h")).read(), flags=re.S)
    assert "torch" not in header and "std::" not in header and "at::" not in header   # outside comments

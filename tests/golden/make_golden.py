"""Mint golden vectors for the hot path FROM THE REFERENCE ITSELF (run in the CPU container,
where /root/reference exists; the fixtures it writes are committed and travel to the GPU box).

  * vectoriser: the reference's unmodified src/ann_solo/spectrum.py is imported with three stub
    modules (mmh3 -> sklearn's MurmurHash3_x86_32, an empty spectrum_utils.spectrum.MsmsSpectrum,
    configargparse -> argparse), exactly as probed in SURVEY.md §8c. Outputs: hash_idx for every
    bin of the default configuration, get_dim for several settings, spectrum_to_vector for random
    spectra held as float32 and as float64 m/z arrays.
  * scorer: the reference's SpectrumMatch.cpp compiled from where it lies (oracle/_ref) and
    called through oracle/ref_shim.cpp with stable buffers: the KAT1-9 cases of SURVEY.md §8c plus
    seeded random batches (inputs and outputs stored).

  * SSM features: the reference's unmodified spectrum_similarity.py is imported
    the same way (scipy 1.18 dropped the Pearson/SpearmanRConstantInputWarning names the module
    refers to; they are aliased to ConstantInputWarning). First the reference's OWN known answers
    (src/tests/spectrum_similarity_test.py, run unmodified with a minimal MsmsSpectrum stub) are
    checked against it here; then SpectrumSimilarityCalculator is driven call for call like
    utils.py:330-455 on the reference test's six spectra pairs and on seeded synthetic SSMs whose
    peak matches come from the reference's own SpectrumMatch.cpp.

Usage: python tests/golden/make_golden.py [vectoriser] [scorer] [features]
"""
import argparse
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/ann_solo"
sys.path.insert(0, ROOT)


def load_reference_spectrum():
    from sklearn.utils import murmurhash3_32

    mmh3 = types.ModuleType("mmh3")
    mmh3.hash = lambda key, seed=0, signed=True: int(murmurhash3_32(key, seed=seed, positive=not signed))
    sys.modules["mmh3"] = mmh3
    su = types.ModuleType("spectrum_utils")
    sus = types.ModuleType("spectrum_utils.spectrum")

    class MsmsSpectrum:  # only used as a type annotation by the reference module
        pass

    sus.MsmsSpectrum = MsmsSpectrum
    su.spectrum = sus
    sys.modules["spectrum_utils"] = su
    sys.modules["spectrum_utils.spectrum"] = sus
    cap = types.ModuleType("configargparse")

    class ArgParser(argparse.ArgumentParser):
        def __init__(self, *a, **k):
            for key in ("default_config_files", "args_for_setting_config_path", "formatter_class"):
                k.pop(key, None)
            super().__init__()

    cap.ArgParser = ArgParser
    cap.ArgumentDefaultsHelpFormatter = argparse.ArgumentDefaultsHelpFormatter
    sys.modules["configargparse"] = cap
    pkg = types.ModuleType("ann_solo")
    pkg.__path__ = [REF]
    sys.modules["ann_solo"] = pkg
    for name in ("config", "spectrum"):
        spec = importlib.util.spec_from_file_location(f"ann_solo.{name}", os.path.join(REF, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"ann_solo.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["ann_solo.spectrum"]


class Spec:
    def __init__(self, mz, intensity):
        self.mz = mz
        self.intensity = intensity


def golden_vectoriser():
    ref = load_reference_spectrum()
    out = {}
    n_bins, min_bound, _ = ref.get_dim(11, 2010, 0.04)
    out["hash_800"] = np.array([ref.hash_idx(b, 800) for b in range(n_bins + 2)], np.uint16)
    out["hash_400_first1000"] = np.array([ref.hash_idx(b, 400) for b in range(1000)], np.uint16)
    dims = []
    for mn, mx, bs in [(11, 2010, 0.04), (50, 1500, 0.05), (0, 2000, 1.0005), (101, 1999, 0.02)]:
        n, s, e = ref.get_dim(mn, mx, bs)
        dims.append([mn, mx, bs, n, s, e])
    out["get_dim"] = np.array(dims, np.float64)
    rng = np.random.default_rng(11)
    n_spec = 64
    counts = rng.integers(10, 51, n_spec)
    off = np.zeros(n_spec + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    mz64 = np.concatenate([np.sort(rng.uniform(11.0, 2010.0, c)) for c in counts])
    # include exact bin boundaries and values that bin differently in f32 vs f64
    mz64[:8] = np.sort(np.array([11.0, 100.0, 200.02, 300.5, 1000.0, 1500.04, 1999.99, 2009.99]))
    inten = rng.uniform(0.01, 1.0, off[-1]).astype(np.float32)
    mz32 = mz64.astype(np.float32)
    v32 = np.zeros((n_spec, 800), np.float32)
    v64 = np.zeros((n_spec, 800), np.float32)
    v64_raw = np.zeros((n_spec, 800), np.float32)
    bins32, bins64 = [], []
    import math
    for i in range(n_spec):
        b, e = off[i], off[i + 1]
        ref.spectrum_to_vector(Spec(mz32[b:e], inten[b:e]), 11, 2010, 0.04, 800, True, v32[i])
        ref.spectrum_to_vector(Spec(mz64[b:e], inten[b:e]), 11, 2010, 0.04, 800, True, v64[i])
        ref.spectrum_to_vector(Spec(mz64[b:e], inten[b:e]), 11, 2010, 0.04, 800, False, v64_raw[i])
        bins32 += [math.floor((m - min_bound) // 0.04) for m in mz32[b:e]]
        bins64 += [math.floor((m - min_bound) // 0.04) for m in mz64[b:e]]
    out.update(mz64=mz64, inten=inten, off=off, v32=v32, v64=v64, v64_raw=v64_raw,
               bins32=np.array(bins32, np.int64), bins64=np.array(bins64, np.int64))
    np.savez_compressed(os.path.join(HERE, "vectoriser.npz"), **out)
    print("vectoriser.npz:", {k: v.shape for k, v in out.items()}, "numpy", np.__version__)


def store(specs):
    mz, inten, chg, off, pm, pz = [], [], [], [0], [], []
    for s in specs:
        mz += list(s["mz"])
        inten += list(s["I"])
        chg += list(s.get("chg", [0] * len(s["mz"])))
        off.append(len(mz))
        pm.append(s["prec"])
        pz.append(s["z"])
    return dict(mz=np.array(mz, np.float32), inten=np.array(inten, np.float32), chg=np.array(chg, np.uint8),
                off=np.array(off, np.int64), prec_mz=np.array(pm, np.float64), prec_z=np.array(pz, np.int32))


def golden_scorer():
    from oracle import solo_oracle as so
    import importlib.util as iu
    spec = iu.spec_from_file_location("synth", os.path.join(ROOT, "ann-solo_b200", "synth.py"))
    synth = iu.module_from_spec(spec)
    spec.loader.exec_module(synth)
    q0 = dict(prec=500.0, z=2, mz=[100, 200, 300, 400], I=[.5] * 4)
    base = dict(prec=490, z=2, mz=[100, 200, 290, 380], I=[.5] * 4)
    kats = [
        ("KAT1", q0, [base], False), ("KAT2", q0, [base], True),
        ("KAT3", q0, [dict(base, chg=[0, 0, 2, 1])], True), ("KAT4", q0, [dict(base, chg=[0, 0, 1, 2])], True),
        ("KAT5", q0, [dict(base, prec=500.005)], True),
        ("KAT6", dict(prec=500.0, z=2, mz=[100.00, 100.03, 300.0], I=[0.6, 0.7, 0.3873]),
         [dict(prec=500, z=2, mz=[100.015, 300.0], I=[0.8, 0.6])], False),
        ("KAT7", q0, [base, base], True),
        ("KAT8", q0, [dict(prec=510, z=2, mz=[100, 200, 310, 420], I=[.5] * 4)], True),
        ("KAT9", dict(q0, z=3), [dict(prec=494, z=3, mz=[100, 194, 291, 382], I=[.5] * 4)], True),
        ("KAT10", q0, [dict(prec=500, z=2, mz=[100, 200.01, 350, 400], I=[.5] * 4)], True),
    ]
    out = []
    for name, q, cands, shift in kats:
        qs, ls = store([q]), store(cands)
        bp, bs, npairs, pairs = so.ref_best_match_batch(qs, ls, np.arange(len(cands)), np.array([0, len(cands)]),
                                                        0.02, shift)
        out.append(dict(name=name, query=q, candidates=cands, allow_shift=shift, tol=0.02, best=int(bp[0]),
                        score=float(bs[0]), pairs=pairs[0, :npairs[0]].tolist()))
        print(name, bp[0], repr(float(bs[0])), pairs[0, :npairs[0]].tolist())
    with open(os.path.join(HERE, "scorer_kat.json"), "w") as f:
        json.dump(out, f, indent=1)
    # seeded random batch: 2,000-spectrum library, 96 queries x 48 candidates, both shift modes
    lib = synth.make_library(1500, seed=21, decoy_seed=22)
    qs = synth.make_queries(lib, 96, seed=23)
    rng = np.random.default_rng(24)
    n_lib = len(lib["prec_mz"])
    cand = np.empty((96, 48), np.int32)
    for i in range(96):
        cand[i] = np.sort(rng.choice(n_lib, 48, replace=False))
        if qs["truth"][i] >= 0:
            cand[i, rng.integers(48)] = qs["truth"][i]
        cand[i] = np.sort(cand[i])
    cand_off = np.arange(0, 96 * 48 + 1, 48, dtype=np.int64)
    res = {}
    for shift in (0, 1):
        bp, bs, npairs, pairs = so.ref_best_match_batch(qs, lib, cand.ravel(), cand_off, 0.02, bool(shift))
        res[f"best_{shift}"] = bp
        res[f"score_{shift}"] = bs
        res[f"npairs_{shift}"] = npairs
        res[f"pairs_{shift}"] = pairs.astype(np.uint8)
    keep = ("mz", "inten", "chg", "off", "prec_mz", "prec_z")
    np.savez_compressed(os.path.join(HERE, "scorer_random.npz"),
                        **{f"lib_{k}": lib[k] for k in keep}, **{f"q_{k}": qs[k] for k in keep if k in qs},
                        cand=cand, cand_off=cand_off, **res)
    print("scorer_random.npz written; mean pairs", res["npairs_1"].mean())


def load_reference_similarity():
    import scipy.stats
    for old in ("PearsonRConstantInputWarning", "SpearmanRConstantInputWarning"):
        if not hasattr(scipy.stats, old):
            setattr(scipy.stats, old, scipy.stats.ConstantInputWarning)
    ref_spectrum = load_reference_spectrum()
    sus = sys.modules["spectrum_utils.spectrum"]

    class MsmsSpectrum:  # the constructor surface the reference's test uses (spectrum_utils 0.4 dtypes)
        def __init__(self, identifier, precursor_mz, precursor_charge, mz, intensity, **kw):
            self.identifier, self.precursor_mz, self.precursor_charge = identifier, precursor_mz, precursor_charge
            order = np.argsort(mz, kind="stable")   # spectrum_utils keeps peaks ordered by m/z
            self.mz = np.asarray(mz, np.float64)[order]
            self.intensity = np.asarray(intensity, np.float32)[order]

    sus.MsmsSpectrum = MsmsSpectrum
    spec = importlib.util.spec_from_file_location("ann_solo.spectrum_similarity",
                                                  os.path.join(REF, "spectrum_similarity.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ann_solo.spectrum_similarity"] = mod
    sys.modules["ann_solo"].spectrum = ref_spectrum
    sys.modules["ann_solo"].spectrum_similarity = mod
    spec.loader.exec_module(mod)
    return ref_spectrum, mod, MsmsSpectrum


def reference_feature_row(ref_spectrum, sim, ssm, q_charge, min_mz=11, max_mz=2010, bin_size=0.04):
    """The similarity columns in the order and with the arguments of reference utils.py:330-455."""
    c, t = sim.SpectrumSimilarityCalculator(ssm), sim.SpectrumSimilarityCalculator(ssm, 5)
    qm, lm = ssm.query_spectrum.precursor_mz, ssm.library_spectrum.precursor_mz
    ppm, da = (qm - lm) / lm * 10 ** 6, qm - lm   # spectrum_utils.utils.mass_diff(mz1, mz2, mode_is_da)
    return [0, int(q_charge <= 2), int(q_charge == 3), int(q_charge == 4), int(q_charge >= 5), qm, lm, ppm, abs(ppm),
            da, abs(da), c.cosine(), t.cosine(), c.n_matched_peaks(), c.frac_n_peaks_query(), c.frac_n_peaks_library(),
            t.frac_n_peaks_library(), c.frac_intensity_query(), c.frac_intensity_library(), t.frac_intensity_library(),
            c.mean_squared_error("mz"), t.mean_squared_error("mz"), c.mean_squared_error("intensity"),
            t.mean_squared_error("intensity"), c.spectral_contrast_angle(), t.spectral_contrast_angle(),
            c.hypergeometric_score(min_mz=min_mz, max_mz=max_mz, fragment_mz_tol=bin_size), c.kendalltau(),
            c.ms_for_id_v1(), c.ms_for_id_v2(), c.entropy(False), c.entropy(True), c.scribe_fragment_acc(),
            t.scribe_fragment_acc(), c.manhattan(), c.euclidean(), c.chebyshev(), c.pearsonr(), t.pearsonr(),
            c.spearmanr(), t.spearmanr(), c.braycurtis(), c.canberra(), c.ruzicka()]


def golden_features():
    import subprocess
    import warnings
    ref_spectrum, sim, MsmsSpectrum = load_reference_similarity()
    # 1. the reference's own known answers, unmodified test file, against the module as imported here
    boot = os.path.join(HERE, "_run_reference_similarity_tests.py")
    rc = subprocess.call([sys.executable, boot])
    assert rc == 0, "the reference's spectrum_similarity_test.py does not pass in this environment"
    # 2. the six pairs of that test file through the utils.py call sequence
    import importlib.util as iu
    tspec = iu.spec_from_file_location("ref_sim_test", "/root/reference/src/tests/spectrum_similarity_test.py")
    tmod = iu.module_from_spec(tspec)
    tspec.loader.exec_module(tmod)
    cases = []

    def add(name, q_mz, q_int, l_mz, l_int, pairs, q_prec, q_z, l_prec):
        q = MsmsSpectrum("q", q_prec, q_z, q_mz, q_int)
        lib = MsmsSpectrum("l", l_prec, q_z, l_mz, l_int)
        q.mz = np.asarray(q_mz)            # keep the caller's m/z precision (float32 or float64)
        lib.mz = np.asarray(l_mz)
        ssm = ref_spectrum.SpectrumSpectrumMatch(q, lib, np.asarray(pairs, np.int64).reshape(-1, 2))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            row = reference_feature_row(ref_spectrum, sim, ssm, q_z)
        cases.append(dict(name=name, q_mz=np.asarray(q_mz), q_int=np.asarray(q_int, np.float32),
                          l_mz=np.asarray(l_mz), l_int=np.asarray(l_int, np.float32),
                          pairs=np.asarray(pairs, np.int64).reshape(-1, 2), q_prec=q_prec, q_z=q_z, l_prec=l_prec,
                          row=np.array(row, np.float64)))

    for fx in ("all_match", "partial_match"):
        f = getattr(tmod, fx)
        f = getattr(f, "__wrapped__", None) or f._get_wrapped_function()
        calc = f()
        # rebuild the SSM the fixture wrapped (the calculator keeps the arrays)
        pm = np.stack([[np.flatnonzero(calc.mz_query == a)[0], np.flatnonzero(calc.mz_library == b)[0]]
                       for a, b in zip(calc.matched_mz_query, calc.matched_mz_library)])
        add("reftest_" + fx, calc.mz_query, calc.int_query, calc.mz_library, calc.int_library, pm, 465.227, 2, 453.75)
    # 3. seeded synthetic SSMs: peak matches from the reference's SpectrumMatch.cpp
    from oracle import solo_oracle as so
    spec = iu.spec_from_file_location("synth", os.path.join(ROOT, "ann-solo_b200", "synth.py"))
    synth = iu.module_from_spec(spec)
    spec.loader.exec_module(synth)
    lib = synth.make_library(1200, seed=31, decoy_seed=32)
    qs = synth.make_queries(lib, 160, seed=33)
    rng = np.random.default_rng(34)
    n_lib = len(lib["prec_mz"])
    cand = np.empty((160, 8), np.int32)
    for i in range(160):
        cand[i] = rng.choice(n_lib, 8, replace=False)
        if qs["truth"][i] >= 0 and i % 4:
            cand[i, 0] = qs["truth"][i]
        cand[i] = np.sort(cand[i])
    cand_off = np.arange(0, 160 * 8 + 1, 8, dtype=np.int64)
    bp, bs, npairs, pairs = so.ref_best_match_batch(qs, lib, cand.ravel(), cand_off, 0.02, True)
    for i in range(160):
        if npairs[i] == 0:
            continue
        r = cand[i, bp[i]]
        a, b, c, d = qs["off"][i], qs["off"][i + 1], lib["off"][r], lib["off"][r + 1]
        q_mz = qs["mz"][a:b] if i % 2 else qs["mz"][a:b].astype(np.float64) + rng.normal(0, 1e-5, b - a)
        q_int, l_int = qs["inten"][a:b].copy(), lib["inten"][c:d].copy()
        if i % 5 == 0:   # 'root'-style scaling with ties inside a spectrum (exercises the tie terms)
            q_int = np.round(np.sqrt(q_int) * 8) / 8 + np.float32(0.125)
            l_int = np.round(np.sqrt(l_int) * 8) / 8 + np.float32(0.125)
            top = np.argsort(-l_int, kind="stable")[:6]   # keep the top-5 cut free of ties
            l_int[top] += np.arange(6, 0, -1, dtype=np.float32)
            q_int /= np.linalg.norm(q_int)
            l_int /= np.linalg.norm(l_int)
        add(f"synth_{i}", q_mz, q_int, lib["mz"][c:d], l_int, pairs[i, :npairs[i]], float(qs["prec_mz"][i]),
            int(qs["prec_z"][i]) if i % 7 else 5, float(lib["prec_mz"][r]))
    # 4. edge cases: one match, no match among the top-5 library peaks, constant matched intensities
    add("one_match", [100., 200, 300, 400, 500, 600, 700, 800, 900, 1000], np.arange(1, 11) / np.sqrt(385.),
        [100.01, 210, 310, 410, 510, 610, 710, 810, 910, 1010], np.arange(10, 0, -1) / np.sqrt(385.), [[0, 0]],
        500.0, 2, 510.0)
    add("no_top5_match", [100., 200, 300, 400, 500, 600, 700, 800, 900, 1000], np.arange(1, 11) / np.sqrt(385.),
        [100.01, 200.01, 310, 410, 510, 610, 710, 810, 910, 1010], np.arange(1, 11) / np.sqrt(385.),
        [[0, 0], [1, 1]], 500.0, 3, 499.0)
    # (five library peaks: the top-5 cut of np.argpartition is then free of ties, which it resolves arbitrarily)
    add("constant_matched", [100., 200, 300, 400], [.5, .5, .5, .5], [100., 200, 300, 400, 450],
        [.4, .4, .4, .4, .6], [[0, 0], [1, 1], [2, 2], [3, 3]], 500.0, 4, 500.0)
    out = dict(n=len(cases), names=np.array([c["name"] for c in cases]))
    for k in ("q_prec", "q_z", "l_prec"):
        out[k] = np.array([c[k] for c in cases], np.float64 if k != "q_z" else np.int32)
    out["rows"] = np.stack([c["row"] for c in cases])
    for k in ("q_mz", "q_int", "l_mz", "l_int"):
        out[k] = np.concatenate([np.asarray(c[k], np.float64 if k.endswith("mz") else np.float32) for c in cases])
        out[k + "_off"] = np.cumsum([0] + [len(c[k]) for c in cases])
    out["q_mz_is_f32"] = np.array([np.asarray(c["q_mz"]).dtype == np.float32 for c in cases])
    out["pairs"] = np.concatenate([c["pairs"] for c in cases])
    out["pairs_off"] = np.cumsum([0] + [len(c["pairs"]) for c in cases])
    import scipy
    out["versions"] = np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}"])
    np.savez_compressed(os.path.join(HERE, "ssm_features.npz"), **out)
    print("ssm_features.npz:", len(cases), "SSMs;", "rows", out["rows"].shape)


if __name__ == "__main__":
    what = sys.argv[1:] or ["vectoriser", "scorer", "features"]
    if "vectoriser" in what:
        golden_vectoriser()
    if "scorer" in what:
        golden_scorer()
    if "features" in what:
        golden_features()

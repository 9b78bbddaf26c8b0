"""SSM feature table (reference utils.py:276-457 / spectrum_similarity.py) on the
host: the oracle restatement, and the K6 kernel's own __host__ __device__ source compiled for the
CPU (oracle/k6_host_check.cpp), against golden rows minted from the reference's unmodified
spectrum_similarity.py (tests/golden/make_golden.py features).

Tolerance. The reference sums float32 intensities in float32 (NumPy) and SciPy 1.18 returns float32
statistics for float32 input; the oracle and the kernel work in float64. Agreed bound:
|got - want| <= 1e-5 * |want| + 5e-6, except
  * contrast_angle(_top5): arccos is ill-conditioned at cosine -> 1 (a float32 cosine one ulp below
    1.0 moves the angle by 3e-4): checked in cosine space to 1e-6 and directly to 4e-4;
  * kendalltau: SciPy's float32 p-value is subnormal below 1.2e-38 (-log p > 87.3) and underflows to 0
    (-> inf) below 0.7e-45: there the float64 p-value must round to the same float32 subnormal
    (|p - p_ref| <= 0.71e-45), resp. the float64 -log p must exceed 103.2.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import ssm_features as sf

NAMES = sf.FEATURE_NAMES
BINS = sf.n_peak_bins(11, 2010, 0.04)


def golden_cases():
    g = np.load(os.path.join(GOLDEN, "ssm_features.npz"))
    for i in range(int(g["n"])):
        a, b = g["q_mz_off"][i:i + 2]
        c, d = g["l_mz_off"][i:i + 2]
        e, f = g["pairs_off"][i:i + 2]
        q_mz = g["q_mz"][a:b].astype(np.float32) if g["q_mz_is_f32"][i] else g["q_mz"][a:b]
        yield dict(name=str(g["names"][i]), q_mz=q_mz, q_int=g["q_int"][a:b], l_mz=g["l_mz"][c:d].astype(np.float32),
                   l_int=g["l_int"][c:d], pairs=g["pairs"][e:f], q_prec=float(g["q_prec"][i]), q_z=int(g["q_z"][i]),
                   l_prec=float(g["l_prec"][i]), row=g["rows"][i])


def assert_rows_close(got, want, label=""):
    """The tolerance of the module docstring, column by column."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape == (len(NAMES),)
    for k, name in enumerate(NAMES):
        g, w = got[k], want[k]
        where = f"{label} {name}: got {g!r}, want {w!r}"
        if name == "kendalltau" and np.isinf(w):
            assert g > 103.2, where
        elif name == "kendalltau" and w > 87.3:
            assert abs(g - w) <= 1e-5 * abs(w) + 5e-6 or abs(np.exp(-g) - np.exp(-w)) <= 0.71e-45, where
        elif np.isinf(w) or np.isnan(w):
            assert (np.isinf(g) and g == w) or (np.isnan(g) and np.isnan(w)), where
        elif name.startswith("contrast_angle"):
            assert abs(g - w) <= 4e-4, where
            assert abs(np.cos(np.pi / 2 * (1 - g)) - np.cos(np.pi / 2 * (1 - w))) <= 1e-6, where
        else:
            assert abs(g - w) <= 1e-5 * abs(w) + 5e-6, where


def test_feature_names_follow_the_reference_dict():
    # reference utils.py:296-340 (keys of `features` without index / sequence / is_target)
    assert len(NAMES) == 44 and NAMES[0] == "sequence_len" and NAMES[11] == "cosine" and NAMES[-1] == "ruzicka"
    assert BINS == 49976
    from ann_solo_b200 import _lib
    lib = _lib.load()
    assert [lib.solo_ssm_feature_name(i).decode() for i in range(_lib.N_SSM_FEATURES)] == NAMES
    assert lib.solo_ssm_feature_name(44) == b""


def test_golden_holds_the_reference_tests_known_answers():
    """Rows of the two matching pairs of reference src/tests/spectrum_similarity_test.py carry the
    numbers that file asserts (:443-844)."""
    rows = {c["name"]: dict(zip(NAMES, c["row"])) for c in golden_cases() if c["name"].startswith("reftest")}
    p, a = rows["reftest_partial_match"], rows["reftest_all_match"]
    want = dict(cosine=0.44582117, cosine_top5=0.85880862, n_matched_peaks=8, frac_n_peaks_query=8 / 14,
                frac_n_peaks_lib=8 / 14, frac_n_peaks_lib_top5=4 / 5, frac_int_query=0.45378598,
                frac_int_lib=0.75759018, contrast_angle=0.29417655, kendalltau=4.25896654, ms_for_id_v1=21.03216848,
                ms_for_id_v2=30.03222119, manhattan=2.98346427, euclidean=1.05278566, chebyshev=0.5802746,
                pearsonr=0.69570652, pearsonr_top5=0.24177300, spearmanr=0.59933680, spearmanr_top5=0.19999999,
                braycurtis=0.58102504, canberra=12.30376030, ruzicka=0.26500210, scribe_fragment_acc=0.86739458,
                scribe_fragment_acc_top5=1.02137350, entropy_unweighted=0.53600209, entropy_weighted=0.59836031)
    for k, v in want.items():
        assert p[k] == pytest.approx(v, rel=2e-6, abs=1e-7), k
    assert a["cosine"] == pytest.approx(1.0) and a["kendalltau"] == pytest.approx(19.29406731)
    assert a["hypergeometric_score"] == 100.0 and a["ms_for_id_v1"] == 1000.0 and a["scribe_fragment_acc"] == 10.0
    assert a["ms_for_id_v2"] == pytest.approx(154.45107128, rel=2e-6)


def test_oracle_matches_golden():
    n = 0
    for c in golden_cases():
        got = sf.ssm_features(c["q_mz"], c["q_int"], c["l_mz"], c["l_int"], c["pairs"], c["q_prec"], c["q_z"],
                              c["l_prec"], 0, BINS)
        assert_rows_close(got, c["row"], c["name"])
        n += 1
    assert n > 100


def _host_kernel():
    from oracle import solo_oracle
    solo_oracle.build()
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libk6_host_check.so"))
    lib.k6_host_ssm_features.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int64,
                                         C.c_void_p]
    return lib


def host_kernel_row(lib, c, seq_len=0):
    f64 = c["q_mz"].dtype == np.float64
    q_mz = np.ascontiguousarray(c["q_mz"])
    q_int, l_mz, l_int = (np.ascontiguousarray(c[k], np.float32) for k in ("q_int", "l_mz", "l_int"))
    pairs = np.ascontiguousarray(c["pairs"], np.uint32)
    out = np.empty(44)
    rc = lib.k6_host_ssm_features(None if f64 else q_mz.ctypes.data, q_mz.ctypes.data if f64 else None,
                                  q_int.ctypes.data, len(q_int), l_mz.ctypes.data, l_int.ctypes.data, len(l_int),
                                  pairs.ctypes.data, len(pairs), c["q_prec"], c["l_prec"], c["q_z"], seq_len, BINS,
                                  out.ctypes.data)
    assert rc == 0
    return out


def test_kernel_source_on_host_matches_golden_and_oracle():
    lib = _host_kernel()
    assert lib.k6_host_n_features() == 44
    for c in golden_cases():
        got = host_kernel_row(lib, c, seq_len=7)
        assert got[0] == 7
        got[0] = 0
        assert_rows_close(got, c["row"], c["name"])
        want = sf.ssm_features(c["q_mz"], c["q_int"], c["l_mz"], c["l_int"], c["pairs"], c["q_prec"], c["q_z"],
                               c["l_prec"], 0, BINS)
        # kernel vs oracle (both float64): tight, except where lgamma-based binomials meet exact ones
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11, err_msg=c["name"])


def test_kendall_exact_and_asymptotic_against_scipy():
    import scipy.stats
    rng = np.random.default_rng(8)
    for n in (2, 3, 5, 12, 33, 34, 50):
        for _ in range(4):
            x, y = rng.random(n), rng.random(n)
            if n == 5:
                x[1] = x[3]            # ties -> asymptotic
            if n == 34:
                y = np.sort(y)[np.argsort(np.argsort(x))]  # perfectly concordant: exact closed form beyond n = 33
            want = scipy.stats.kendalltau(x, y)[1]
            assert sf.kendalltau_pvalue(x, y) == pytest.approx(want, rel=1e-10), n
    assert np.isnan(sf.kendalltau_pvalue([1.0], [2.0])) and np.isnan(sf.kendalltau_pvalue([1.0, 1.0], [2.0, 3.0]))


def test_kernel_source_on_host_equals_oracle_on_random_ssms():
    """Seeded random SSMs beyond the golden set: tiny spectra (fewer than five library peaks), single
    matches, tied and zero intensities, up to 128 peaks — the kernel's source (host build) against the
    float64 oracle."""
    lib = _host_kernel()
    rng = np.random.default_rng(99)
    n_done = 0
    for t in range(400):
        nq = int(rng.integers(1, 129 if t % 40 == 0 else 60))
        nl = int(rng.integers(1, 129 if t % 40 == 1 else 60))
        q_mz = np.sort(rng.uniform(100, 1900, nq)).astype(np.float32 if t % 2 else np.float64)
        l_mz = np.sort(rng.uniform(100, 1900, nl)).astype(np.float32)
        if t % 3 == 0:      # ties (and zeros) inside the spectra; the top-5 cut kept free of ties
            q_int = rng.integers(0, 6, nq).astype(np.float32)
            l_int = rng.integers(1, 6, nl).astype(np.float32)
            top = np.argsort(-l_int, kind="stable")[:6]
            l_int[top] += np.arange(len(top), 0, -1, dtype=np.float32) * 10
            if not q_int.any():
                q_int[0] = 1
        else:
            q_int = rng.random(nq).astype(np.float32) + np.float32(0.01)
            l_int = rng.random(nl).astype(np.float32) + np.float32(0.01)
        q_int /= np.linalg.norm(q_int)
        l_int /= np.linalg.norm(l_int)
        m = int(rng.integers(1, min(nq, nl) + 1))
        pairs = np.stack([rng.choice(nq, m, replace=False), rng.choice(nl, m, replace=False)], axis=1)
        c = dict(q_mz=q_mz, q_int=q_int, l_mz=l_mz, l_int=l_int, pairs=pairs, q_prec=float(rng.uniform(300, 900)),
                 q_z=int(rng.integers(1, 7)), l_prec=float(rng.uniform(300, 900)))
        got = host_kernel_row(lib, c, seq_len=t % 31)
        want = sf.ssm_features(q_mz, q_int, l_mz, l_int, pairs, c["q_prec"], c["q_z"], c["l_prec"], t % 31, BINS)
        both_nan = np.isnan(got) & np.isnan(want)
        np.testing.assert_allclose(np.where(both_nan, 0, got), np.where(both_nan, 0, want), rtol=1e-9, atol=1e-11,
                                   err_msg=f"case {t}: nq {nq} nl {nl} m {m}")
        n_done += 1
    assert n_done == 400

"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's SSM feature computation
 — the 44 numeric columns `_compute_ssm_features` builds for every spectrum–
spectrum match (reference utils.py:276-457) from `SpectrumSimilarityCalculator`
(reference spectrum_similarity.py:13-730), in float64, one SSM at a time with plain loops.

Pinned: tests/golden/ssm_features.npz holds inputs and outputs minted by importing the
reference's unmodified spectrum_similarity.py (tests/golden/make_golden.py, which also re-runs the
reference's own spectrum_similarity_test.py known answers against it in this container). Where the
reference accumulates in float32 (NumPy sums/dots of float32 intensities) this restatement and the
CUDA kernel accumulate in float64; the agreed tolerance is written in tests/test_ssm_features.py.

SciPy pieces restated (scipy is the reference's dependency, unpinned; behaviour of 1.18.1 here):
  kendalltau(x, y)[1]   tau-b p-value, method 'auto': exact (Kendall 1970 recursion) without ties and
                        n <= 33 (or min(dis, tot-dis) <= 1), else the normal approximation with tie terms
  pearsonr / spearmanr  NaN (-> 0.0) for a constant input; Spearman = Pearson of average ranks
  special.comb          here exact integer binomials (math.comb) instead of float approximations
  stats.entropy         natural log of the normalised intensities
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

FEATURE_NAMES = [
    "sequence_len", "precursor_charge_2", "precursor_charge_3", "precursor_charge_4", "precursor_charge_5",
    "query_prec_mz", "lib_prec_mz", "mz_diff_ppm", "abs_mz_diff_ppm", "mz_diff_da", "abs_mz_diff_da",
    "cosine", "cosine_top5", "n_matched_peaks", "frac_n_peaks_query", "frac_n_peaks_lib", "frac_n_peaks_lib_top5",
    "frac_int_query", "frac_int_lib", "frac_int_lib_top5", "mse_mz", "mse_mz_top5", "mse_int", "mse_int_top5",
    "contrast_angle", "contrast_angle_top5", "hypergeometric_score", "kendalltau", "ms_for_id_v1", "ms_for_id_v2",
    "entropy_unweighted", "entropy_weighted", "scribe_fragment_acc", "scribe_fragment_acc_top5", "manhattan",
    "euclidean", "chebyshev", "pearsonr", "pearsonr_top5", "spearmanr", "spearmanr_top5", "braycurtis", "canberra",
    "ruzicka",
]
TOP = 5  # utils.py:336


def n_peak_bins(min_mz: float, max_mz: float, bin_size: float) -> int:
    """spectrum.get_dim(...)[0] (reference spectrum.py:123-143) as used by hypergeometric_score."""
    start = min_mz - min_mz % bin_size
    end = max_mz + bin_size - max_mz % bin_size
    return math.ceil((end - start) / bin_size)


# ------------------------------------------------------------------ scipy restatements
def kendall_p_exact(n: int, c: int) -> float:
    """Two-sided exact p-value for `c` concordant pairs of n untied observations (Kendall 1970)."""
    tot = n * (n - 1) // 2
    c = min(c, tot - c)
    if n == 1 or n == 2:
        return 1.0
    if c == 0:
        return 2.0 / math.factorial(n)
    if c == 1:
        return 2.0 / math.factorial(n - 1)
    if 4 * c == n * (n - 1):
        return 1.0
    new = np.zeros(c + 1)
    new[0:2] = 1.0
    for j in range(3, n + 1):
        cum = np.cumsum(new)
        if j <= c:
            cum[j:] = cum[j:] - cum[:c + 1 - j].copy()
        new = cum
    return float(min(max(2.0 * new.sum() / math.factorial(n), 0.0), 1.0))


def kendalltau_pvalue(x, y) -> float:
    """scipy.stats.kendalltau(x, y)[1] (variant b, method auto, two-sided); NaN when undefined."""
    n = len(x)
    if n == 0:
        return math.nan
    dis = xtie = ytie = ntie = 0
    for i in range(n):
        for j in range(i + 1, n):
            dx, dy = x[i] - x[j], y[i] - y[j]
            if dx == 0 and dy == 0:
                ntie += 1
            if dx == 0:
                xtie += 1
            if dy == 0:
                ytie += 1
            if dx * dy < 0:
                dis += 1
    tot = n * (n - 1) // 2
    if xtie == tot or ytie == tot:
        return math.nan
    con_minus_dis = tot - xtie - ytie + ntie - 2 * dis
    if xtie == 0 and ytie == 0 and (n <= 33 or min(dis, tot - dis) <= 1):
        return kendall_p_exact(n, tot - dis)

    def tie_terms(v):
        _, cnt = np.unique(np.asarray(v), return_counts=True)
        cnt = cnt[cnt > 1].astype(np.float64)
        return (cnt * (cnt - 1) * (cnt - 2)).sum(), (cnt * (cnt - 1) * (2 * cnt + 5)).sum()

    x0, x1 = tie_terms(x)
    y0, y1 = tie_terms(y)
    m = n * (n - 1.0)
    var = (m * (2 * n + 5) - x1 - y1) / 18 + (2 * xtie * ytie) / m + x0 * y0 / (9 * m * (n - 2))
    z = con_minus_dis / math.sqrt(var)
    return math.erfc(abs(z) / math.sqrt(2.0))  # 2 * norm.sf(|z|)


def pearson(x, y) -> float:
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    if len(x) < 2 or (x == x[0]).all() or (y == y[0]).all():
        return 0.0  # NaN in SciPy -> 0.0 in the reference (spectrum_similarity.py:487, :514)
    xm, ym = x - x.mean(), y - y.mean()
    r = float((xm * ym).sum() / math.sqrt((xm * xm).sum() * (ym * ym).sum()))
    return min(max(r, -1.0), 1.0)


def average_ranks(v) -> np.ndarray:
    v = np.asarray(v, np.float64)
    r = np.empty(len(v))
    for i in range(len(v)):
        r[i] = (v < v[i]).sum() + ((v == v[i]).sum() + 1) / 2.0
    return r


def entropy(p) -> float:
    p = np.asarray(p, np.float64)
    p = p / p.sum()
    p = p[p > 0]
    return float(-(p * np.log(p)).sum())


def spectrum_entropy(inten, weighted: bool) -> float:
    """reference spectrum_similarity.py:703-730."""
    s = entropy(inten)
    if not weighted or s > 3:
        return s
    w = np.asarray(inten, np.float64) ** (0.25 + 0.25 * s)
    return entropy(w)


def hypergeometric(n_matched: int, n_lib: int, bins: int) -> float:
    """reference spectrum_similarity.py:251-309, with exact integer binomials."""
    denom = math.comb(bins, n_lib)
    p = Fraction(0)
    for i in range(n_matched + 1, n_lib + 1):
        p += Fraction(math.comb(n_lib, i) * math.comb(bins - n_lib, n_lib - i), denom)
    p = float(p)
    return 100.0 if p <= 0.0 else min(-math.log(p), 100.0)


# ------------------------------------------------------------------ one SSM
def ssm_features(q_mz, q_int, l_mz, l_int, pairs, q_prec_mz: float, q_charge: int, l_prec_mz: float,
                 sequence_len: int, bins: int) -> np.ndarray:
    """One row of the feature table, columns FEATURE_NAMES. `pairs` is (M, 2) (query peak, library
    peak), M >= 1 (the reference skips SSMs without peak matches, utils.py:332-333)."""
    q_mz, l_mz = np.asarray(q_mz, np.float64), np.asarray(l_mz, np.float64)
    q_int, l_int = np.asarray(q_int, np.float64), np.asarray(l_int, np.float64)
    pairs = np.asarray(pairs, np.int64).reshape(-1, 2)
    f = dict.fromkeys(FEATURE_NAMES, 0.0)
    f["sequence_len"] = float(sequence_len)
    z = "precursor_charge_%d" % (2 if q_charge <= 2 else 5 if q_charge >= 5 else q_charge)
    f[z] = 1.0
    f["query_prec_mz"], f["lib_prec_mz"] = q_prec_mz, l_prec_mz
    f["mz_diff_da"] = q_prec_mz - l_prec_mz
    f["mz_diff_ppm"] = (q_prec_mz - l_prec_mz) / l_prec_mz * 10 ** 6
    f["abs_mz_diff_da"], f["abs_mz_diff_ppm"] = abs(f["mz_diff_da"]), abs(f["mz_diff_ppm"])

    qi, li = pairs[:, 0], pairs[:, 1]
    mq, ml = q_int[qi], l_int[li]
    uq = np.delete(q_int, qi)
    lib_unmatched = np.setdiff1d(np.arange(len(l_int)), li)
    ul = l_int[lib_unmatched]
    m = len(pairs)
    # ---- all peaks
    f["cosine"] = float((mq * ml).sum())
    f["n_matched_peaks"] = float(m)
    f["frac_n_peaks_query"] = m / len(q_mz)
    f["frac_n_peaks_lib"] = m / len(l_mz)
    f["frac_int_query"] = mq.sum() / q_int.sum()
    f["frac_int_lib"] = ml.sum() / l_int.sum()
    f["mse_mz"] = ((q_mz[qi] - l_mz[li]) ** 2).sum() / m
    f["mse_int"] = ((mq - ml) ** 2).sum() / m
    f["contrast_angle"] = 1.0 - 2 * math.acos(min(max(f["cosine"], 0.0), 1.0)) / math.pi
    f["hypergeometric_score"] = hypergeometric(m, len(l_int), bins)
    p = kendalltau_pvalue(mq, ml)
    f["kendalltau"] = 0.0 if math.isnan(p) else (math.inf if p == 0 else -math.log(p))
    abs_diff = np.abs(mq - ml).sum()
    f["ms_for_id_v1"] = min(m ** 4 / (len(q_mz) * len(l_mz) * max(abs_diff, np.finfo(float).eps) ** 0.25), 1000.0)
    f["ms_for_id_v2"] = (m ** 4 * (q_int.sum() + 2 * l_int.sum()) ** 1.25) / (
        (len(q_mz) + 2 * len(l_mz)) ** 2 + abs_diff + np.abs(q_mz[qi] - l_mz[li]).sum())
    merged = np.concatenate([mq + ml, uq, ul]) / 2
    for name, w in (("entropy_unweighted", False), ("entropy_weighted", True)):
        f[name] = 1 - (2 * spectrum_entropy(merged, w) - spectrum_entropy(q_int, w) -
                       spectrum_entropy(l_int, w)) / math.log(4)
    den = ((mq - ml) ** 2).sum() + (ul ** 2).sum()
    f["scribe_fragment_acc"] = 10.0 if den == 0.0 else math.log(1 / den)
    f["manhattan"] = abs_diff + uq.sum() + ul.sum()
    f["euclidean"] = math.sqrt(((mq - ml) ** 2).sum() + (uq ** 2).sum() + (ul ** 2).sum())
    f["chebyshev"] = max(np.abs(mq - ml).max(), uq.max() if len(uq) else 0.0, ul.max() if len(ul) else 0.0)
    x, y = np.concatenate([mq, np.zeros(len(ul))]), np.concatenate([ml, ul])
    f["pearsonr"] = pearson(x, y)
    f["spearmanr"] = pearson(average_ranks(x), average_ranks(y))
    f["braycurtis"] = (abs_diff + uq.sum() + ul.sum()) / (np.abs(mq + ml).sum() + uq.sum() + ul.sum())
    s = mq + ml
    f["canberra"] = (np.where(s != 0, np.abs(mq - ml) / np.where(s != 0, s, 1), 0.0).sum() +
                     np.count_nonzero(uq) + np.count_nonzero(ul))
    f["ruzicka"] = np.minimum(mq, ml).sum() / (np.maximum(mq, ml).sum() + uq.sum() + ul.sum())
    # ---- restricted to the TOP most intense library peaks (spectrum_similarity.py:50-75)
    top = np.argsort(-l_int, kind="stable")[:TOP]
    keep = np.isin(li, top)
    ul_t = l_int[np.intersect1d(lib_unmatched, top)]
    if keep.any():
        mq_t, ml_t = mq[keep], ml[keep]
        with np.errstate(invalid="ignore", divide="ignore"):  # all-zero matched intensities: NaN, like the reference
            f["cosine_top5"] = float(np.float64((mq_t * ml_t).sum()) /
                                     np.float64(math.sqrt((mq_t ** 2).sum()) * math.sqrt((ml_t ** 2).sum())))
        f["frac_n_peaks_lib_top5"] = len(ml_t) / (len(ml_t) + len(ul_t))
        f["frac_int_lib_top5"] = ml_t.sum() / (ml_t.sum() + ul_t.sum())
        f["mse_mz_top5"] = ((q_mz[qi[keep]] - l_mz[li[keep]]) ** 2).sum() / len(ml_t)
        f["mse_int_top5"] = ((mq_t - ml_t) ** 2).sum() / len(ml_t)
        den = ((mq_t - ml_t) ** 2).sum() + (ul_t ** 2).sum()
        f["scribe_fragment_acc_top5"] = 10.0 if den == 0.0 else math.log(1 / den)
        x, y = np.concatenate([mq_t, np.zeros(len(ul_t))]), np.concatenate([ml_t, ul_t])
        f["pearsonr_top5"] = pearson(x, y)
        f["spearmanr_top5"] = pearson(average_ranks(x), average_ranks(y))
    else:
        f["mse_mz_top5"] = f["mse_int_top5"] = math.inf
    f["contrast_angle_top5"] = 1.0 - 2 * math.acos(min(max(f["cosine_top5"], 0.0), 1.0)) / math.pi
    return np.array([f[k] for k in FEATURE_NAMES], np.float64)


def ssm_features_batch(q: dict, lib: dict, lib_row, pairs, n_pairs, q_charge, sequence_len, bins: int,
                       q_mz64=None) -> np.ndarray:
    """Batch over CSR stores (see ann_solo_b200.synth): q/lib dicts with mz, inten, off, prec_mz."""
    out = np.full((len(lib_row), len(FEATURE_NAMES)), np.nan)
    qmz = q["mz"] if q_mz64 is None else q_mz64
    for i, r in enumerate(lib_row):
        if r < 0 or n_pairs[i] <= 0:
            continue
        a, b = q["off"][i], q["off"][i + 1]
        c, d = lib["off"][r], lib["off"][r + 1]
        out[i] = ssm_features(qmz[a:b], q["inten"][a:b], lib["mz"][c:d], lib["inten"][c:d], pairs[i][:n_pairs[i]],
                              float(q["prec_mz"][i]), int(q_charge[i]), float(lib["prec_mz"][r]),
                              0 if sequence_len is None else int(sequence_len[i]), bins)
    return out

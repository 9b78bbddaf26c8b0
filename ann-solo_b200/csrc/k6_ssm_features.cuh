// K6 core: the 44 numeric SSM feature columns of reference utils.py:276-457 (`_compute_ssm_features`)
// for ONE spectrum-spectrum match, as computed by `SpectrumSimilarityCalculator`
// (reference spectrum_similarity.py:13-730) with and without the top-5 library-peak filter.
//
// Plain scalar code over small per-thread arrays, float64 throughout (the reference accumulates
// float32 intensities in float32; see tests/test_ssm_features.py for the agreed tolerance). The
// function is __host__ __device__ so that the -m "not gpu" suite can compile this very source for
// the host (oracle/k6_host_check.cpp, test infrastructure) and check it against the golden vectors;
// the product only ever calls it from k6_ssm_features_kernel.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define K6_HD __host__ __device__ __forceinline__
#else
#define K6_HD inline
#endif

namespace solo {
namespace k6 {

constexpr int MAX_PEAKS = 128;  // peaks per spectrum (library capacity, DESIGN.md "Capacities")
constexpr int N_FEATURES = 44;
constexpr int TOP = 5;          // utils.py:336
constexpr int KENDALL_EXACT_MAX_N = 33;  // scipy.stats.kendalltau method='auto'
constexpr int KENDALL_DP = KENDALL_EXACT_MAX_N * (KENDALL_EXACT_MAX_N - 1) / 4 + 2;

// column order = FEATURE_NAMES (utils.py:296-340 minus index / sequence / is_target)
enum Col {
    C_SEQUENCE_LEN = 0, C_CHARGE_2, C_CHARGE_3, C_CHARGE_4, C_CHARGE_5, C_QUERY_PREC_MZ, C_LIB_PREC_MZ, C_MZ_DIFF_PPM,
    C_ABS_MZ_DIFF_PPM, C_MZ_DIFF_DA, C_ABS_MZ_DIFF_DA, C_COSINE, C_COSINE_TOP5, C_N_MATCHED, C_FRAC_N_QUERY,
    C_FRAC_N_LIB, C_FRAC_N_LIB_TOP5, C_FRAC_INT_QUERY, C_FRAC_INT_LIB, C_FRAC_INT_LIB_TOP5, C_MSE_MZ, C_MSE_MZ_TOP5,
    C_MSE_INT, C_MSE_INT_TOP5, C_CONTRAST, C_CONTRAST_TOP5, C_HYPERGEOMETRIC, C_KENDALLTAU, C_MSFORID_V1, C_MSFORID_V2,
    C_ENTROPY_UNWEIGHTED, C_ENTROPY_WEIGHTED, C_SCRIBE, C_SCRIBE_TOP5, C_MANHATTAN, C_EUCLIDEAN, C_CHEBYSHEV,
    C_PEARSONR, C_PEARSONR_TOP5, C_SPEARMANR, C_SPEARMANR_TOP5, C_BRAYCURTIS, C_CANBERRA, C_RUZICKA
};

struct SsmIn {
    const float *q_mz32;   // exactly one of q_mz32 / q_mz64 is set (the precision the caller holds)
    const double *q_mz64;
    const float *q_int;
    int nq;
    const float *l_mz;
    const float *l_int;
    int nl;
    const uint32_t *pairs;  // np x (query peak, library peak)
    int np;
    double q_prec_mz, l_prec_mz;
    int q_charge, sequence_len;
    int64_t n_peak_bins;    // get_dim(min_mz, max_mz, bin_size)[0] (spectrum_similarity.py:297)
    // log-gamma tables for the binomials of hypergeometric_score (fill_log_tables):
    const double *lfact;    // [MAX_PEAKS + 1]      lfact[k] = lgamma(k + 1)
    const double *lbig;     // [2 * MAX_PEAKS + 1]  lbig[k]  = lgamma(n_peak_bins + 1 - k)
};

constexpr int LFACT_LEN = MAX_PEAKS + 1, LBIG_LEN = 2 * MAX_PEAKS + 1;

// host side: the two tables for a given number of m/z bins
inline void fill_log_tables(int64_t bins, double *lfact, double *lbig) {
    for (int k = 0; k < LFACT_LEN; ++k) lfact[k] = lgamma((double)k + 1.0);
    for (int k = 0; k < LBIG_LEN; ++k) lbig[k] = bins + 1 - k > 0 ? lgamma((double)(bins + 1 - k)) : HUGE_VAL;
}

struct Scratch {  // per thread
    uint8_t qflag[MAX_PEAKS];  // bit 0: matched
    uint8_t lflag[MAX_PEAKS];  // bit 0: matched, bit 1: among the TOP most intense library peaks
    double x[MAX_PEAKS], y[MAX_PEAKS], rx[2 * MAX_PEAKS], ry[MAX_PEAKS];  // rx also holds the merged spectrum
    double dp[KENDALL_DP];
    double lv[2 * MAX_PEAKS];  // per-element logs of the spectrum whose entropy is being taken
};

K6_HD double k6_inf() { return HUGE_VAL; }
K6_HD double k6_nan() { return HUGE_VAL - HUGE_VAL; }

// spectrum_similarity.py:251-309: -log P(more than `m` of the `L` library peaks match by chance);
// comb(n, k) through the log-gamma tables: comb(L, i), comb(bins - L, L - i), comb(bins, L)
K6_HD double hypergeometric_score(int m, int L, int64_t bins, const double *lfact, const double *lbig) {
    double p = 0.0;
    const double ld = lbig[0] - lfact[L] - lbig[L];
    for (int i = m + 1; i <= L; ++i) {
        if ((int64_t)(L - i) > bins - L) continue;  // comb() == 0
        const double a = lfact[L] - lfact[i] - lfact[L - i];
        const double b = lbig[L] - lfact[L - i] - lbig[2 * L - i];
        p += exp(a + b - ld);
    }
    if (!(p > 0.0)) return 100.0;  // guard against infinity for identical spectra (:308)
    const double s = -log(p);
    return s < 100.0 ? s : 100.0;
}

K6_HD double factorial(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f *= (double)i;
    return f;
}

// scipy.stats._mstats_basic._kendall_p_exact, two-sided, n < 171
K6_HD double kendall_p_exact(int n, int64_t c, double *dp) {
    const int64_t tot = (int64_t)n * (n - 1) / 2;
    if (tot - c < c) c = tot - c;
    double prob;
    if (n == 1 || n == 2) prob = 1.0;
    else if (c == 0) prob = 2.0 / factorial(n);
    else if (c == 1) prob = 2.0 / factorial(n - 1);
    else if (4 * c == (int64_t)n * (n - 1)) prob = 1.0;
    else {
        const int C = (int)c;  // only reached with n <= KENDALL_EXACT_MAX_N: C + 1 <= KENDALL_DP
        for (int t = 0; t <= C; ++t) dp[t] = t < 2 ? 1.0 : 0.0;
        for (int j = 3; j <= n; ++j) {
            for (int t = 1; t <= C; ++t) dp[t] += dp[t - 1];          // new = cumsum(new)
            if (j <= C)
                for (int t = C; t >= j; --t) dp[t] -= dp[t - j];       // new[j:] -= new[:c+1-j] (old values)
        }
        double s = 0.0;
        for (int t = 0; t <= C; ++t) s += dp[t];
        prob = 2.0 * s / factorial(n);
    }
    return prob < 0.0 ? 0.0 : (prob > 1.0 ? 1.0 : prob);
}

// scipy.stats.kendalltau(x, y)[1]: tau-b, method 'auto', two-sided; NaN when undefined
K6_HD double kendalltau_pvalue(const double *x, const double *y, int n, double *dp) {
    if (n <= 0) return k6_nan();
    int64_t dis = 0, xtie = 0, ytie = 0, ntie = 0;
    double x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    for (int i = 0; i < n; ++i) {
        int cx = 0, cy = 0;
        for (int j = 0; j < n; ++j) {
            const bool ex = x[i] == x[j], ey = y[i] == y[j];
            cx += ex;
            cy += ey;
            if (j > i) {
                xtie += ex;
                ytie += ey;
                ntie += ex && ey;
                dis += (x[i] < x[j] && y[i] > y[j]) || (x[i] > x[j] && y[i] < y[j]);
            }
        }
        // sums over tie groups of cnt(cnt-1)(cnt-2) and cnt(cnt-1)(2cnt+5), element by element
        x0 += (double)(cx - 1) * (cx - 2);
        x1 += (double)(cx - 1) * (2 * cx + 5);
        y0 += (double)(cy - 1) * (cy - 2);
        y1 += (double)(cy - 1) * (2 * cy + 5);
    }
    const int64_t tot = (int64_t)n * (n - 1) / 2;
    if (xtie == tot || ytie == tot) return k6_nan();
    const int64_t con_minus_dis = tot - xtie - ytie + ntie - 2 * dis;
    const int64_t lo = dis < tot - dis ? dis : tot - dis;
    if (xtie == 0 && ytie == 0 && (n <= KENDALL_EXACT_MAX_N || lo <= 1)) return kendall_p_exact(n, tot - dis, dp);
    const double m = (double)n * ((double)n - 1.0);
    const double var = (m * (2.0 * n + 5.0) - x1 - y1) / 18.0 + (2.0 * (double)xtie * (double)ytie) / m +
                       x0 * y0 / (9.0 * m * ((double)n - 2.0));
    const double z = (double)con_minus_dis / sqrt(var);
    return erfc(fabs(z) / 1.4142135623730951);  // 2 * norm.sf(|z|)
}

// scipy.stats.pearsonr(x, y)[0]; 0.0 where SciPy yields NaN (constant input), spectrum_similarity.py:487
K6_HD double pearson(const double *x, const double *y, int n) {
    if (n < 2) return 0.0;
    bool cx = true, cy = true;
    double sx = 0, sy = 0;
    for (int i = 0; i < n; ++i) {
        cx &= x[i] == x[0];
        cy &= y[i] == y[0];
        sx += x[i];
        sy += y[i];
    }
    if (cx || cy) return 0.0;
    const double mx = sx / n, my = sy / n;
    double sxy = 0, sxx = 0, syy = 0;
    for (int i = 0; i < n; ++i) {
        const double a = x[i] - mx, b = y[i] - my;
        sxy += a * b;
        sxx += a * a;
        syy += b * b;
    }
    double r = sxy / (sqrt(sxx) * sqrt(syy));
    if (!(r == r)) return 0.0;
    return r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
}

K6_HD void average_ranks(const double *v, int n, double *r) {
    for (int i = 0; i < n; ++i) {
        int less = 0, equal = 0;
        for (int j = 0; j < n; ++j) {
            less += v[j] < v[i];
            equal += v[j] == v[i];
        }
        r[i] = (double)less + ((double)equal + 1.0) * 0.5;
    }
}

// spectrum_similarity.py:703-730 (_spectrum_entropy) for both variants at once: scipy.stats.entropy of
// v (unweighted) and, when that is <= 3, of v ** (0.25 + 0.25 * entropy) (weighted; above 3 the two
// coincide). log v is taken once per element (lv) and reused: v ** w = exp(w * lv).
K6_HD void spectrum_entropies(const double *v, int n, double *lv, double &unweighted, double &weighted) {
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += v[i];
    const double lsum = log(sum);
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        lv[i] = v[i] > 0.0 ? log(v[i]) : 0.0;
        if (v[i] > 0.0) s -= (v[i] / sum) * (lv[i] - lsum);
    }
    unweighted = s;
    weighted = s;
    if (s > 3.0) return;
    const double w = 0.25 + 0.25 * s;
    double wsum = 0.0;
    for (int i = 0; i < n; ++i)
        if (v[i] > 0.0) wsum += exp(w * lv[i]);
    const double lwsum = log(wsum);
    double sw = 0.0;
    for (int i = 0; i < n; ++i)
        if (v[i] > 0.0) {
            const double lp = w * lv[i] - lwsum;
            sw -= exp(lp) * lp;
        }
    weighted = sw;
}

K6_HD double contrast_angle(double cosine) {  // spectrum_similarity.py:233-249
    const double c = cosine < 0.0 ? 0.0 : (cosine > 1.0 ? 1.0 : cosine);
    return 1.0 - 2.0 * acos(c) / 3.141592653589793;
}

K6_HD double scribe(double den) {  // spectrum_similarity.py:632-657 (math.isclose(x, 0.0) <=> x == 0)
    return den == 0.0 ? 10.0 : log(1.0 / den);
}

K6_HD void ssm_features(const SsmIn &in, Scratch &S, double *out) {
    const int nq = in.nq, nl = in.nl, np = in.np;
    for (int c = 0; c < N_FEATURES; ++c) out[c] = 0.0;
    out[C_SEQUENCE_LEN] = (double)in.sequence_len;
    out[in.q_charge <= 2 ? C_CHARGE_2 : in.q_charge == 3 ? C_CHARGE_3 : in.q_charge == 4 ? C_CHARGE_4 : C_CHARGE_5] = 1.0;
    out[C_QUERY_PREC_MZ] = in.q_prec_mz;
    out[C_LIB_PREC_MZ] = in.l_prec_mz;
    const double da = in.q_prec_mz - in.l_prec_mz;            // spectrum_utils mass_diff(mz1, mz2, True)
    const double ppm = da / in.l_prec_mz * 1000000.0;         // ... (mz1, mz2, False)
    out[C_MZ_DIFF_DA] = da;
    out[C_ABS_MZ_DIFF_DA] = fabs(da);
    out[C_MZ_DIFF_PPM] = ppm;
    out[C_ABS_MZ_DIFF_PPM] = fabs(ppm);

    for (int i = 0; i < nq; ++i) S.qflag[i] = 0;
    for (int i = 0; i < nl; ++i) S.lflag[i] = 0;
    // the TOP most intense library peaks (np.argpartition(int_library, -top)[-top:]); equal
    // intensities at the cut are resolved towards the lower peak index
    for (int t = 0; t < TOP && t < nl; ++t) {
        int best = -1;
        for (int i = 0; i < nl; ++i)
            if (!(S.lflag[i] & 2) && (best < 0 || in.l_int[i] > in.l_int[best])) best = i;
        S.lflag[best] |= 2;
    }
    double q_sum = 0, l_sum = 0;
    for (int i = 0; i < nq; ++i) q_sum += (double)in.q_int[i];
    for (int i = 0; i < nl; ++i) l_sum += (double)in.l_int[i];

    // ---- sums over the matched peaks, all and top-filtered
    double dot = 0, mq_sum = 0, ml_sum = 0, abs_diff = 0, sq_diff = 0, abs_sum = 0, mz_sq = 0, mz_abs = 0, max_diff = 0,
           canb = 0, min_sum = 0, max_sum = 0;
    double t_dot = 0, t_qq = 0, t_ll = 0, t_ml_sum = 0, t_sq_diff = 0, t_mz_sq = 0;
    int t_m = 0;
    for (int k = 0; k < np; ++k) {
        const int qi = (int)in.pairs[2 * k], li = (int)in.pairs[2 * k + 1];
        S.qflag[qi] |= 1;
        S.lflag[li] |= 1;
        const double a = (double)in.q_int[qi], b = (double)in.l_int[li];
        const double qm = in.q_mz64 ? in.q_mz64[qi] : (double)in.q_mz32[qi];
        const double dm = qm - (double)in.l_mz[li], d = a - b;
        dot += a * b;
        mq_sum += a;
        ml_sum += b;
        abs_diff += fabs(d);
        sq_diff += d * d;
        abs_sum += fabs(a + b);
        mz_sq += dm * dm;
        mz_abs += fabs(dm);
        if (fabs(d) > max_diff) max_diff = fabs(d);
        if (a + b != 0.0) canb += fabs(d) / (a + b);   // nan_to_num(0/0) = 0 (:606-612)
        min_sum += a < b ? a : b;
        max_sum += a > b ? a : b;
        S.x[k] = a;
        S.y[k] = b;
        if (S.lflag[li] & 2) {
            ++t_m;
            t_dot += a * b;
            t_qq += a * a;
            t_ll += b * b;
            t_ml_sum += b;
            t_sq_diff += d * d;
            t_mz_sq += dm * dm;
        }
    }
    // ---- unmatched peaks
    double uq_sum = 0, uq_sq = 0, uq_max = 0, ul_sum = 0, ul_sq = 0, ul_max = 0, t_ul_sum = 0, t_ul_sq = 0;
    int uq_nz = 0, ul_nz = 0, n_ul = 0, t_n_ul = 0;
    for (int i = 0; i < nq; ++i)
        if (!(S.qflag[i] & 1)) {
            const double v = (double)in.q_int[i];
            uq_sum += v;
            uq_sq += v * v;
            if (v > uq_max) uq_max = v;
            uq_nz += v != 0.0;
        }
    for (int i = 0; i < nl; ++i)
        if (!(S.lflag[i] & 1)) {
            const double v = (double)in.l_int[i];
            ++n_ul;
            ul_sum += v;
            ul_sq += v * v;
            if (v > ul_max) ul_max = v;
            ul_nz += v != 0.0;
            if (S.lflag[i] & 2) {
                ++t_n_ul;
                t_ul_sum += v;
                t_ul_sq += v * v;
            }
        }
    const double m = (double)np;
    out[C_COSINE] = dot;                                   // :88-106 (norm = 1 without the top filter)
    out[C_N_MATCHED] = m;
    out[C_FRAC_N_QUERY] = m / (double)nq;
    out[C_FRAC_N_LIB] = m / (double)nl;
    out[C_FRAC_INT_QUERY] = mq_sum / q_sum;
    out[C_FRAC_INT_LIB] = ml_sum / l_sum;
    out[C_MSE_MZ] = mz_sq / m;
    out[C_MSE_INT] = sq_diff / m;
    out[C_CONTRAST] = contrast_angle(dot);
    out[C_HYPERGEOMETRIC] = hypergeometric_score(np, nl, in.n_peak_bins, in.lfact, in.lbig);
    {
        const double p = kendalltau_pvalue(S.x, S.y, np, S.dp);
        out[C_KENDALLTAU] = p == p ? -log(p) : 0.0;        // :334-338
    }
    {
        const double eps = 2.220446049250313e-16;
        const double v1 = m * m * m * m / ((double)nq * (double)nl * pow(abs_diff > eps ? abs_diff : eps, 0.25));
        out[C_MSFORID_V1] = v1 < 1000.0 ? v1 : 1000.0;     // :340-378
        const double npk = (double)nq + 2.0 * (double)nl;
        out[C_MSFORID_V2] = m * m * m * m * pow(q_sum + 2.0 * l_sum, 1.25) / (npk * npk + abs_diff + mz_abs);  // :380-413
    }
    out[C_SCRIBE] = scribe(sq_diff + ul_sq);
    out[C_MANHATTAN] = abs_diff + uq_sum + ul_sum;
    out[C_EUCLIDEAN] = sqrt(sq_diff + uq_sq + ul_sq);
    out[C_CHEBYSHEV] = max_diff > uq_max ? (max_diff > ul_max ? max_diff : ul_max) : (uq_max > ul_max ? uq_max : ul_max);
    out[C_BRAYCURTIS] = (abs_diff + uq_sum + ul_sum) / (abs_sum + uq_sum + ul_sum);
    out[C_CANBERRA] = canb + (double)uq_nz + (double)ul_nz;
    out[C_RUZICKA] = min_sum / (max_sum + uq_sum + ul_sum);

    // ---- correlations over (matched, unmatched library) peaks: x = [mq, 0...], y = [ml, ul] (:476-520)
    int n = np;
    for (int i = 0; i < nl; ++i)
        if (!(S.lflag[i] & 1)) {
            S.x[n] = 0.0;
            S.y[n] = (double)in.l_int[i];
            ++n;
        }
    out[C_PEARSONR] = pearson(S.x, S.y, n);
    average_ranks(S.x, n, S.rx);
    average_ranks(S.y, n, S.ry);
    out[C_SPEARMANR] = pearson(S.rx, S.ry, n);

    // ---- spectral entropy (:659-700): merged = [(mq + ml) / 2, uq / 2, ul / 2]
    {
        int k = 0;
        for (int j = 0; j < np; ++j)
            S.rx[k++] = ((double)in.q_int[in.pairs[2 * j]] + (double)in.l_int[in.pairs[2 * j + 1]]) * 0.5;
        for (int i = 0; i < nq; ++i)
            if (!(S.qflag[i] & 1)) S.rx[k++] = (double)in.q_int[i] * 0.5;
        for (int i = 0; i < nl; ++i)
            if (!(S.lflag[i] & 1)) S.rx[k++] = (double)in.l_int[i] * 0.5;
        double mu, mw, qu, qw, lu, lw;
        spectrum_entropies(S.rx, k, S.lv, mu, mw);
        for (int i = 0; i < nq; ++i) S.ry[i] = (double)in.q_int[i];
        spectrum_entropies(S.ry, nq, S.lv, qu, qw);
        for (int i = 0; i < nl; ++i) S.ry[i] = (double)in.l_int[i];
        spectrum_entropies(S.ry, nl, S.lv, lu, lw);
        out[C_ENTROPY_UNWEIGHTED] = 1.0 - (2.0 * mu - qu - lu) / 1.3862943611198906;
        out[C_ENTROPY_WEIGHTED] = 1.0 - (2.0 * mw - qw - lw) / 1.3862943611198906;
    }

    // ---- restricted to the TOP most intense library peaks (:50-75)
    if (t_m > 0) {
        out[C_COSINE_TOP5] = t_dot / (sqrt(t_qq) * sqrt(t_ll));
        out[C_FRAC_N_LIB_TOP5] = (double)t_m / (double)(t_m + t_n_ul);
        out[C_FRAC_INT_LIB_TOP5] = t_ml_sum / (t_ml_sum + t_ul_sum);
        out[C_MSE_MZ_TOP5] = t_mz_sq / (double)t_m;
        out[C_MSE_INT_TOP5] = t_sq_diff / (double)t_m;
        out[C_SCRIBE_TOP5] = scribe(t_sq_diff + t_ul_sq);
        n = 0;
        for (int k = 0; k < np; ++k) {
            const int qi = (int)in.pairs[2 * k], li = (int)in.pairs[2 * k + 1];
            if (S.lflag[li] & 2) {
                S.x[n] = (double)in.q_int[qi];
                S.y[n] = (double)in.l_int[li];
                ++n;
            }
        }
        for (int i = 0; i < nl; ++i)
            if ((S.lflag[i] & 3) == 2) {
                S.x[n] = 0.0;
                S.y[n] = (double)in.l_int[i];
                ++n;
            }
        out[C_PEARSONR_TOP5] = pearson(S.x, S.y, n);
        average_ranks(S.x, n, S.rx);
        average_ranks(S.y, n, S.ry);
        out[C_SPEARMANR_TOP5] = pearson(S.rx, S.ry, n);
    } else {
        out[C_MSE_MZ_TOP5] = k6_inf();   // :205-231
        out[C_MSE_INT_TOP5] = k6_inf();
    }
    out[C_CONTRAST_TOP5] = contrast_angle(out[C_COSINE_TOP5]);
}

}  // namespace k6
}  // namespace solo

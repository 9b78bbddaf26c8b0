// K6: SSM feature table for rescoring. Replaces the per-SSM Python loop of
// reference utils.py:330-455 (two SpectrumSimilarityCalculator objects and ~40 NumPy/SciPy calls per
// SSM) by one launch over all SSMs of a batch: one thread per SSM, every peak of the two spectra
// read once from the device-resident stores, 44 float64 columns written per SSM.
#include "k6_ssm_features.cuh"
#include "solo_common.cuh"

namespace solo {

constexpr int K6_THREADS = 64;

__global__ void __launch_bounds__(K6_THREADS)
k6_ssm_features_kernel(const FeatureArgs a) {
    const int i = blockIdx.x * K6_THREADS + threadIdx.x;
    if (i >= a.n) return;
    double *out = a.out + (int64_t)i * k6::N_FEATURES;
    const int row = a.lib_row[i];
    const int np = a.n_pairs[i];
    if (row < 0 || np <= 0) {  // the reference skips SSMs without peak matches (utils.py:332-333)
        const double nan = k6::k6_nan();
        for (int c = 0; c < k6::N_FEATURES; ++c) out[c] = nan;
        return;
    }
    const int64_t qb = a.q_off[i], lb = a.l_off[row];
    k6::SsmIn in;
    in.q_mz32 = a.q_mz64 ? nullptr : a.q_mz32 + qb;
    in.q_mz64 = a.q_mz64 ? a.q_mz64 + qb : nullptr;
    in.q_int = a.q_int + qb;
    in.nq = (int)(a.q_off[i + 1] - qb);
    in.l_mz = a.l_mz + lb;
    in.l_int = a.l_int + lb;
    in.nl = (int)(a.l_off[row + 1] - lb);
    in.pairs = a.pairs + (int64_t)i * a.max_pairs * 2;
    in.np = np;
    in.q_prec_mz = a.q_prec_mz[i];
    in.l_prec_mz = a.l_prec_mz[row];
    in.q_charge = a.q_charge ? a.q_charge[i] : a.q_charge_all;
    in.sequence_len = a.sequence_len ? a.sequence_len[i] : 0;
    in.n_peak_bins = a.n_peak_bins;
    in.lfact = a.lfact;
    in.lbig = a.lbig;
    bool ok = in.nq <= k6::MAX_PEAKS && in.nl <= k6::MAX_PEAKS && np <= k6::MAX_PEAKS && np <= a.max_pairs;
    for (int k = 0; ok && k < np; ++k)  // a pair naming a peak outside either spectrum is refused, never read
        ok = in.pairs[2 * k] < (uint32_t)in.nq && in.pairs[2 * k + 1] < (uint32_t)in.nl;
    if (!ok) {
        atomicAdd(a.bad, 1);
        return;
    }
    k6::Scratch S;
    k6::ssm_features(in, S, out);
}

void launch_ssm_features(solo_handle *h, const FeatureArgs &a) {
    if (a.n <= 0) return;
    k6_ssm_features_kernel<<<div_up(a.n, K6_THREADS), K6_THREADS, 0, h->stream>>>(a);
    SOLO_CUDA(cudaGetLastError());
    h->launches++;
}

}  // namespace solo

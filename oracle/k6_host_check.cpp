// TEST INFRASTRUCTURE ONLY. Compiles the K6 device function's own source
// (ann-solo_b200/csrc/k6_ssm_features.cuh, __host__ __device__) for the host, so that the
// -m "not gpu" suite can check the kernel's arithmetic against the golden feature vectors without a
// GPU. Nothing in the product loads this library; the product path is k6_ssm_features_kernel only.
#include "../ann-solo_b200/csrc/k6_ssm_features.cuh"

extern "C" int k6_host_n_features() { return solo::k6::N_FEATURES; }

extern "C" int k6_host_ssm_features(const float *q_mz32, const double *q_mz64, const float *q_int, int nq,
                                    const float *l_mz, const float *l_int, int nl, const uint32_t *pairs, int np,
                                    double q_prec_mz, double l_prec_mz, int q_charge, int sequence_len,
                                    int64_t n_peak_bins, double *out) {
    using namespace solo::k6;
    if (nq > MAX_PEAKS || nl > MAX_PEAKS || np > MAX_PEAKS || np <= 0) return -1;
    static thread_local double lfact[LFACT_LEN], lbig[LBIG_LEN];
    static thread_local int64_t tables_for = -1;
    if (tables_for != n_peak_bins) {
        fill_log_tables(n_peak_bins, lfact, lbig);
        tables_for = n_peak_bins;
    }
    SsmIn in{q_mz64 ? nullptr : q_mz32, q_mz64, q_int, nq, l_mz, l_int, nl, pairs, np, q_prec_mz, l_prec_mz,
             q_charge, sequence_len, n_peak_bins, lfact, lbig};
    static thread_local Scratch S;
    ssm_features(in, S, out);
    return 0;
}

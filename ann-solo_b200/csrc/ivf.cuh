// IVF-Flat inner-product index: declarations shared by ivf.cu / ivf_tc.cu / solo_api.cu.
#pragma once
#include <functional>

#include "solo_common.cuh"

namespace solo {

constexpr int IVF_MAX_K = 2048;        // top-k rows returned per query
constexpr int IVF_MAX_NLIST = 32768;   // coarse scores of one query are selected in shared memory
constexpr int64_t IVF_ROUND0_SCORES = 24576;  // longest inverted list: round 0 appends a whole list unconditionally and the per-query buffer holds 32,768 entries
constexpr unsigned long long IVF_PACKED_PAD = 0xFF800000FFFFFFFFull;  // score -inf, row -1
constexpr float IVF_REL_EPS = 1.25e-3f;  // bound on |approx - exact| / sum|q_d c_d| for the fp16 tensor path

// composite selection key: larger is better; (score desc, id asc) is a strict total order
__host__ __device__ __forceinline__ uint32_t ivf_f2o(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ivf_o2f(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

struct IvfSearchScratch;

// One work item of the tcgen05 scan = one chunk of <= NB consecutive vectors of an inverted list, scored against
// the list's whole query group of this round.
struct __align__(16) TcItem {
    int32_t p0;   // first list position of the chunk
    int32_t nv;   // vectors in the chunk (1..NB)
    int32_t g0;   // offset of the list's query group in gq
    int32_t G;    // queries in the group
};

// Everything the search needs, all device pointers.
struct IvfSearchArgs {
    const float *q;      // (nq, dim) fp32 query vectors
    int nq;
    int k;
    int nprobe;
    // outputs (any may be null)
    int64_t *I;          // (nq, k) sorted (score desc, id asc), -1 padded
    float *D;            // (nq, k) exact fp32 scores, -inf padded
    unsigned long long *packed;  // (nq, k) unsorted (float score bits << 32 | row), IVF_PACKED_PAD padded (mode B exchange)
    int32_t *sel_ids;    // (nq, k) unsorted selected rows, first sel_cnt[q] valid
    int32_t *sel_cnt;    // (nq)
    int32_t *probes;     // (nq, nprobe) selected lists (unsorted set unless sort_probes)
    int sort_probes;
    int coarse_only;
    const int32_t *given_probes;  // (nq, nprobe) device: skip coarse scoring / probe selection (mode B: another GPU selected them)
    // optional precursor-window mask fused into the (unsorted) selection output; tol_mode < 0: off
    const double *win_q_prec_mz;
    const float *win_lib_prec_mz32;
    const uint8_t *win_lib_valid;
    int win_charge;
    double win_tol;
    int win_tol_mode;
};

// exclusive scan of n int32 counts into n+1 int64 offsets (single CTA)
void scan_counts_i32(solo_handle *h, const int32_t *cnt, int64_t n, int64_t *off);
// tcgen05 scan engine (ivf_tc.cu)
bool tc_scan_supported(const IvfIndex &ix);
void tc_make_tensor_map(IvfIndex &ix);
void tc_make_centroid_map(IvfIndex &ix);
void launch_coarse_tc(solo_handle *h, IvfIndex &ix, const __half *qh, const uint32_t *qmask, int nq, int q_scale_log2,
                      float *out, int ld, int n_lists = -1, const float *tau = nullptr, unsigned long long *buf = nullptr,
                      int32_t *cnt = nullptr, int cap = 0);
void launch_coarse_tc_listed(solo_handle *h, IvfIndex &ix, const __half *qh, const uint32_t *qmask, int q_scale_log2,
                             float *out, int ld, const int32_t *q_list, const int32_t *n_listed);
void tc_prepare_queries(solo_handle *h, const IvfIndex &ix, const float *q, int nq, int q_scale_log2, __half *qh,
                        uint32_t *qmask);
void launch_scan_tc(solo_handle *h, IvfIndex &ix, const int64_t *goff, const int32_t *gq, const __half *qh,
                    const uint32_t *qmask, int q_scale_log2, const float *tau, unsigned long long *buf, int32_t *cnt,
                    int cap, DevBuf &item_cnt, DevBuf &item_off, DevBuf &items, bool dense_round = false);
void ivf_set_centroids(solo_handle *h, IvfIndex &ix, const float *h_cent, int nlist, int dim);
void ivf_add_device(solo_handle *h, IvfIndex &ix, const float *d_x, int64_t n, bool assign = true);  // d_x on device
void ivf_train_rows(solo_handle *h, IvfIndex &ix, int64_t n, int dim, int nlist, int iters, uint64_t seed,
                    const std::function<void(int64_t, int64_t, float *)> &fill);
void ivf_finalize(solo_handle *h, IvfIndex &ix);
void ivf_reset(IvfIndex &ix);
void ivf_search(solo_handle *h, IvfIndex &ix, const IvfSearchArgs &a);
void ivf_merge_select(solo_handle *h, IvfIndex &ix, const unsigned long long *d_parts, int n_parts, int S, int k,
                      const float *d_q_slice, const float *d_qnorm_slice, int n, const IvfSearchArgs &win,
                      int32_t *sel_ids, int32_t *sel_cnt);
void ivf_train(solo_handle *h, IvfIndex &ix, const float *h_x, int64_t n, int dim, int nlist, int iters,
               uint64_t seed);
// faiss_io.cu: Faiss ".idxann" files, explicit list assignment, dense read-back
void ivf_add_assigned(solo_handle *h, IvfIndex &ix, const float *h_x, int64_t n, const int32_t *list_of_row);
void ivf_read_index(solo_handle *h, IvfIndex &ix, const char *path, int64_t *nprobe_out);
void ivf_write_index(solo_handle *h, IvfIndex &ix, const char *path, int64_t nprobe);
void ivf_reconstruct(solo_handle *h, IvfIndex &ix, int64_t row0, int64_t n, float *h_out);
void idxann_inspect(const char *path, solo_idxann_info *info);

}  // namespace solo

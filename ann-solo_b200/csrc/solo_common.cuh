// Shared host/device declarations for libsolo_b200.so (sm_100a only).
#pragma once
#include <nvtx3/nvToolsExt.h>

#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/solo_b200.h"

namespace solo {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------- errors
struct Error {
    int code;
    std::string msg;
};

#define SOLO_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            char _b[512];                                                                       \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                     __FILE__, __LINE__);                                                       \
            throw ::solo::Error{_e == cudaErrorMemoryAllocation ? SOLO_ENOMEM : SOLO_ECUDA, _b}; \
        }                                                                                       \
    } while (0)

#define SOLO_REQUIRE(cond, code, ...)              \
    do {                                           \
        if (!(cond)) {                             \
            char _b[512];                          \
            snprintf(_b, sizeof _b, __VA_ARGS__);  \
            throw ::solo::Error{code, _b};         \
        }                                          \
    } while (0)

// ---------------------------------------------------------------- device buffers
// Grow-only device allocation owned by the handle.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            char b[256];
            snprintf(b, sizeof b, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            throw Error{SOLO_ENOMEM, b};
        }
        cap = want;
    }
    // grow while keeping the first `keep` bytes (device-to-device copy on `stream`)
    void ensure_keep(size_t bytes, size_t keep, cudaStream_t stream) {
        if (bytes <= cap) return;
        void *old = p;
        size_t want = bytes + bytes / 2 + 256;
        void *np = nullptr;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) {
            char b[256];
            snprintf(b, sizeof b, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            throw Error{SOLO_ENOMEM, b};
        }
        if (old && keep) {
            cudaMemcpyAsync(np, old, keep, cudaMemcpyDeviceToDevice, stream);
            cudaStreamSynchronize(stream);
        }
        if (old) cudaFree(old);
        p = np;
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

// ---------------------------------------------------------------- profiling stages
enum Stage {
    ST_VECTORIZE = 0,  // K1
    ST_COARSE,         // K2 coarse scoring
    ST_PROBE_SELECT,   // K2 nprobe selection
    ST_GROUP,          // invert (query, probe) -> per-list query groups
    ST_SCAN,           // K3 list scan (dominant kernel)
    ST_TOPK,           // K4 threshold / band re-rank / top-k
    ST_CANDIDATES,     // window mask + compaction
    ST_SCORE,          // K5 shifted dot
    ST_H2D,
    ST_D2H,
    ST_COUNT
};

struct StageProf {
    double ms = 0.0;
    int64_t launches = 0;
    int64_t k5_overflow_pairs = 0;  // (query, candidate) pairs with > 256 tentative matches seen so far
    double units = 0.0;  // stage-specific unit count (bytes or flops), summed
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

// ---------------------------------------------------------------- library peak store
struct LibraryStore {
    int64_t n = 0;
    int64_t n_peaks = 0;
    int max_peaks = 0;
    DevBuf mz, inten, chg, off, prec_mz, prec_mz32, prec_z, valid;
    DevBuf sorted_mz32, sorted_row;  // m/z-sorted view (float32 precursor m/z, library row) for window-only candidates
    DevBuf meta, table;  // scorer fast path: packed per-row metadata and m/z bucket tables (k5_build_aux)
};

// ---------------------------------------------------------------- IVF index
struct IvfIndex {
    int dim = 0;
    int nlist = 0;
    int64_t ntotal = 0;     // rows offered to add() so far (ids run 0..ntotal-1)
    int64_t nstored = 0;    // rows actually stored (NaN rows are skipped)
    bool nonneg = true;     // every stored value and centroid value is >= 0
    float max_norm = 0.f;   // max L2 norm over stored rows
    float cent_max_norm = 0.f;
    bool cent_nonneg = true;  // every centroid value is >= 0
    int scale_log2 = 10;    // fp16 copies hold x * 2^scale_log2
    // centroids
    DevBuf cent;            // (nlist, dim) fp32 row-major
    DevBuf cent_h;          // (nlist, dim) fp16 scaled
    // inverted lists (rebuilt on add): positions are list-ordered
    DevBuf list_off;        // int64 [nlist+1]
    DevBuf list_ids;        // int32 [nstored] library row of position p
    DevBuf vec_h;           // (nstored, dim) fp16 scaled, list order
    DevBuf sp_off;          // int64 [nstored+1] sparse row offsets, list order
    DevBuf sp_idx;          // uint16 [nnz]
    DevBuf sp_val;          // float  [nnz]
    DevBuf row_list;        // int32 [ntotal] list of row (-1 = skipped)
    std::vector<int64_t> h_list_off;  // host copy
    std::vector<uint8_t> owned;       // mode B (lists sharded over GPUs): owned[l] == 0 -> list l is not stored here; empty = all
    // every row ever added, in insertion order, as sparse rows (lists are rebuilt from these)
    DevBuf row_off;         // int64 [ntotal+1]
    DevBuf row_idx;         // uint16 [nnz]
    DevBuf row_val;         // float  [nnz]
    DevBuf stats;           // int32 [4]: {any_negative, max_norm_bits, cent_any_negative, cent_max_norm_bits}
    int64_t nnz = 0;
    bool dirty = true;      // lists need rebuilding before the next search
    int64_t max_list_len = 0;
    int buf_cap = 0;        // per-query candidate buffer capacity used by the scan
    float last_eps_rel = 0.f;   // error model of the scores the last search produced (the mode-B merge re-uses it)
    int last_eps_nonneg = 0;
    // TMA descriptor (CUtensorMap, 128 bytes) of vec_h for the tcgen05 scan engine
    alignas(64) unsigned char tmap_storage[128];
    bool tmap_valid = false;
    alignas(64) unsigned char tmap16_storage[128];     // vec_h again with 16-row boxes (CTA-pair scan kernel)
    alignas(64) unsigned char tmap_cent_storage[128];  // the same for cent_h (tensor-core coarse quantizer)
    bool tmap_cent_valid = false;
    int cent_scale_log2 = 10;  // cent_h holds centroid * 2^cent_scale_log2
    DevBuf coarse_items;       // work items of the tensor-core coarse pass, valid for coarse_items_nq queries
    int coarse_items_nq = -1;
    int coarse_items_used = -1;
    DevBuf coarse_fb_items;    // the same items with a device-written group size (fall-back of the compact probe selection)
};

}  // namespace solo

struct solo_handle {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int64_t launches = 0;
    int64_t k5_overflow_pairs = 0;  // (query, candidate) pairs with > 256 tentative matches seen so far
    bool profile = false;
    bool opt_scan_exact = false;  // solo_set_option("scan_engine", 1): CUDA-core exact list scan
    bool opt_scan_wide = false;   // solo_set_option("scan_wide", 1): list vectors streamed per tile, chunks of up to 256 rows
    int opt_scan_hybrid = 0;      // > 0: lists longer than this use the streamed-chunk scan variant
    int opt_scan_ts = 0;          // solo_set_option("scan_ts", 96 | 112): swapped-operand list scan (list chunk in tensor memory), queries per tile
    int opt_round0_wide = 0;      // > 0: the first scan round streams list chunks of this many rows through the ring
    bool opt_nvtx = getenv("SOLO_NVTX") != nullptr && getenv("SOLO_NVTX")[0] == '1';   // NVTX range per stage
    bool opt_sort_items = true;   // scan items ordered by their tile count, largest first (balance over the persistent CTAs)
    int opt_tc_stages = 0;        // > 0: cap on the query stages of scan_tc_kernel (tuning)
    int opt_tc_debug = 0;         // timing experiments only (solo_set_option("tc_debug", bits)): see TcScanArgs::debug
    bool opt_scan_pairs = false;  // solo_set_option("scan_pairs", 1): cta_group::2 list scan
    int opt_round0_scores = 4096;   // scores per query appended unconditionally by the first scan round
    bool opt_front_probes = true;   // probe selection writes the closest lists first
    int opt_train_balance = 0;      // balancing rounds after the Lloyd iterations of ivf_train (0: plain Lloyd)
    bool opt_compact_probes = true; // thresholded coarse pass + compact probe selection (no (Q, nlist) score matrix)
    int64_t compact_probe_batches = 0;
    int opt_tc_nb = 0;              // > 0: rows of the resident list chunk of the tcgen05 scan (multiple of 32; tuning)
    int opt_tc_kbb = 2;             // k-blocks per MMA issue batch of the tcgen05 scan (tuning)
    solo::StageProf prof[solo::ST_COUNT];

    // vectoriser
    double min_mz = 11.0, max_mz = 2010.0, bin_size = 0.04;
    int hash_len = 800;
    int64_t n_bins = 0;
    double min_bound = 0.0;
    solo::DevBuf lut;  // uint16 [n_bins + 2]
    std::vector<uint16_t> h_lut;

    std::map<int, solo::LibraryStore> libs;
    std::map<int, solo::IvfIndex> ivf;

    // staged query batch
    int nq = 0;
    int64_t q_peaks = 0;
    int q_max_peaks = 0;
    int q_mz_is_f64 = 0;
    solo::DevBuf q_mz, q_mz_vec, q_int, q_off, q_prec_mz;

    // scratch
    solo::DevBuf scratch[32];
    // results of the last staged search
    solo::DevBuf r_best_row, r_best_score, r_n_pairs, r_pairs, r_n_cand, r_ovf;
    int r_nq = 0, r_max_pairs = 0;
    // query slots: the staged batch + its results above belong to the active slot; the others are parked
    struct Slot {
        int nq = 0;
        int64_t q_peaks = 0;
        int q_max_peaks = 0, q_mz_is_f64 = 0, r_nq = 0, r_max_pairs = 0;
        solo::DevBuf b[11];
    };
    int active_slot = 0;
    std::map<int, Slot> parked;
    // asynchronous staging / fetching (solo_stage_queries_async, solo_fetch_results_async): a copy stream of its
    // own and three events per slot order it against the compute stream
    struct SlotSync {
        cudaEvent_t staged = nullptr, done = nullptr, fetched = nullptr;
        bool wait_staged = false, has_done = false, has_fetched = false;
        int32_t *n_over = nullptr;   // pinned
    };
    cudaStream_t copy_stream = nullptr;
    std::map<int, SlotSync> slot_sync;
};

namespace solo {

// RAII stage timer: records CUDA events around a stage when profiling is enabled.
// NVTX ranges around the stages (SOLO_NVTX=1 or solo_set_option("nvtx", 1)): nvtx3 is header-only, the ranges cost nothing
// unless a tool is attached
const char *stage_name(int st);
struct StageTimer {
    solo_handle *h;
    int st;
    cudaEvent_t a = nullptr, b = nullptr;
    bool range = false;
    StageTimer(solo_handle *h_, int st_, int64_t launches, double units = 0.0, bool enabled = true) : h(h_), st(st_) {
        if (!enabled) return;
        if (h->opt_nvtx) {
            nvtxRangePushA(stage_name(st));
            range = true;
        }
        h->launches += launches;
        h->prof[st].launches += launches;
        h->prof[st].units += units;
        if (h->profile) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, h->stream);
        }
    }
    ~StageTimer() {
        if (a) {
            cudaEventRecord(b, h->stream);
            h->prof[st].pending.emplace_back(a, b);
        }
        if (range) nvtxRangePop();
    }
};

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// kernels' host launchers (defined in the .cu files)
void launch_vectorize(solo_handle *h, const void *d_mz, int mz_is_f64, const float *d_int, const int64_t *d_off,
                      int64_t n, int64_t n_peaks, int norm, float *d_out, __half *d_out_h, int scale_log2);

struct ScoreArgs {
    const float *q_mz;
    const float *q_int;
    const int64_t *q_off;
    const double *q_prec_mz;
    int nq;
    int q_max_peaks;
    const LibraryStore *lib;
    const int32_t *cand_ids;
    const int64_t *cand_off;   // CSR, or null with (cand_stride, cand_cnt)
    const int32_t *cand_cnt = nullptr;
    int cand_stride = 0;
    int tie_by_row = 0;
    double tol;
    int allow_shift;
    int max_pairs;
    int32_t *best_pos;   // position in candidate list
    int32_t *best_row;   // library row (may be null)
    double *best_score;
    int32_t *n_pairs;
    uint32_t *pairs;
    int32_t *overflow;   // device counter of match-list overflows
};
void launch_best_match(solo_handle *h, const ScoreArgs &a);

// K6 (k6_ssm_features.cu): SSM feature table, all device pointers
struct FeatureArgs {
    const float *q_mz32;
    const double *q_mz64;
    const float *q_int;
    const int64_t *q_off;
    const double *q_prec_mz;
    const int32_t *q_charge;      // per SSM, or null with q_charge_all
    int q_charge_all;
    const float *l_mz;
    const float *l_int;
    const int64_t *l_off;
    const double *l_prec_mz;
    const int32_t *lib_row;       // library store row matched to query i (< 0: no SSM)
    const uint32_t *pairs;        // (n, max_pairs, 2)
    const int32_t *n_pairs;
    int max_pairs;
    const int32_t *sequence_len;  // may be null
    int64_t n_peak_bins;
    const double *lfact, *lbig;   // log-gamma tables (k6::fill_log_tables)
    int n;
    double *out;                  // (n, N_FEATURES)
    int32_t *bad;                 // counter of SSMs beyond the peak capacity
};
void launch_ssm_features(solo_handle *h, const FeatureArgs &a);

// K0 (k0_process.cu): batched process_spectrum
struct ProcessArgs {
    const void *mz;          // float32 or float64, ascending inside every spectrum
    const float *inten;
    const int64_t *off;
    const double *prec_mz;
    const int32_t *prec_charge;
    int n;
    solo_process_params p;
    // outputs, row stride = p.max_peaks
    void *out_mz;            // same type as mz
    float *out_int;
    int32_t *out_idx;        // index of the kept peak inside its raw spectrum
    int32_t *out_cnt;        // peaks kept (0 for invalid spectra)
    uint8_t *out_valid;
    int32_t *err;            // [0]: spectra with more than K0_MAX_RAW peaks
};
void launch_process(solo_handle *h, const ProcessArgs &a, int mz_is_f64);
// mgf_io.cu: MGF query files -> CSR arrays (host only)
void mgf_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_title_bytes, int64_t *n_seq_bytes);
void mgf_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t n_title_bytes, int64_t n_seq_bytes,
              double *prec_mz, int32_t *prec_charge, double *rt, uint8_t *is_decoy, int64_t *peak_off, double *mz,
              float *inten, int64_t *title_off, char *titles, int64_t *seq_off, char *seqs);
// mzml_io.cu: mzML query files -> CSR arrays (host only)
void mzml_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_skipped);
void mzml_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t *scan_nr, int32_t *index, double *prec_mz,
               int32_t *prec_charge, double *rt, int64_t *peak_off, double *mz, float *inten);
void mzxml_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_skipped);
void mzxml_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t *scan_nr, int32_t *index, double *prec_mz,
                int32_t *prec_charge, double *rt, int64_t *peak_off, double *mz, float *inten);
// splib_io.cu: SpectraST .splib -> CSR arrays (host only)
void splib_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_peptide_bytes);
void splib_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t n_peptide_bytes, uint32_t *id,
                double *prec_mz, int32_t *prec_charge, uint8_t *is_decoy, int64_t *file_offset, int64_t *peak_off,
                float *mz, float *inten, uint8_t *peak_charge, int64_t *pep_off, char *pep);
void k5_build_aux(solo_handle *h, LibraryStore &L);

}  // namespace solo

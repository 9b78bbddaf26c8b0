// K5 — (shifted) dot product with greedy intensity-ordered peak assignment, best candidate
// per query.
//
// Replaces reference src/ann_solo/SpectrumMatch.cpp:8-133 (SpectrumMatcher::dot) and its
// Cython marshalling spectrum_match.pyx:28-108. One CTA per query, one warp per
// (query, candidate) pair:
//   * the query's peaks live in registers (lane = query peak), the candidate's m/z-sorted
//     peaks are staged in shared memory;
//   * per (query peak, shift) the reference's monotone two-pointer advance
//     (SpectrumMatch.cpp:39-46) is evaluated as a binary search for the first candidate peak
//     that does NOT satisfy `q_mz - tol > c_mz + mass_diff` (clamped to n-1 exactly like the
//     `< n - 1` guard), then the run of in-tolerance peaks is emitted (:49-85);
//   * tentative matches are packed into 64-bit keys (product | query peak | candidate peak),
//     rank-sorted in shared memory under the total order (product desc, query peak asc,
//     candidate peak asc) and consumed greedily (:95-111), the score summed in double in that
//     order;
//   * mixed precision is reproduced exactly: peaks float32, tolerance / mass shifts / compares
//     in double, product = (float)((mult * q_int) * c_int) evaluated in double, score in double.
// Per-warp best (first maximum wins, :118) is reduced across the CTA's warps.
#include "solo_common.cuh"

namespace solo {

constexpr int K5_WARPS = 8;
constexpr int K5_MAXM = 256;   // tentative matches per pair held on chip
constexpr int K5_MAXSHIFT = 8; // precursor charge <= 7

struct K5Params {
    const float *q_mz;
    const float *q_int;
    const int64_t *q_off;
    const double *q_prec_mz;
    const float *lib_mz;
    const float *lib_int;
    const uint8_t *lib_chg;
    const int64_t *lib_off;
    const double *lib_prec_mz;
    const int32_t *lib_prec_z;
    const int32_t *cand_ids;
    const int64_t *cand_off;   // CSR offsets, or null: strided lists (cand_stride, cand_cnt)
    const int32_t *cand_cnt;
    int cand_stride;
    int tie_by_row;            // ties -> lowest library row instead of lowest list position
    double tol;
    int allow_shift;
    int max_pairs;
    int32_t *best_pos;
    int32_t *best_row;
    double *best_score;
    int32_t *n_pairs;
    uint32_t *pairs;
    int32_t *overflow;
};

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

template <int PPL>
struct K5WarpMem {
    static constexpr int MAXP = 32 * PPL;
    double c_mz[MAXP];
    unsigned long long keys[K5_MAXM];
    float c_int[MAXP];
    uint16_t cur_pairs[MAXP];
    uint16_t best_pairs[MAXP];
    uint8_t c_chg[MAXP];
    int count;
    int pad;
};

template <int PPL>
__global__ void __launch_bounds__(K5_WARPS * 32) k5_best_match_kernel(K5Params p) {
    constexpr int MAXP = 32 * PPL;
    constexpr int NW = (MAXP + 63) / 64;  // 64-bit words in a peak-used mask
    constexpr int MAXR = K5_MAXM / 32;
    extern __shared__ __align__(16) unsigned char k5_smem[];
    K5WarpMem<PPL> *wm_all = reinterpret_cast<K5WarpMem<PPL> *>(k5_smem);
    __shared__ double s_best_score[K5_WARPS];
    __shared__ int s_best_pos[K5_WARPS];
    __shared__ int s_best_np[K5_WARPS];
    __shared__ int s_best_row[K5_WARPS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    K5WarpMem<PPL> &wm = wm_all[warp];
    const int q = blockIdx.x;
    const int64_t qb = p.q_off[q];
    const int nqp = (int)(p.q_off[q + 1] - qb);
    const double q_prec = p.q_prec_mz[q];
    const double tol = p.tol;

    // query peaks in registers: lane holds peaks lane, lane+32, ...
    double qm[PPL];
    float qi[PPL];
#pragma unroll
    for (int r = 0; r < PPL; ++r) {
        int i = lane + 32 * r;
        qm[r] = i < nqp ? (double)p.q_mz[qb + i] : 0.0;
        qi[r] = i < nqp ? p.q_int[qb + i] : 0.f;
    }

    const int64_t cb = p.cand_off ? p.cand_off[q] : (int64_t)q * p.cand_stride;
    const int64_t ce = p.cand_off ? p.cand_off[q + 1] : cb + p.cand_cnt[q];
    double best_score = 0.0;
    int best_pos = -1, best_np = 0, best_rowid = 0x7fffffff;

    for (int64_t c = cb + warp; c < ce; c += K5_WARPS) {
        const int row = p.cand_ids[c];
        const int64_t b = p.lib_off[row];
        const int n = (int)(p.lib_off[row + 1] - b);
#pragma unroll
        for (int r = 0; r < PPL; ++r) {
            int j = lane + 32 * r;
            if (j < n) {
                wm.c_mz[j] = (double)p.lib_mz[b + j];
                wm.c_int[j] = p.lib_int[b + j];
                wm.c_chg[j] = p.lib_chg[b + j];
            }
        }
        if (lane == 0) wm.count = 0;
        __syncwarp();

        // SpectrumMatch.cpp:18-31
        const int z = p.lib_prec_z[row];
        const double delta = __dmul_rn(__dsub_rn(q_prec, p.lib_prec_mz[row]), (double)z);
        const int nshift = (p.allow_shift && fabs(delta) >= tol) ? z + 1 : 1;
        const double md_lane = (lane > 0 && lane < nshift) ? __ddiv_rn(delta, (double)lane) : 0.0;

        if (n > 0) {
            for (int s = 0; s < nshift; ++s) {
                const double md = __shfl_sync(0xffffffffu, md_lane, s);
#pragma unroll
                for (int r = 0; r < PPL; ++r) {
                    const int i = lane + 32 * r;
                    if (i < nqp) {
                        const double thr = __dsub_rn(qm[r], tol);
                        int lo = 0, hi = n - 1;
                        while (lo < hi) {
                            int mid = (lo + hi) >> 1;
                            if (thr > __dadd_rn(wm.c_mz[mid], md)) lo = mid + 1;
                            else hi = mid;
                        }
                        for (int j = lo; j < n; ++j) {
                            double d = fabs(__dsub_rn(qm[r], __dadd_rn(wm.c_mz[j], md)));
                            if (!(d <= tol)) break;
                            const int cz = wm.c_chg[j];
                            double mult = 0.0;
                            if (s == 0) mult = 1.0;
                            else if (cz == s) mult = 1.0;
                            else if (cz == 0) mult = 2.0 / 3.0;
                            if (mult > 0.0) {
                                float prod = __double2float_rn(
                                    __dmul_rn(__dmul_rn(mult, (double)qi[r]), (double)wm.c_int[j]));
                                int slot = atomicAdd(&wm.count, 1);
                                if (slot < K5_MAXM) {
                                    wm.keys[slot] = ((unsigned long long)float_to_ordered(prod) << 32) |
                                                    ((unsigned long long)(0xFFFFu - (unsigned)i) << 16) |
                                                    (unsigned long long)(0xFFFFu - (unsigned)j);
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        int M = wm.count;
        if (M > K5_MAXM) {  // reported, never silently truncated
            if (lane == 0) atomicAdd(p.overflow, 1);
            M = K5_MAXM;
        }

        // rank sort, descending, in place (keys are unique up to exact duplicates, which are
        // ordered by slot and are interchangeable for the greedy pass)
        if (M > 1) {
            unsigned long long mine[MAXR];
            int rank[MAXR];
            const int R = (M + 31) >> 5;
#pragma unroll
            for (int r = 0; r < MAXR; ++r) {
                int idx = lane + 32 * r;
                mine[r] = (r < R && idx < M) ? wm.keys[idx] : 0ull;
                rank[r] = 0;
            }
            for (int j = 0; j < M; ++j) {
                const unsigned long long kj = wm.keys[j];
#pragma unroll
                for (int r = 0; r < MAXR; ++r) {
                    if (r < R) {
                        int idx = lane + 32 * r;
                        rank[r] += (kj > mine[r]) || (kj == mine[r] && j < idx);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < MAXR; ++r) {
                int idx = lane + 32 * r;
                if (r < R && idx < M) wm.keys[rank[r]] = mine[r];
            }
            __syncwarp();
        }

        // greedy assignment (SpectrumMatch.cpp:95-111), warp-uniform
        double score = 0.0;
        int np = 0;
        unsigned long long qu[NW], cu[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) qu[w] = cu[w] = 0ull;
        const int max_np = min(nqp, n);
        for (int m = 0; m < M && np < max_np; ++m) {
            const unsigned long long key = wm.keys[m];
            const int i = 0xFFFF - (int)((key >> 16) & 0xFFFFu);
            const int j = 0xFFFF - (int)(key & 0xFFFFu);
            bool used = false;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                if ((i >> 6) == w) used |= (qu[w] >> (i & 63)) & 1ull;
                if ((j >> 6) == w) used |= (cu[w] >> (j & 63)) & 1ull;
            }
            if (!used) {
                score = __dadd_rn(score, (double)ordered_to_float((uint32_t)(key >> 32)));
                if (lane == 0) wm.cur_pairs[np] = (uint16_t)((i << 8) | j);
                ++np;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    if ((i >> 6) == w) qu[w] |= 1ull << (i & 63);
                    if ((j >> 6) == w) cu[w] |= 1ull << (j & 63);
                }
            }
        }

        // SpectrumMatch.cpp:118 — first candidate, then strictly greater only
        if (best_pos < 0 || best_score < score || (p.tie_by_row && best_score == score && row < best_rowid)) {
            best_score = score;
            best_pos = (int)(c - cb);
            best_rowid = row;
            best_np = np;
            __syncwarp();
            for (int t = lane; t < np; t += 32) wm.best_pairs[t] = wm.cur_pairs[t];
        }
        __syncwarp();
    }

    if (lane == 0) {
        s_best_score[warp] = best_score;
        s_best_pos[warp] = best_pos;
        s_best_np[warp] = best_np;
        s_best_row[warp] = best_rowid;
    }
    __syncthreads();
    // winner: maximum score, ties -> earliest candidate position (== the sequential rule)
    int win = -1;
    double ws = 0.0;
    int wp = -1, wr = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < K5_WARPS; ++w) {
        int pos = s_best_pos[w];
        double sc = s_best_score[w];
        int rw = s_best_row[w];
        bool earlier = p.tie_by_row ? (rw < wr) : (pos < wp);
        if (pos >= 0 && (win < 0 || sc > ws || (sc == ws && earlier))) {
            win = w;
            ws = sc;
            wp = pos;
            wr = rw;
        }
    }
    if (win < 0) {
        if (threadIdx.x == 0) {
            p.best_pos[q] = -1;
            if (p.best_row) p.best_row[q] = -1;
            p.best_score[q] = 0.0;
            p.n_pairs[q] = 0;
        }
        return;
    }
    if (warp == win) {
        const int np = s_best_np[win];
        if (lane == 0) {
            p.best_pos[q] = wp;
            if (p.best_row) p.best_row[q] = p.cand_ids[cb + wp];
            p.best_score[q] = ws;
            p.n_pairs[q] = np;
        }
        uint32_t *out = p.pairs + (size_t)q * p.max_pairs * 2;
        for (int t = lane; t < np && t < p.max_pairs; t += 32) {
            uint16_t pr = wm.best_pairs[t];
            out[2 * t] = pr >> 8;
            out[2 * t + 1] = pr & 0xFFu;
        }
    }
}

template <int PPL>
static void launch_k5(solo_handle *h, const K5Params &p, int nq) {
    auto k = k5_best_match_kernel<PPL>;
    size_t smem = sizeof(K5WarpMem<PPL>) * K5_WARPS;
    SOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<nq, K5_WARPS * 32, smem, h->stream>>>(p);
    SOLO_CUDA(cudaGetLastError());
}

void launch_best_match(solo_handle *h, const ScoreArgs &a) {
    if (a.nq <= 0) return;
    const LibraryStore &L = *a.lib;
    int maxp = a.q_max_peaks > L.max_peaks ? a.q_max_peaks : L.max_peaks;
    SOLO_REQUIRE(maxp <= 128, SOLO_ECAPACITY,
                 "spectra with more than 128 peaks are not supported by the scorer (got %d)", maxp);
    K5Params p;
    p.q_mz = a.q_mz;
    p.q_int = a.q_int;
    p.q_off = a.q_off;
    p.q_prec_mz = a.q_prec_mz;
    p.lib_mz = L.mz.as<float>();
    p.lib_int = L.inten.as<float>();
    p.lib_chg = L.chg.as<uint8_t>();
    p.lib_off = L.off.as<int64_t>();
    p.lib_prec_mz = L.prec_mz.as<double>();
    p.lib_prec_z = L.prec_z.as<int32_t>();
    p.cand_ids = a.cand_ids;
    p.cand_off = a.cand_off;
    p.cand_cnt = a.cand_cnt;
    p.cand_stride = a.cand_stride;
    p.tie_by_row = a.tie_by_row;
    p.tol = a.tol;
    p.allow_shift = a.allow_shift;
    p.max_pairs = a.max_pairs;
    p.best_pos = a.best_pos;
    p.best_row = a.best_row;
    p.best_score = a.best_score;
    p.n_pairs = a.n_pairs;
    p.pairs = a.pairs;
    p.overflow = a.overflow;
    StageTimer t(h, ST_SCORE, 1);
    if (maxp <= 64) launch_k5<2>(h, p, a.nq);
    else launch_k5<4>(h, p, a.nq);
}

}  // namespace solo

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
constexpr int K5_NBUCKET = 256;   // bucket table entries per library spectrum (fast path)
constexpr int K5_BUCKET_MZ = 8;    // m/z width of one bucket

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


// ---------------------------------------------------------------------------------------------
// Overflow path: a (query, candidate) pair with more than K5_MAXM tentative peak matches (wide
// fragment tolerances, dense spectra). The greedy assignment over the matches sorted by
// (product desc, query peak asc, candidate peak asc) is the same as repeatedly taking the largest
// remaining match whose two peaks are still free — so the matches are re-enumerated once per
// accepted pair instead of being stored: no capacity, no scratch memory, exact. Warp-cooperative;
// the candidate's peaks are in `c_mz / c_int / c_chg` (shared memory), the query's peaks are read
// through `qmz(i) / qint(i)`, the mass shift of shift s is md_of_shift[s] (shared memory). Returns the score; pairs go to `cur_pairs`, *np_out their count.
template <typename QMz, typename QInt>
__device__ double k5_greedy_by_reselection(const double *c_mz, const float *c_int, const uint8_t *c_chg, int n, int nqp,
                                           int nshift, const double *md_of_shift, double tol, QMz qmz, QInt qint,
                                           uint16_t *cur_pairs, int *np_out) {
    const int lane = threadIdx.x & 31;
    unsigned long long qu[2] = {0ull, 0ull}, cu[2] = {0ull, 0ull};  // used peaks (<= 128 each)
    double score = 0.0;
    int np = 0;
    const int max_np = min(nqp, n);
    const int W = nshift * nqp;
    while (np < max_np) {
        unsigned long long best = 0ull;
        for (int w = lane; w < W; w += 32) {
            const int s = w / nqp, i = w - s * nqp;
            if ((qu[i >> 6] >> (i & 63)) & 1ull) continue;
            const double md = md_of_shift[s];
            const double qm = qmz(i);
            const double thr = __dsub_rn(qm, tol);
            int lo = 0, hi = n - 1;  // SpectrumMatch.cpp:39-46 as a binary search (clamped to n - 1)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (thr > __dadd_rn(c_mz[mid], md)) lo = mid + 1;
                else hi = mid;
            }
            for (int j = lo; j < n; ++j) {
                const double d = fabs(__dsub_rn(qm, __dadd_rn(c_mz[j], md)));
                if (!(d <= tol)) break;
                if ((cu[j >> 6] >> (j & 63)) & 1ull) continue;
                const int cz = c_chg[j];
                double mult = 0.0;
                if (s == 0) mult = 1.0;
                else if (cz == s) mult = 1.0;
                else if (cz == 0) mult = 2.0 / 3.0;
                if (mult > 0.0) {
                    const float prod = __double2float_rn(__dmul_rn(__dmul_rn(mult, (double)qint(i)), (double)c_int[j]));
                    const unsigned long long key = ((unsigned long long)float_to_ordered(prod) << 32) |
                                                   ((unsigned long long)(0xFFFFu - (unsigned)i) << 16) |
                                                   (unsigned long long)(0xFFFFu - (unsigned)j);
                    best = key > best ? key : best;
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
            best = y > best ? y : best;
        }
        if (best == 0ull) break;
        const int i = 0xFFFF - (int)((best >> 16) & 0xFFFFu);
        const int j = 0xFFFF - (int)(best & 0xFFFFu);
        score = __dadd_rn(score, (double)ordered_to_float((uint32_t)(best >> 32)));
        if (lane == 0) cur_pairs[np] = (uint16_t)((i << 8) | j);
        ++np;
        qu[i >> 6] |= 1ull << (i & 63);
        cu[j >> 6] |= 1ull << (j & 63);
    }
    __syncwarp();
    *np_out = np;
    return score;
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
    double md[8];  // mass shift per shift index (overflow path)
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
        // delta / lane: lanes 1, 2 and 4 divide by a power of two (exact scaling, |delta| >= tol here, so no
        // subnormals); only a lane 3, 5, 6 or 7 (precursor charge >= 3) runs the ~40-instruction double division
        double md_lane = 0.0;
        if (lane > 0 && lane < nshift) {
            if (lane == 1) md_lane = delta;
            else if (lane == 2) md_lane = __dmul_rn(delta, 0.5);
            else if (lane == 4) md_lane = __dmul_rn(delta, 0.25);
            else md_lane = __ddiv_rn(delta, (double)lane);
        }

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
        const int M = wm.count;
        double score = 0.0;
        int np = 0;
        if (M > K5_MAXM) {
            // more tentative matches than the on-chip list holds: exact greedy by re-enumeration (counted, never truncated)
            if (lane == 0) atomicAdd(p.overflow, 1);
            if (lane < 8) wm.md[lane] = md_lane;
            __syncwarp();
            score = k5_greedy_by_reselection(
                wm.c_mz, wm.c_int, wm.c_chg, n, nqp, nshift, wm.md, tol, [&](int i) { return (double)p.q_mz[qb + i]; },
                [&](int i) { return p.q_int[qb + i]; }, wm.cur_pairs, &np);
        } else {
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

// ======================================================================= fast path (<= 64 peaks)
//
// Same arithmetic and ordering as k5_best_match_kernel, restructured for throughput:
//  * the next candidate's metadata (one 32-byte record per library spectrum), peaks and bucket
//    table are fetched into registers while the current candidate is scored (two-deep software
//    pipeline), candidate ids are fetched 32 at a time;
//  * the start of the reference's pointer advance comes from a per-spectrum bucket table
//    (K5_NBUCKET buckets of K5_BUCKET_MZ m/z: number of peaks below the bucket's lower edge, built
//    once at load time) followed by the reference's own advance loop, instead of a binary search;
//  * (query peak, shift) combinations are flattened over the lanes, so a 35-peak query with three
//    shifts takes 4 passes instead of 6.

struct __align__(16) LibMeta {
    int64_t off;
    int32_t n;
    int32_t z;
    double prec_mz;
    double pad;
};

struct K5FastWarpMem {
    double c_mz[64];
    unsigned long long keys[K5_MAXM];
    float c_int[64];
    uint16_t cur_pairs[64];
    uint16_t best_pairs[64];
    uint8_t c_chg[64];
    __align__(8) uint8_t table[K5_NBUCKET];
    int count;
    int pad;
    double md[8];  // mass shift per shift index
    float mdf[8];  // ... as float (bucket lookup only)
};

struct K5FastCtaMem {
    K5FastWarpMem w[K5_WARPS];
    double qmz[64];
    double qthr[64];
    double best_score[K5_WARPS];
    float qtf[64];
    float qint[64];
    int best_pos[K5_WARPS];
    int best_np[K5_WARPS];
    int best_row[K5_WARPS];
};

__global__ void k5_build_meta_kernel(const int64_t *__restrict__ off, const double *__restrict__ prec_mz,
                                     const int32_t *__restrict__ prec_z, int64_t n, LibMeta *__restrict__ meta) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    LibMeta m;
    m.off = off[r];
    m.n = (int32_t)(off[r + 1] - off[r]);
    m.z = prec_z[r];
    m.prec_mz = prec_mz[r];
    m.pad = 0.0;
    meta[r] = m;
}

// table[r][b] = number of peaks of row r with m/z < b * K5_BUCKET_MZ (saturated at 255)
__global__ void k5_build_table_kernel(const float *__restrict__ mz, const int64_t *__restrict__ off, int64_t n,
                                      uint8_t *__restrict__ table) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * K5_NBUCKET) return;
    const int64_t r = t / K5_NBUCKET;
    const int b = (int)(t % K5_NBUCKET);
    const float edge = (float)(b * K5_BUCKET_MZ);
    const float *m = mz + off[r];
    int lo = 0, hi = (int)(off[r + 1] - off[r]);
    while (lo < hi) {  // first peak with mz >= edge
        const int mid = (lo + hi) >> 1;
        if (m[mid] < edge) lo = mid + 1;
        else hi = mid;
    }
    table[t] = (uint8_t)min(lo, 255);
}

void k5_build_aux(solo_handle *h, LibraryStore &L) {
    L.meta.ensure(std::max<int64_t>(L.n, 1) * sizeof(LibMeta));
    L.table.ensure(std::max<int64_t>(L.n, 1) * K5_NBUCKET);
    if (L.n == 0) return;
    k5_build_meta_kernel<<<div_up(L.n, 256), 256, 0, h->stream>>>(L.off.as<int64_t>(), L.prec_mz.as<double>(),
                                                                  L.prec_z.as<int32_t>(), L.n, L.meta.as<LibMeta>());
    k5_build_table_kernel<<<div_up(L.n * K5_NBUCKET, 256), 256, 0, h->stream>>>(L.mz.as<float>(), L.off.as<int64_t>(),
                                                                                L.n, L.table.as<uint8_t>());
    SOLO_CUDA(cudaGetLastError());
    h->launches += 2;
}

struct K5FastParams {
    K5Params p;
    const LibMeta *meta;
    const uint8_t *table;
    int debug;  // timing experiments only (SOLO_K5_DEBUG): 1 = skip the match passes, 2 = skip sort + greedy
};

__global__ void __launch_bounds__(K5_WARPS * 32, 3) k5_fast_kernel(K5FastParams fp) {
    const K5Params &p = fp.p;
    // one block of dynamic shared memory for everything (per-warp areas, the query, the per-warp winners): every
    // access is base register + immediate (separate static arrays cost an address computation per use)
    extern __shared__ __align__(16) unsigned char k5_smem[];
    K5FastCtaMem &cm = *reinterpret_cast<K5FastCtaMem *>(k5_smem);
    K5FastWarpMem *wm_all = cm.w;
    double (&s_qmz)[64] = cm.qmz;
    double (&s_qthr)[64] = cm.qthr;  // q_mz - tol (SpectrumMatch.cpp:41)
    float (&s_qtf)[64] = cm.qtf;     // float copy of it minus the bucket safety margin
    float (&s_qint)[64] = cm.qint;
    double (&s_best_score)[K5_WARPS] = cm.best_score;
    int (&s_best_pos)[K5_WARPS] = cm.best_pos;
    int (&s_best_np)[K5_WARPS] = cm.best_np;
    int (&s_best_row)[K5_WARPS] = cm.best_row;
    constexpr int MAXR = K5_MAXM / 32;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    K5FastWarpMem &wm = wm_all[warp];
    const int q = blockIdx.x;
    const int64_t qb = p.q_off[q];
    const int nqp = (int)(p.q_off[q + 1] - qb);
    const double q_prec = p.q_prec_mz[q];
    const double tol = p.tol;
    for (int i = threadIdx.x; i < nqp; i += blockDim.x) {
        const double qm = (double)p.q_mz[qb + i];
        s_qmz[i] = qm;
        s_qthr[i] = __dsub_rn(qm, tol);
        s_qtf[i] = __double2float_rn(__dsub_rn(qm, tol)) - 0.004f;
        s_qint[i] = p.q_int[qb + i];
    }
    __syncthreads();

    const int64_t cb = p.cand_off ? p.cand_off[q] : (int64_t)q * p.cand_stride;
    const int64_t ce = p.cand_off ? p.cand_off[q + 1] : cb + p.cand_cnt[q];
    const int64_t ncand = ce - cb;
    // this warp scores candidates warp, warp + 8, ...: T of them
    const int64_t T = ncand > warp ? (ncand - warp + K5_WARPS - 1) / K5_WARPS : 0;
    double best_score = 0.0;
    int best_pos = -1, best_np = 0, best_rowid = 0x7fffffff;

    // candidate ids, 32 iterations per register; the block after the current one is in flight
    auto load_ids = [&](int64_t t0) -> int {
        const int64_t t = t0 + lane;
        return t < T ? p.cand_ids[cb + warp + t * K5_WARPS] : 0;
    };
    int ids_cur = load_ids(0), ids_nxt = load_ids(32);

    // pipeline registers
    LibMeta m_cur, m_nxt;   // metadata of candidates t and t + 1
    int row_cur = 0, row_nxt = 0;
    float pk_mz[2], pk_int[2];
    uint32_t pk_chg[2];
    uint2 pk_tab = make_uint2(0u, 0u);
    static_assert(K5_NBUCKET == 256, "the bucket table is fetched as one uint2 per lane");
    m_cur.off = 0; m_cur.n = 0; m_cur.z = 0; m_cur.prec_mz = 0.0; m_cur.pad = 0.0;
    m_nxt = m_cur;

    auto fetch_meta = [&](int row, LibMeta &m) {
        const uint4 a = *reinterpret_cast<const uint4 *>(&fp.meta[row]);
        m.off = (int64_t)(((unsigned long long)a.y << 32) | a.x);
        m.n = (int32_t)a.z;
        m.z = (int32_t)a.w;
        m.prec_mz = fp.meta[row].prec_mz;
    };
    auto fetch_peaks = [&](int row, const LibMeta &m) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = lane + 32 * r;
            const bool ok = j < m.n;
            pk_mz[r] = ok ? p.lib_mz[m.off + j] : 0.f;
            pk_int[r] = ok ? p.lib_int[m.off + j] : 0.f;
            pk_chg[r] = ok ? (uint32_t)p.lib_chg[m.off + j] : 0u;
        }
        pk_tab = reinterpret_cast<const uint2 *>(fp.table + (size_t)row * K5_NBUCKET)[lane];
    };

    if (T > 0) {
        row_cur = __shfl_sync(0xffffffffu, ids_cur, 0);
        fetch_meta(row_cur, m_cur);
        fetch_peaks(row_cur, m_cur);
        if (T > 1) {
            row_nxt = __shfl_sync(0xffffffffu, ids_cur, 1);
            fetch_meta(row_nxt, m_nxt);
        }
    }

    for (int64_t t = 0; t < T; ++t) {
        const int row = row_cur;
        const int n = m_cur.n;
        const int z = m_cur.z;
        const double c_prec = m_cur.prec_mz;
        // ---- registers -> shared memory for candidate t
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = lane + 32 * r;
            if (j < n) {
                wm.c_mz[j] = (double)pk_mz[r];
                wm.c_int[j] = pk_int[r];
                wm.c_chg[j] = (uint8_t)pk_chg[r];
            }
        }
        reinterpret_cast<uint2 *>(wm.table)[lane] = pk_tab;
        if (lane == 0) wm.count = 0;
        __syncwarp();
        // ---- prefetch: peaks of t + 1 (its metadata arrived during the previous iteration), metadata of t + 2
        if (t + 1 < T) {
            m_cur = m_nxt;
            row_cur = row_nxt;
            fetch_peaks(row_cur, m_cur);
            if (t + 2 < T) {
                const int64_t t2 = t + 2;
                if ((t2 & 31) == 0) {  // entering a new id block: rotate
                    ids_cur = ids_nxt;
                    ids_nxt = load_ids(t2 + 32);
                }
                // ids_cur covers the block containing t2 once rotated; before rotation t2 is in the same block as t
                row_nxt = __shfl_sync(0xffffffffu, ids_cur, (int)(t2 & 31));
                fetch_meta(row_nxt, m_nxt);
            }
        }

        // SpectrumMatch.cpp:18-31
        const double delta = __dmul_rn(__dsub_rn(q_prec, c_prec), (double)z);
        const int nshift = (p.allow_shift && fabs(delta) >= tol) ? z + 1 : 1;
        // delta / lane: lanes 1, 2 and 4 divide by a power of two (exact scaling, |delta| >= tol here, so no
        // subnormals); only a lane 3, 5, 6 or 7 (precursor charge >= 3) runs the ~40-instruction double division
        double md_lane = 0.0;
        if (lane > 0 && lane < nshift) {
            if (lane == 1) md_lane = delta;
            else if (lane == 2) md_lane = __dmul_rn(delta, 0.5);
            else if (lane == 4) md_lane = __dmul_rn(delta, 0.25);
            else md_lane = __ddiv_rn(delta, (double)lane);
        }
        if (lane < 8) {
            wm.md[lane] = md_lane;
            wm.mdf[lane] = __double2float_rn(md_lane);
        }
        __syncwarp();

        if (n > 0 && !(fp.debug & 1)) {
            const int W = nshift * nqp;
            const float inv_nqp = 1.0f / (float)nqp;
            for (int w0 = 0; w0 < W; w0 += 32) {
                const int w = w0 + lane;
                const bool active = w < W;
                // w = s * nqp + i; exact for these small integers
                const int s = active ? __float2int_rz(((float)w + 0.5f) * inv_nqp) : 0;
                const int i = active ? w - s * nqp : 0;
                if (active) {
                    const double md = wm.md[s];
                    const double qm = s_qmz[i];
                    const double thr = s_qthr[i];
                    // bucket start: every peak below the bucket edge satisfies thr > c_mz + md with a wide margin
                    // (float rounding at m/z 2000 is 1.2e-4 per term, far inside the 0.004 margin)
                    int b = __float2int_rd((s_qtf[i] - wm.mdf[s]) * (1.0f / K5_BUCKET_MZ));
                    b = max(0, min(K5_NBUCKET - 1, b));
                    int j = min((int)wm.table[b], n - 1);
                    double x = __dadd_rn(wm.c_mz[j], md);  // candidate peak m/z + mass shift, reused by the match test
                    while (j < n - 1 && thr > x) {          // SpectrumMatch.cpp:39-46 (at most a few steps from the bucket start)
                        ++j;
                        x = __dadd_rn(wm.c_mz[j], md);
                    }
                    for (;;) {
                        const double d = fabs(__dsub_rn(qm, x));
                        if (!(d <= tol)) break;
                        const int cz = wm.c_chg[j];
                        double mult = 0.0;
                        if (s == 0) mult = 1.0;
                        else if (cz == s) mult = 1.0;
                        else if (cz == 0) mult = 2.0 / 3.0;
                        if (mult > 0.0) {
                            const float prod = __double2float_rn(
                                __dmul_rn(__dmul_rn(mult, (double)s_qint[i]), (double)wm.c_int[j]));
                            const int slot = atomicAdd(&wm.count, 1);
                            if (slot < K5_MAXM) {
                                wm.keys[slot] = ((unsigned long long)float_to_ordered(prod) << 32) |
                                                ((unsigned long long)(0xFFFFu - (unsigned)i) << 16) |
                                                (unsigned long long)(0xFFFFu - (unsigned)j);
                            }
                        }
                        if (++j >= n) break;
                        x = __dadd_rn(wm.c_mz[j], md);
                    }
                }
            }
        }
        __syncwarp();
        const int M = (fp.debug & 2) ? 0 : wm.count;
        double score = 0.0;
        int np = 0;
        if (M > K5_MAXM) {
            // more tentative matches than the on-chip list holds: exact greedy by re-enumeration (counted, never truncated)
            if (lane == 0) atomicAdd(p.overflow, 1);
            score = k5_greedy_by_reselection(
                wm.c_mz, wm.c_int, wm.c_chg, n, nqp, nshift, wm.md, tol, [&](int i) { return s_qmz[i]; },
                [&](int i) { return s_qint[i]; }, wm.cur_pairs, &np);
        } else {
            // rank sort, descending, in place
            if (M > 1) {
                unsigned long long mine[MAXR];
                int rank[MAXR];
                const int R = (M + 31) >> 5;
    #pragma unroll
                for (int r = 0; r < MAXR; ++r) {
                    const int idx = lane + 32 * r;
                    mine[r] = (r < R && idx < M) ? wm.keys[idx] : 0ull;
                    rank[r] = 0;
                }
                for (int j = 0; j < M; ++j) {
                    const unsigned long long kj = wm.keys[j];
    #pragma unroll
                    for (int r = 0; r < MAXR; ++r) {
                        if (r < R) {
                            const int idx = lane + 32 * r;
                            rank[r] += (kj > mine[r]) || (kj == mine[r] && j < idx);
                        }
                    }
                }
                __syncwarp();
    #pragma unroll
                for (int r = 0; r < MAXR; ++r) {
                    const int idx = lane + 32 * r;
                    if (r < R && idx < M) wm.keys[rank[r]] = mine[r];
                }
                __syncwarp();
            }

            // greedy assignment (SpectrumMatch.cpp:95-111), warp-uniform
            unsigned long long qu = 0ull, cu = 0ull;
            const int max_np = min(nqp, n);
            for (int m = 0; m < M && np < max_np; ++m) {
                const unsigned long long key = wm.keys[m];
                const int i = 0xFFFF - (int)((key >> 16) & 0xFFFFu);
                const int j = 0xFFFF - (int)(key & 0xFFFFu);
                if (!(((qu >> i) | (cu >> j)) & 1ull)) {
                    score = __dadd_rn(score, (double)ordered_to_float((uint32_t)(key >> 32)));
                    if (lane == 0) wm.cur_pairs[np] = (uint16_t)((i << 8) | j);
                    ++np;
                    qu |= 1ull << i;
                    cu |= 1ull << j;
                }
            }

        }

        // SpectrumMatch.cpp:118 — first candidate, then strictly greater only
        const int pos = (int)(warp + t * K5_WARPS);
        if (best_pos < 0 || best_score < score || (p.tie_by_row && best_score == score && row < best_rowid)) {
            best_score = score;
            best_pos = pos;
            best_rowid = row;
            best_np = np;
            __syncwarp();
            for (int u = lane; u < np; u += 32) wm.best_pairs[u] = wm.cur_pairs[u];
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
        const int pos = s_best_pos[w];
        const double sc = s_best_score[w];
        const int rw = s_best_row[w];
        const bool earlier = p.tie_by_row ? (rw < wr) : (pos < wp);
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
            if (p.best_row) p.best_row[q] = wr;
            p.best_score[q] = ws;
            p.n_pairs[q] = np;
        }
        uint32_t *out = p.pairs + (size_t)q * p.max_pairs * 2;
        for (int u = lane; u < np && u < p.max_pairs; u += 32) {
            const uint16_t pr = wm.best_pairs[u];
            out[2 * u] = pr >> 8;
            out[2 * u + 1] = pr & 0xFFu;
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
    static const bool v_old = getenv("SOLO_K5_OLD") != nullptr;
    if (maxp <= 64 && L.meta.p && !v_old) {
        K5FastParams fp;
        fp.p = p;
        fp.meta = L.meta.as<LibMeta>();
        fp.table = L.table.as<uint8_t>();
        static const int v_dbg = getenv("SOLO_K5_DEBUG") ? atoi(getenv("SOLO_K5_DEBUG")) : 0;
        fp.debug = v_dbg;
        const size_t smem = sizeof(K5FastCtaMem);
        SOLO_CUDA(cudaFuncSetAttribute(k5_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k5_fast_kernel<<<a.nq, K5_WARPS * 32, smem, h->stream>>>(fp);
        SOLO_CUDA(cudaGetLastError());
    } else if (maxp <= 64) launch_k5<2>(h, p, a.nq);
    else launch_k5<4>(h, p, a.nq);
}

}  // namespace solo

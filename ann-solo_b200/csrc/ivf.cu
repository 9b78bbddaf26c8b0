// IVF-Flat inner-product index on the device (K2/K3/K4 of the hot path).
//
// Replaces the Faiss objects used at reference src/ann_solo/spectral_library.py:167-181
// (IndexFlatIP + IndexIVFFlat(METRIC_INNER_PRODUCT): train / add) and :443-444, :490-497
// (nprobe, search). Faiss is a third-party library that is absent from the reference tree, so the
// arithmetic is the one the oracle defines (oracle/solo_oracle.cpp §4):
//   score(q, x) = fp32 sequential fmaf over dimensions 0..d-1; order = (score desc, id asc).
//
// Pipeline for one batch of queries:
//   1. queries -> sparse rows (hashed spectra have <= ~50 non-zeros of 800)
//   2. K2 coarse: exact scores against all centroids (centroid tile transposed in shared
//      memory, lane = centroid, sequential fmaf over the query's non-zeros in index order)
//   3. nprobe selection: block radix-select on (score, id) keys
//   4. (query, probe) pairs inverted into per-list query groups (two rounds, see below)
//   5. K3 list scan: every list is read once per round and scored against its query group;
//      a score reaches HBM only if it passes the query's running threshold
//   6. K4: k-th largest by radix select; ties broken by id; optional fused precursor-window
//      mask (applied AFTER the top-k, reference spectral_library.py:441-446).
// Round 0 scans each query's first probes without a threshold (bounded by C0 scores) to seed
// tau[q] = k-th best so far; round 1 scans the rest and appends only scores >= tau[q].
#include "ivf.cuh"

#include <algorithm>
#include <cmath>
#include <functional>
#include <random>

namespace solo {

// ======================================================================= small utilities

__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int32_t *__restrict__ cnt, int64_t n, int64_t *__restrict__ off, int64_t base) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = base;
    __syncthreads();
    for (int64_t tile = 0; tile < n; tile += 1024) {
        int64_t i = tile + tid;
        int64_t v = i < n ? (int64_t)max(cnt[i], 0) : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        int64_t prefix = (warp > 0 ? s_warp[warp - 1] : 0) + s_carry;
        if (i < n) off[i] = prefix + x - v;
        __syncthreads();
        if (tid == 1023) s_carry = prefix + x;
        __syncthreads();
    }
    if (tid == 0) off[n] = s_carry;
}

static void scan_counts(solo_handle *h, const int32_t *cnt, int64_t n, int64_t *off, int64_t base) {
    scan_counts_kernel<<<1, 1024, 0, h->stream>>>(cnt, n, off, base);
    SOLO_CUDA(cudaGetLastError());
    h->launches++;
}

void scan_counts_i32(solo_handle *h, const int32_t *cnt, int64_t n, int64_t *off) { scan_counts(h, cnt, n, off, 0); }

constexpr int KTH_BINS = 2048;  // shared-memory histogram bins of block_kth_largest_u32
constexpr int KTH_BC = 48;      // broadcast / warp-sum scratch words

// k-th largest (k >= 1, k <= number of participating keys) of `n` 32-bit keys in shared memory,
// restricted to entries with key != skip_key when use_skip. Every thread of the block must call
// it (blockDim.x a multiple of 32, at most 1024). On return *count_gt holds the number of
// participating keys strictly greater than the result.
// Radix select, most significant digit first, 11 bits per pass, starting at the highest bit in
// which the block's minimum and maximum differ: scores of one query share their upper bits, so
// the first histogram is already well spread (few same-bin shared-memory atomics).
// When k2 > 0 (k2 <= k), *front_key receives a key F (resolution: the first radix pass) such that at
// least k2 participating keys are >= F — a cheap "roughly the k2 best" split.
__device__ uint32_t block_kth_largest_u32(const uint32_t *keys, int n, int k, bool use_skip, uint32_t skip_key,
                                          uint32_t *s_hist /*KTH_BINS*/, uint32_t *s_bc /*KTH_BC*/, int *count_gt,
                                          int k2 = 0, uint32_t *front_key = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        s_bc[4] = 0xFFFFFFFFu;
        s_bc[5] = 0u;
    }
    __syncthreads();
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t key = keys[i];
        if (use_skip && key == skip_key) continue;
        mn = min(mn, key);
        mx = max(mx, key);
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) {
        atomicMin(&s_bc[4], mn);
        atomicMax(&s_bc[5], mx);
    }
    __syncthreads();
    mn = s_bc[4];
    mx = s_bc[5];
    __syncthreads();
    if (mn >= mx) {  // all participating keys equal (or none participates)
        *count_gt = 0;
        if (front_key) *front_key = mx;
        return mx;
    }
    bool first_pass = true;
    int top = 32 - __clz(mn ^ mx);  // unresolved low bits
    uint32_t mask = top >= 32 ? 0u : ~((1u << top) - 1u);
    uint32_t prefix = mx & mask;
    int kk = k, gt = 0;
    while (top > 0) {
        const int w = min(11, top);
        const int shift = top - w;
        const int nb = 1 << w;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t key = keys[i];
            if (use_skip && key == skip_key) continue;
            if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & (uint32_t)(nb - 1)], 1u);
        }
        __syncthreads();
        // thread t owns bins [t * per, (t + 1) * per); suffix sums over threads, highest bins first
        const int per = (nb + (int)blockDim.x - 1) / (int)blockDim.x;
        const int b0 = threadIdx.x * per;
        uint32_t sum = 0;
        for (int b = 0; b < per; ++b)
            if (b0 + b < nb) sum += s_hist[b0 + b];
        uint32_t incl = sum;  // inclusive suffix sum within the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += y;
        }
        if (lane == 0) s_bc[8 + warp] = incl;  // warp total
        __syncthreads();
        uint32_t above = incl - sum;  // threads of this warp with higher bins
        for (int w2 = warp + 1; w2 < nwarps; ++w2) above += s_bc[8 + w2];
        if (above < (uint32_t)kk && above + sum >= (uint32_t)kk) {
            uint32_t cum = above;
            for (int b = per - 1; b >= 0; --b) {
                if (b0 + b >= nb) continue;
                const uint32_t c = s_hist[b0 + b];
                if (cum < (uint32_t)kk && cum + c >= (uint32_t)kk) {
                    s_bc[0] = (uint32_t)(b0 + b);
                    s_bc[1] = cum;  // participating keys with a larger digit at this level
                }
                cum += c;
            }
        }
        if (first_pass && k2 > 0 && above < (uint32_t)k2 && above + sum >= (uint32_t)k2) {
            uint32_t cum = above;
            for (int b = per - 1; b >= 0; --b) {
                if (b0 + b >= nb) continue;
                const uint32_t c = s_hist[b0 + b];
                if (cum < (uint32_t)k2 && cum + c >= (uint32_t)k2) s_bc[2] = prefix | ((uint32_t)(b0 + b) << shift);
                cum += c;
            }
        }
        __syncthreads();
        if (first_pass && front_key) *front_key = k2 > 0 ? s_bc[2] : 0u;
        first_pass = false;
        const uint32_t digit = s_bc[0];
        const uint32_t larger = s_bc[1];
        gt += (int)larger;
        kk -= (int)larger;
        prefix |= digit << shift;
        mask |= (uint32_t)(nb - 1) << shift;
        top = shift;
        __syncthreads();
    }
    *count_gt = gt;
    return prefix;
}

// bitonic sort, descending, of n (power of two) 64-bit keys in shared memory
__device__ void block_bitonic_desc_u64(unsigned long long *a, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
                int i = 2 * t - (t & (stride - 1));
                int j = i + stride;
                bool desc = ((i & size) == 0);
                unsigned long long x = a[i], y = a[j];
                if ((x < y) == desc) {
                    a[i] = y;
                    a[j] = x;
                }
            }
        }
    }
    __syncthreads();
}

// Error bound of the approximate (fp16 tensor-core) scores against the oracle's fp32 fmaf chain:
// |approx - exact| <= eps. rel == 0 means the scan engine was exact. With non-negative data the
// bound is relative to the score itself (sum |q_d c_d| equals the dot product); otherwise it is
// relative to ||q|| * max ||c||. See DESIGN.md "Exact top-k out of tensor cores".
struct EpsArgs {
    float rel;
    int nonneg;
    const float *qnorm;  // [nq] L2 norms of the queries
    float max_norm;      // max L2 norm of the stored vectors
};
__device__ __forceinline__ float band_eps(const EpsArgs &ea, float t, int q) {
    if (ea.rel <= 0.f) return 0.f;
    if (ea.nonneg) return ea.rel * fmaxf(t, 0.f) / (1.f - 2.f * ea.rel) + 4e-6f;
    return ea.rel * ea.qnorm[q] * ea.max_norm + 4e-6f;
}

// ======================================================================= dense -> sparse rows

__global__ void dense_count_kernel(const float *__restrict__ x, int64_t n, int d, int32_t *__restrict__ nnz,
                                   uint8_t *__restrict__ bad, int32_t *__restrict__ stats, int stat_base,
                                   float *__restrict__ row_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *r = x + row * d;
    int cnt = 0;
    bool isbad = false, neg = false;
    float ss = 0.f;
    for (int j = lane; j < d; j += 32) {
        float v = r[j];
        cnt += (v != 0.f) ? 1 : 0;  // NaN != 0 counts, but the row is dropped below
        isbad |= !isfinite(v);
        neg |= (v < 0.f);
        ss += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    isbad = __any_sync(0xffffffffu, isbad);
    neg = __any_sync(0xffffffffu, neg);
    if (lane == 0) {
        nnz[row] = isbad ? 0 : cnt;
        bad[row] = isbad ? 1 : 0;
        if (row_norm) row_norm[row] = isbad ? 0.f : sqrtf(ss);
        if (!isbad) {
            if (neg) atomicOr(&stats[stat_base], 1);
            atomicMax(&stats[stat_base + 1], __float_as_int(sqrtf(ss)));
        }
    }
}

__global__ void dense_fill_kernel(const float *__restrict__ x, int64_t n, int d, const int64_t *__restrict__ off,
                                  const uint8_t *__restrict__ bad, uint16_t *__restrict__ idx,
                                  float *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    if (bad[row]) return;
    const float *r = x + row * d;
    int64_t o = off[row];
    for (int base = 0; base < d; base += 32) {
        int j = base + lane;
        float v = j < d ? r[j] : 0.f;
        bool nz = v != 0.f;
        unsigned m = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            int pos = __popc(m & ((1u << lane) - 1u));
            idx[o + pos] = (uint16_t)j;
            val[o + pos] = v;
        }
        o += __popc(m);
    }
}

// ======================================================================= K2: exact coarse scores

constexpr int CO_WARPS = 8;
constexpr int CO_PAD = 33;

// MODE 0: scores[row * nlist + c]; MODE 1: best[row] = max over c of (score, -c) composite key
template <int MODE>
__global__ void __launch_bounds__(CO_WARPS * 32)
coarse_exact_kernel(const int64_t *__restrict__ r_off, const uint16_t *__restrict__ r_idx,
                    const float *__restrict__ r_val, int64_t row0, int64_t nrows, const float *__restrict__ cent,
                    int nlist, int d, float *__restrict__ scores, unsigned long long *__restrict__ best, int pitch) {
    extern __shared__ float s_c[];  // [d][CO_PAD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * 32;
    for (int c = warp; c < 32; c += CO_WARPS) {
        const int cc = c0 + c;
        for (int j = lane; j < d; j += 32) s_c[j * CO_PAD + c] = cc < nlist ? cent[(int64_t)cc * d + j] : 0.f;
    }
    __syncthreads();
    const int cmine = c0 + lane;
    for (int64_t r = (int64_t)blockIdx.y * CO_WARPS + warp; r < nrows; r += (int64_t)gridDim.y * CO_WARPS) {
        const int64_t b = r_off[row0 + r], e = r_off[row0 + r + 1];
        float acc = 0.f;
        for (int64_t t0 = b; t0 < e; t0 += 32) {
            int64_t t = t0 + lane;
            int ei = t < e ? (int)r_idx[t] : 0;
            float ev = t < e ? r_val[t] : 0.f;
            int cnt = (int)min((int64_t)32, e - t0);
            for (int u = 0; u < cnt; ++u) {
                int i = __shfl_sync(0xffffffffu, ei, u);
                float v = __shfl_sync(0xffffffffu, ev, u);
                acc = __fmaf_rn(v, s_c[i * CO_PAD + lane], acc);
            }
        }
        if (MODE == 0) {
            if (cmine < nlist) scores[(row0 + r) * pitch + cmine] = acc;
        } else {
            unsigned long long key = 0ull;
            if (cmine < nlist && acc == acc)
                key = ((unsigned long long)ivf_f2o(acc) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)cmine);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long y = __shfl_xor_sync(0xffffffffu, key, o);
                key = y > key ? y : key;
            }
            if (lane == 0 && key) atomicMax(&best[row0 + r], key);
        }
    }
}

template <int MODE>
static void launch_coarse(solo_handle *h, const IvfIndex &ix, const int64_t *r_off, const uint16_t *r_idx,
                          const float *r_val, int64_t row0, int64_t nrows, float *scores,
                          unsigned long long *best, int pitch = 0) {
    if (nrows <= 0) return;
    auto k = coarse_exact_kernel<MODE>;
    size_t smem = (size_t)ix.dim * CO_PAD * sizeof(float);
    SOLO_REQUIRE(smem <= 220 * 1024, SOLO_ECAPACITY, "dim %d too large for the coarse kernel", ix.dim);
    SOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int tiles = div_up(ix.nlist, 32);
    int per_sm = smem > 110 * 1024 ? 1 : 2;
    int want = kNumSMs * per_sm * 4;
    int ysplit = std::max(1, std::min(div_up(want, tiles), div_up(nrows, CO_WARPS)));
    ysplit = std::min(ysplit, 65535);
    dim3 grid(tiles, ysplit);
    k<<<grid, CO_WARPS * 32, smem, h->stream>>>(r_off, r_idx, r_val, row0, nrows, ix.cent.as<float>(), ix.nlist,
                                                ix.dim, scores, best, pitch ? pitch : ix.nlist);
    SOLO_CUDA(cudaGetLastError());
}

__global__ void decode_assign_kernel(const unsigned long long *__restrict__ best, const uint8_t *__restrict__ bad,
                                     int64_t n, int32_t *__restrict__ row_list) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = best[i];
    const bool isbad = bad ? (bad[i] != 0) : (row_list[i] < 0);  // re-assignment keeps skipped rows skipped
    row_list[i] = (isbad || k == 0ull) ? -1 : (int32_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
}

// ======================================================================= probe selection

constexpr int SEL_THREADS = 1024;
constexpr int SEL_BCAP = 1024;  // band entries re-scored on chip by select_probes_kernel

// one CTA per query; keys = ordered score; ties at the threshold resolved towards lower list id.
// With approximate (tensor-core) coarse scores (ea.rel > 0) the band around the nprobe-th score is
// re-scored exactly — the oracle's sequential fmaf over the query's non-zeros in ascending
// dimension — so the selected set is the exact one. `exact_all` re-scores everything that can be
// selected (sorted probe output carries exact scores).
struct SelectArgs {
    const float *scores;   // (nq, pitch)
    int pitch;
    int nlist;
    int nprobe;
    int32_t *probes;                  // (nq, nprobe) or null
    unsigned long long *probe_keys;   // (nq, nprobe) or null
    EpsArgs eps;
    int exact_all;
    const int64_t *q_off;             // sparse queries
    const uint16_t *q_idx;
    const float *q_val;
    const float *cent;                // (nlist, d) fp32
    int d;
    int n_front;                      // fast path: roughly the n_front best lists are written first
};

__device__ void select_probes_one(const SelectArgs &a, const int q) {
    extern __shared__ __align__(16) uint32_t s_sel_keys[];  // [nlist]
    uint32_t *s_keys = s_sel_keys;
    __shared__ uint32_t s_hist[KTH_BINS];
    __shared__ uint32_t s_bc[KTH_BC];
    __shared__ int s_cnt, s_eq_taken, s_back;
    const int nlist = a.nlist, nprobe = a.nprobe;
    int32_t *probes = a.probes;
    unsigned long long *probe_keys = a.probe_keys;
    const float *row = a.scores + (int64_t)q * a.pitch;
    {   // all loads of a thread in flight together (the row pitch is a multiple of 32 floats: 16-byte aligned)
        const float4 *row4 = reinterpret_cast<const float4 *>(row);
        const int n4 = nlist >> 2;
        for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                v[u] = i < n4 ? row4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < n4) {
                    uint4 k;
                    k.x = (v[u].x == v[u].x) ? ivf_f2o(v[u].x) : 0u;  // NaN -> lowest key
                    k.y = (v[u].y == v[u].y) ? ivf_f2o(v[u].y) : 0u;
                    k.z = (v[u].z == v[u].z) ? ivf_f2o(v[u].z) : 0u;
                    k.w = (v[u].w == v[u].w) ? ivf_f2o(v[u].w) : 0u;
                    reinterpret_cast<uint4 *>(s_keys)[i] = k;
                }
            }
        }
        for (int i = (n4 << 2) + threadIdx.x; i < nlist; i += blockDim.x) {
            const float s = row[i];
            s_keys[i] = (s == s) ? ivf_f2o(s) : 0u;
        }
    }
    if (threadIdx.x == 0) {
        s_cnt = 0;
        s_eq_taken = 0;
        s_back = 0;
    }
    __syncthreads();
    int gt;
    uint32_t fkey = 0u;
    uint32_t T = block_kth_largest_u32(s_keys, nlist, nprobe, false, 0u, s_hist, s_bc, &gt,
                                       min(a.n_front, nprobe), &fkey);
    if (a.eps.rel > 0.f && T != 0u && !a.exact_all) {
        // fast path: everything above the band is in; only the (small) band is re-scored exactly and
        // ranked under (exact score desc, list id asc)
        __shared__ int s_nband;
        __shared__ uint32_t s_band_id[SEL_BCAP];
        __shared__ unsigned long long s_band_key[SEL_BCAP];
        const float t = ivf_o2f(T);
        const float e2 = 2.f * band_eps(a.eps, t, q);
        const float lo = t - e2, hi = t + e2;
        const int64_t qb = a.q_off[q], qe = a.q_off[q + 1];
        if (threadIdx.x == 0) s_nband = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < nlist; i += blockDim.x) {
            const uint32_t k0 = s_keys[i];
            if (k0 == 0u) continue;
            const float s = ivf_o2f(k0);
            if (s > hi) {  // front: close lists first (they seed the scan's running threshold)
                const int slot = (a.n_front <= 0 || k0 >= fkey) ? atomicAdd(&s_cnt, 1) : nprobe - 1 - atomicAdd(&s_back, 1);
                probes[(int64_t)q * nprobe + slot] = i;
            } else if (s >= lo) {
                const int pos = atomicAdd(&s_nband, 1);
                if (pos < SEL_BCAP) s_band_id[pos] = (uint32_t)i;
            }
        }
        __syncthreads();
        const int nband = s_nband, certain = s_cnt + s_back;
        if (nband <= SEL_BCAP) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
            for (int b = warp; b < nband; b += nwarps) {
                const uint32_t i = s_band_id[b];
                const float *c = a.cent + (int64_t)i * a.d;
                float acc = 0.f;
                for (int64_t e0 = qb; e0 < qe; e0 += 32) {
                    const int64_t e = e0 + lane;
                    const float v = e < qe ? a.q_val[e] : 0.f;
                    const float cv = e < qe ? c[a.q_idx[e]] : 0.f;
                    const int cnt = (int)min((int64_t)32, qe - e0);
                    for (int u = 0; u < cnt; ++u)
                        acc = __fmaf_rn(__shfl_sync(0xffffffffu, v, u), __shfl_sync(0xffffffffu, cv, u), acc);
                }
                if (lane == 0)
                    s_band_key[b] = ((unsigned long long)((acc == acc) ? ivf_f2o(acc) : 0u) << 32) |
                                    (unsigned long long)(0xFFFFFFFFu - i);
            }
            __syncthreads();
            const int need = nprobe - certain;
            for (int b = threadIdx.x; b < nband; b += blockDim.x) {
                const unsigned long long key = s_band_key[b];
                int r = 0;
                for (int b2 = 0; b2 < nband; ++b2) r += s_band_key[b2] > key;
                if (r < need) probes[(int64_t)q * nprobe + (nprobe - 1 - atomicAdd(&s_back, 1))] = (int32_t)s_band_id[b];
            }
            return;
        }
        __syncthreads();  // band larger than the on-chip list: generic path below
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
    }
    if (a.eps.rel > 0.f && T != 0u) {
        const float t = ivf_o2f(T);
        const float e2 = 2.f * band_eps(a.eps, t, q);
        const float lo = t - e2, hi = t + e2;
        const int64_t qb = a.q_off[q], qe = a.q_off[q + 1];
        for (int i = threadIdx.x; i < nlist; i += blockDim.x) {
            const uint32_t k0 = s_keys[i];
            const float s = ivf_o2f(k0);
            uint32_t key;
            if (k0 == 0u || s < lo) key = 0u;                        // certainly out
            else if (s > hi && !a.exact_all) key = 0xFFFFFFFFu;      // certainly in
            else {
                const float *c = a.cent + (int64_t)i * a.d;
                float acc = 0.f;
                for (int64_t e = qb; e < qe; ++e) acc = __fmaf_rn(a.q_val[e], c[a.q_idx[e]], acc);
                key = (acc == acc) ? ivf_f2o(acc) : 0u;
            }
            s_keys[i] = key;
        }
        __syncthreads();
        T = block_kth_largest_u32(s_keys, nlist, nprobe, true, 0u, s_hist, s_bc, &gt);
    }
    const int need_eq = nprobe - gt;
    // strictly greater: any order
    for (int i = threadIdx.x; i < nlist; i += blockDim.x) {
        if (s_keys[i] > T) {
            int slot = atomicAdd(&s_cnt, 1);
            if (probes) probes[(int64_t)q * nprobe + slot] = i;
            if (probe_keys)
                probe_keys[(int64_t)q * nprobe + slot] =
                    ((unsigned long long)s_keys[i] << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
        }
    }
    __syncthreads();
    // equal to the threshold: lowest ids first (ordered, chunk by chunk)
    for (int base = 0; base < nlist; base += blockDim.x) {
        if (s_eq_taken >= need_eq) break;  // uniform: read after the barrier below
        int i = base + threadIdx.x;
        bool eq = i < nlist && s_keys[i] == T;
        unsigned m = __ballot_sync(0xffffffffu, eq);
        __shared__ int s_wsum[SEL_THREADS / 32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) s_wsum[warp] = __popc(m);
        __syncthreads();
        int before = s_eq_taken;
        for (int w = 0; w < warp; ++w) before += s_wsum[w];
        int mypos = before + __popc(m & ((1u << lane) - 1u));
        if (eq && mypos < need_eq) {
            int slot = gt + mypos;
            if (probes) probes[(int64_t)q * nprobe + slot] = i;
            if (probe_keys)
                probe_keys[(int64_t)q * nprobe + slot] =
                    ((unsigned long long)T << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_wsum[w];
            s_eq_taken += tot;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SEL_THREADS) select_probes_kernel(SelectArgs a) { select_probes_one(a, blockIdx.x); }

// the same for the queries of a device-side list (the compact path's fall-back): a few persistent CTAs
__global__ void __launch_bounds__(SEL_THREADS)
select_probes_listed_kernel(SelectArgs a, const int32_t *__restrict__ qlist, const int32_t *__restrict__ n_listed) {
    const int n = *n_listed;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        select_probes_one(a, qlist[i]);
        __syncthreads();
    }
}

// ---- compact probe selection: the coarse pass kept only scores >= tau[q] (a quantile estimated on a sample of the
// centroids), so a query's candidates are a few thousand (score, list) pairs instead of a dense row of nlist scores.

// tau[q] = the m-th largest of the query's scores against the first `ns` centroids (a random sample: k-means
// starts from randomly drawn rows)
__global__ void __launch_bounds__(128)
coarse_tau_kernel(const float *__restrict__ sample, int ld, int ns, int m, float *__restrict__ tau,
                  int32_t *__restrict__ ccnt, int32_t *__restrict__ n_fail) {
    extern __shared__ uint32_t s_tau_keys[];
    __shared__ uint32_t s_hist[KTH_BINS];
    __shared__ uint32_t s_bc[KTH_BC];
    const int q = blockIdx.x;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        const float v = sample[(int64_t)q * ld + i];
        s_tau_keys[i] = (v == v) ? ivf_f2o(v) : 0u;
    }
    if (q == 0 && threadIdx.x == 0) *n_fail = 0;
    __syncthreads();
    int gt;
    const uint32_t T = block_kth_largest_u32(s_tau_keys, ns, m, false, 0u, s_hist, s_bc, &gt);
    if (threadIdx.x == 0) {
        tau[q] = ivf_o2f(T);
        ccnt[q] = 0;
    }
}

struct CompactSelectArgs {
    const unsigned long long *cbuf;   // (nq, cap) (score bits << 32 | list)
    const int32_t *ccnt;
    const float *tau;
    int cap;
    int32_t *fail_list;               // queries that must go through the dense path
    int32_t *n_fail;
    SelectArgs s;                     // probes, eps, sparse queries, centroids, n_front
};

__global__ void __launch_bounds__(256) select_probes_compact_kernel(CompactSelectArgs ca) {
    extern __shared__ __align__(16) uint32_t s_ckeys[];  // [cap] keys, then [cap] list ids
    const SelectArgs &a = ca.s;
    uint32_t *s_keys = s_ckeys, *s_ids = s_ckeys + ca.cap;
    __shared__ uint32_t s_hist[KTH_BINS];
    __shared__ uint32_t s_bc[KTH_BC];
    __shared__ int s_cnt, s_back, s_nband;
    __shared__ uint32_t s_band_id[SEL_BCAP];
    __shared__ unsigned long long s_band_key[SEL_BCAP];
    const int q = blockIdx.x;
    const int nprobe = a.nprobe;
    const int n = ca.ccnt[q];
    auto fail = [&]() {
        if (threadIdx.x == 0) ca.fail_list[atomicAdd(ca.n_fail, 1)] = q;
    };
    if (n < nprobe || n > ca.cap) {   // the sampled threshold was too tight (or far too loose)
        fail();
        return;
    }
    const unsigned long long *row = ca.cbuf + (int64_t)q * ca.cap;
    {   // four loads in flight per thread (n is about 1.8 nprobe: seven 8-byte loads per thread otherwise one by one)
        const int step = blockDim.x;
        int i = threadIdx.x;
        for (; i + 3 * step < n; i += 4 * step) {
            unsigned long long e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e[u] = row[i + u * step];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s_keys[i + u * step] = ivf_f2o(__uint_as_float((uint32_t)(e[u] >> 32)));
                s_ids[i + u * step] = (uint32_t)(e[u] & 0xFFFFFFFFull);
            }
        }
        for (; i < n; i += step) {
            const unsigned long long e = row[i];
            s_keys[i] = ivf_f2o(__uint_as_float((uint32_t)(e >> 32)));
            s_ids[i] = (uint32_t)(e & 0xFFFFFFFFull);
        }
    }
    if (threadIdx.x == 0) {
        s_cnt = 0;
        s_back = 0;
        s_nband = 0;
    }
    __syncthreads();
    int gt;
    uint32_t fkey = 0u;
    const uint32_t T = block_kth_largest_u32(s_keys, n, nprobe, false, 0u, s_hist, s_bc, &gt, min(a.n_front, nprobe), &fkey);
    const float t = ivf_o2f(T);
    const float e2 = 2.f * band_eps(a.eps, t, q);
    const float lo = t - e2, hi = t + e2;
    if (!(lo >= ca.tau[q])) {         // part of the band may lie below the coarse pass's threshold
        fail();
        return;
    }
    int32_t *probes = a.probes + (int64_t)q * nprobe;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t k0 = s_keys[i];
        const float s = ivf_o2f(k0);
        if (s > hi) {   // front: close lists first (they seed the scan's running threshold)
            const int slot = (a.n_front <= 0 || k0 >= fkey) ? atomicAdd(&s_cnt, 1) : nprobe - 1 - atomicAdd(&s_back, 1);
            probes[slot] = (int32_t)s_ids[i];
        } else if (s >= lo) {
            const int pos = atomicAdd(&s_nband, 1);
            if (pos < SEL_BCAP) s_band_id[pos] = s_ids[i];
        }
    }
    __syncthreads();
    const int nband = s_nband, certain = s_cnt + s_back;
    if (nband > SEL_BCAP) {
        fail();
        return;
    }
    // the band is re-scored exactly (the oracle's sequential fmaf over the query's non-zeros) and ranked under
    // (exact score desc, list id asc)
    const int64_t qb = a.q_off[q], qe = a.q_off[q + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int b = warp; b < nband; b += nwarps) {
        const uint32_t i = s_band_id[b];
        const float *c = a.cent + (int64_t)i * a.d;
        float acc = 0.f;
        for (int64_t e0 = qb; e0 < qe; e0 += 32) {
            const int64_t e = e0 + lane;
            const float v = e < qe ? a.q_val[e] : 0.f;
            const float cv = e < qe ? c[a.q_idx[e]] : 0.f;
            const int cnt = (int)min((int64_t)32, qe - e0);
            for (int u = 0; u < cnt; ++u)
                acc = __fmaf_rn(__shfl_sync(0xffffffffu, v, u), __shfl_sync(0xffffffffu, cv, u), acc);
        }
        if (lane == 0)
            s_band_key[b] = ((unsigned long long)((acc == acc) ? ivf_f2o(acc) : 0u) << 32) | (unsigned long long)(0xFFFFFFFFu - i);
    }
    __syncthreads();
    const int need = nprobe - certain;
    for (int b = threadIdx.x; b < nband; b += blockDim.x) {
        const unsigned long long key = s_band_key[b];
        int r = 0;
        for (int b2 = 0; b2 < nband; ++b2) r += s_band_key[b2] > key;
        if (r < need) probes[nprobe - 1 - atomicAdd(&s_back, 1)] = (int32_t)s_band_id[b];
    }
}

// sort each row of n u64 keys descending (n <= 4096), decode to list ids
__global__ void __launch_bounds__(512)
sort_probe_rows_kernel(unsigned long long *__restrict__ keys, int n, int npad, int32_t *__restrict__ out) {
    extern __shared__ unsigned long long s_k[];
    const int q = blockIdx.x;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) s_k[i] = i < n ? keys[(int64_t)q * n + i] : 0ull;
    __syncthreads();
    block_bitonic_desc_u64(s_k, npad);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        out[(int64_t)q * n + i] = (int32_t)(0xFFFFFFFFu - (uint32_t)(s_k[i] & 0xFFFFFFFFull));
}

// ======================================================================= grouping (two rounds)

// r0[q] = number of leading probes whose lists together hold <= c0 vectors (at least one)
__global__ void round0_split_kernel(const int32_t *__restrict__ probes, int nq, int nprobe,
                                    const int64_t *__restrict__ list_off, int64_t c0, int32_t *__restrict__ r0) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int64_t acc = 0;
    int r = 0;
    for (; r < nprobe; ++r) {
        int l = probes[(int64_t)q * nprobe + r];
        int64_t len = list_off[l + 1] - list_off[l];
        if (r > 0 && acc + len > c0) break;
        acc += len;
    }
    r0[q] = r;
}

// (lists without vectors on this GPU — mode B: lists another GPU owns — get no group: 7 of 8 probes at 8 GPUs)
__global__ void group_count_kernel(const int32_t *__restrict__ probes, int nq, int nprobe,
                                   const int32_t *__restrict__ r0, int nlist, const int64_t *__restrict__ list_off,
                                   int32_t *__restrict__ gcnt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nq * nprobe) return;
    int q = (int)(t / nprobe), j = (int)(t % nprobe);
    int l = probes[t];
    if (list_off[l + 1] == list_off[l]) return;
    int round = j >= r0[q] ? 1 : 0;
    atomicAdd(&gcnt[round * nlist + l], 1);
}

__global__ void group_fill_kernel(const int32_t *__restrict__ probes, int nq, int nprobe,
                                  const int32_t *__restrict__ r0, int nlist, const int64_t *__restrict__ goff,
                                  const int64_t *__restrict__ list_off, int32_t *__restrict__ gcur, int32_t *__restrict__ gq) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nq * nprobe) return;
    int q = (int)(t / nprobe), j = (int)(t % nprobe);
    int l = probes[t];
    if (list_off[l + 1] == list_off[l]) return;
    int round = j >= r0[q] ? 1 : 0;
    int pos = atomicAdd(&gcur[round * nlist + l], 1);
    gq[goff[round * nlist + l] + pos] = q;
}

// ======================================================================= K3 engine v1 (CUDA cores, exact)

constexpr int EN_WARPS = 8;
constexpr int EN_VCH = 128;    // vectors per shared-memory chunk
constexpr int EN_ENT = 6144;   // sparse entries per chunk

struct ScanArgs {
    const int64_t *goff;   // [nlist+1] group offsets of this round
    const int32_t *gq;     // grouped query ids
    const int64_t *list_off;
    const int64_t *sp_off;
    const uint16_t *sp_idx;
    const float *sp_val;
    const float *q;        // (nq, d)
    int d;
    const float *tau;      // [nq] append threshold (score >= tau)
    unsigned long long *buf;  // [nq][cap] (score bits << 32 | position)
    int32_t *cnt;          // [nq]
    int cap;
};

__global__ void __launch_bounds__(EN_WARPS * 32) scan_exact_kernel(ScanArgs a) {
    extern __shared__ __align__(16) unsigned char en_smem[];
    uint2 *s_ent = reinterpret_cast<uint2 *>(en_smem);                         // [EN_ENT]
    int *s_voff = reinterpret_cast<int *>(s_ent + EN_ENT);                     // [EN_VCH + 1]
    float *s_q = reinterpret_cast<float *>(s_voff + EN_VCH + 4);               // [EN_WARPS][d]
    __shared__ int s_nv;
    const int l = blockIdx.x;
    const int64_t g0 = a.goff[l], g1 = a.goff[l + 1];
    const int64_t G = g1 - g0;
    if (G == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *myq = s_q + (size_t)warp * a.d;
    int64_t p0 = a.list_off[l];
    const int64_t pend = a.list_off[l + 1];
    while (p0 < pend) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int nv = 0;
            const int64_t e0 = a.sp_off[p0];
            while (nv < EN_VCH && p0 + nv < pend && a.sp_off[p0 + nv + 1] - e0 <= EN_ENT) ++nv;
            s_nv = nv;
        }
        __syncthreads();
        const int nv = s_nv;
        if (nv == 0) break;  // a single row larger than the chunk: cannot happen for dim <= EN_ENT
        const int64_t e0 = a.sp_off[p0];
        const int nent = (int)(a.sp_off[p0 + nv] - e0);
        for (int t = threadIdx.x; t < nent; t += blockDim.x)
            s_ent[t] = make_uint2((unsigned)a.sp_idx[e0 + t], __float_as_uint(a.sp_val[e0 + t]));
        for (int v = threadIdx.x; v <= nv; v += blockDim.x) s_voff[v] = (int)(a.sp_off[p0 + v] - e0);
        __syncthreads();
        for (int64_t g = (int64_t)blockIdx.y * EN_WARPS + warp; g < G; g += (int64_t)gridDim.y * EN_WARPS) {
            const int q = a.gq[g0 + g];
            const float tau = a.tau[q];
            const float *qrow = a.q + (int64_t)q * a.d;
            for (int j = lane; j < a.d; j += 32) myq[j] = qrow[j];
            __syncwarp();
            for (int vb = 0; vb < nv; vb += 32) {
                const int v = vb + lane;
                float acc = 0.f;
                bool pass = false;
                if (v < nv) {
                    const int b = s_voff[v], e = s_voff[v + 1];
                    for (int t = b; t < e; ++t) {
                        uint2 en = s_ent[t];
                        acc = __fmaf_rn(myq[en.x], __uint_as_float(en.y), acc);
                    }
                    pass = acc >= tau;  // NaN never passes
                }
                unsigned m = __ballot_sync(0xffffffffu, pass);
                if (m) {
                    int leader = __ffs(m) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(&a.cnt[q], __popc(m));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (pass) {
                        int slot = base + __popc(m & ((1u << lane) - 1u));
                        if (slot < a.cap)
                            a.buf[(int64_t)q * a.cap + slot] =
                                ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned long long)(uint32_t)(p0 + v);
                    }
                }
            }
            __syncwarp();
        }
        p0 += nv;
    }
}

// ======================================================================= K4: thresholds and top-k

constexpr int TK_THREADS = 1024;

// After round 0: tau[q] = k-th best score so far (or -inf), buffer compacted to scores >= tau - margin.
// retry != 0: used after an overflow — recompute tau from the (full) buffer, keep only the
// round-0 part [0, n0) that still passes, mark non-overflowed queries done (tau = +inf).
// candidate buffer -> decision keys in shared memory, four loads in flight per thread (the plain strided loop spent a
// quarter of K4's samples waiting on one 8-byte load at a time)
__device__ __forceinline__ void load_score_keys(const unsigned long long *__restrict__ b, int n, uint32_t *s_keys) {
    const int step = blockDim.x;
    int i = threadIdx.x;
    for (; i + 3 * step < n; i += 4 * step) {
        const unsigned long long e0 = b[i], e1 = b[i + step], e2 = b[i + 2 * step], e3 = b[i + 3 * step];
        s_keys[i] = ivf_f2o(__uint_as_float((uint32_t)(e0 >> 32)));
        s_keys[i + step] = ivf_f2o(__uint_as_float((uint32_t)(e1 >> 32)));
        s_keys[i + 2 * step] = ivf_f2o(__uint_as_float((uint32_t)(e2 >> 32)));
        s_keys[i + 3 * step] = ivf_f2o(__uint_as_float((uint32_t)(e3 >> 32)));
    }
    for (; i < n; i += step) s_keys[i] = ivf_f2o(__uint_as_float((uint32_t)(b[i] >> 32)));
}

__global__ void __launch_bounds__(TK_THREADS)
threshold_kernel(unsigned long long *__restrict__ buf, int32_t *__restrict__ cnt, int32_t *__restrict__ n0,
                 float *__restrict__ tau, int cap, int scap, int k, EpsArgs ea, int retry) {
    extern __shared__ uint32_t s_keys[];  // [scap] (scap <= 32 * blockDim.x)
    __shared__ uint32_t s_hist[KTH_BINS];
    __shared__ uint32_t s_bc[KTH_BC];
    __shared__ int s_out;
    const int q = blockIdx.x;
    unsigned long long *b = buf + (int64_t)q * cap;
    const int raw = cnt[q];
    if (retry) {
        if (raw <= cap) {  // this query finished cleanly in the previous attempt
            if (threadIdx.x == 0) tau[q] = INFINITY;
            return;
        }
    }
    const int n = min(raw, cap);
    if (n < k || n > scap) {  // too few to bound the k-th score — or more than this launch holds: no threshold (still exact)
        if (threadIdx.x == 0) {
            tau[q] = -INFINITY;
            n0[q] = n;
        }
        return;
    }
    load_score_keys(b, n, s_keys);
    if (threadIdx.x == 0) s_out = 0;
    __syncthreads();
    int gt;
    const uint32_t T = block_kth_largest_u32(s_keys, n, k, false, 0u, s_hist, s_bc, &gt);
    const float t = ivf_o2f(T);
    // every member of the exact top-k has approx >= (final approx k-th) - 2 eps >= t - 2 eps
    const float thr = t - 2.f * band_eps(ea, t, q);
    const int limit = retry ? n0[q] : n;  // retry: only round-0 entries survive
    // compaction through registers (each thread owns a strided subset; two-phase to stay in place)
    unsigned long long keep[32];
    int nk = 0;
    for (int i = threadIdx.x, r = 0; i < limit && r < 32; i += blockDim.x, ++r) {
        unsigned long long e = b[i];
        if (__uint_as_float((uint32_t)(e >> 32)) >= thr) keep[nk++] = e;
    }
    __syncthreads();
    int base = nk ? atomicAdd(&s_out, nk) : 0;
    __syncthreads();
    for (int r = 0; r < nk; ++r) b[base + r] = keep[r];
    __syncthreads();
    if (threadIdx.x == 0) {
        tau[q] = thr;
        cnt[q] = s_out;
        n0[q] = s_out;
    }
}

struct FinalArgs {
    const unsigned long long *buf;
    const int32_t *cnt;
    const int32_t *list_ids;
    int cap;
    int k;
    int scap;              // entries this launch holds in shared memory; queries with more are deferred
    int min_cnt;           // process only queries with more than min_cnt entries (second pass)
    int32_t *overflow;     // [0] incremented when cnt > cap, [1] when deferred (cnt > scap)
    // sorted API output
    int64_t *I;
    float *D;
    // fused output
    int32_t *sel_ids;
    int32_t *sel_cnt;
    // mode B exchange format: (nq, k) entries (float score bits << 32 | row id), unsorted, padded with IVF_PACKED_PAD;
    // the score is the approximate one (exact for entries that went through the band) — within eps of the exact score
    unsigned long long *packed;
    // optional precursor-window mask (applied after the top-k)
    const double *q_prec_mz;
    const float *lib_prec_mz32;
    const uint8_t *lib_valid;
    int charge;
    double tol;
    int tol_mode;          // -1: no mask
    // exact re-scoring of the band around the k-th approximate score (tensor-core engine)
    EpsArgs eps;
    const float *q;        // (nq, d) fp32 queries
    int d;
    const int64_t *sp_off; // list-ordered sparse rows
    const uint16_t *sp_idx;
    const float *sp_val;
};

// exact score of list position `pos` against the query held in shared memory: the oracle's
// sequential fp32 fmaf over ascending dimensions (zeros of the stored row skipped)
__device__ __forceinline__ float exact_score(const FinalArgs &a, const float *s_q, uint32_t pos) {
    float acc = 0.f;
    const int64_t b = a.sp_off[pos], e = a.sp_off[pos + 1];
    for (int64_t t = b; t < e; ++t) acc = __fmaf_rn(s_q[a.sp_idx[t]], a.sp_val[t], acc);
    return acc;
}

__device__ __forceinline__ bool window_pass(const FinalArgs &a, int q, int id) {
    if (a.tol_mode < 0) return true;
    const double qm = a.q_prec_mz[q];
    const double lm = (double)a.lib_prec_mz32[id];
    bool ok;
    // spectral_library.py:421-427, numexpr evaluates in float64
    if (a.tol_mode == SOLO_TOL_DA) ok = __dmul_rn(fabs(__dsub_rn(qm, lm)), (double)a.charge) <= a.tol;
    else ok = __dmul_rn(__ddiv_rn(fabs(__dsub_rn(qm, lm)), lm), 1000000.0) <= a.tol;
    return ok && a.lib_valid[id];  // spectral_library.py:453
}

// K4: per query, the exact top-k (score desc, id asc) of everything the scan appended.
//  1. T = k-th largest APPROXIMATE score; eps = error bound at T.
//  2. entries above T + 2 eps are certainly in, entries below T - 2 eps certainly out; the band in
//     between is re-scored EXACTLY (sorted-output mode re-scores everything that can be in, because
//     it returns exact scores D).
//  3. the remaining slots are filled with the best band entries under (exact score desc, id asc).
// With an exact scan engine (eps.rel == 0) step 2 is the identity.
constexpr int TK_BCAP = 1024;  // band entries re-scored on chip by the fast path of final_topk_kernel

// (two 1024-thread CTAs per SM: 32 registers per thread, so that three 512-thread CTAs of the common launch stay resident)
template <bool PACKED>
__global__ void __launch_bounds__(TK_THREADS, 2) final_topk_kernel(FinalArgs a) {
    extern __shared__ uint32_t s_keys[];  // [scap] decision keys; s_q [d] follows; then s_sorted [IVF_MAX_K] (sorted output)
    __shared__ uint32_t s_hist[KTH_BINS];
    __shared__ uint32_t s_bc[KTH_BC];
    __shared__ int s_nout, s_neq, s_nband, s_ncert;
    __shared__ uint32_t s_band_pos[TK_BCAP];
    __shared__ unsigned long long s_band_key[TK_BCAP];
    float *s_q = reinterpret_cast<float *>(s_keys + a.scap);
    unsigned long long *s_sorted = reinterpret_cast<unsigned long long *>(s_q + ((a.d + 1) & ~1));
    const int q = blockIdx.x;
    const unsigned long long *b = a.buf + (int64_t)q * a.cap;
    const int raw = a.cnt[q];
    if (raw <= a.min_cnt) return;  // finished by the first pass
    if (raw > a.cap) {
        if (threadIdx.x == 0) atomicAdd(a.overflow, 1);
        return;
    }
    if (raw > a.scap) {
        if (threadIdx.x == 0) atomicAdd(a.overflow + 1, 1);
        return;
    }
    const int n = raw;
    const int kk = min(a.k, n);
    const bool sorted_out = a.I != nullptr;
    // list position -> library row (the merge of mode B runs on row ids directly: list_ids == null)
    auto row_of = [&](uint32_t pos) -> int { return a.list_ids ? a.list_ids[pos] : (int)pos; };
    auto emit = [&](int id, float score) {
        if (PACKED)
            a.packed[(int64_t)q * a.k + atomicAdd(&s_nout, 1)] =
                ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(uint32_t)id;
        else if (window_pass(a, q, id))
            a.sel_ids[(int64_t)q * a.k + atomicAdd(&s_nout, 1)] = id;
    };
    if (threadIdx.x == 0) {
        s_nout = 0;
        s_neq = 0;
        s_nband = 0;
        s_ncert = 0;
    }
    load_score_keys(b, n, s_keys);
    __syncthreads();
    if (kk > 0 && a.eps.rel > 0.f && !sorted_out) {
        // ---- fast path: everything above the band is in; the (small) band is re-scored exactly, one
        // warp per entry, and ranked under (exact score desc, id asc)
        int gt;
        const uint32_t T = block_kth_largest_u32(s_keys, n, kk, false, 0u, s_hist, s_bc, &gt);
        const float t = ivf_o2f(T);
        const float e2 = 2.f * band_eps(a.eps, t, q);
        const float lo = t - e2, hi = t + e2;
        for (int j = threadIdx.x; j < a.d; j += blockDim.x) s_q[j] = a.q[(int64_t)q * a.d + j];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float s = ivf_o2f(s_keys[i]);
            if (s > hi) {
                atomicAdd(&s_ncert, 1);
                emit(row_of((uint32_t)(b[i] & 0xFFFFFFFFull)), s);
            } else if (s >= lo) {
                const int p = atomicAdd(&s_nband, 1);
                if (p < TK_BCAP) s_band_pos[p] = (uint32_t)(b[i] & 0xFFFFFFFFull);
            }
        }
        __syncthreads();
        const int nband = s_nband;
        if (nband <= TK_BCAP) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
            for (int w = warp; w < nband; w += nwarps) {
                const uint32_t pos = s_band_pos[w];
                const int64_t rb = a.sp_off[pos], re = a.sp_off[pos + 1];
                float acc = 0.f;
                for (int64_t e0 = rb; e0 < re; e0 += 32) {
                    const int64_t e = e0 + lane;
                    const float qv = e < re ? s_q[a.sp_idx[e]] : 0.f;
                    const float xv = e < re ? a.sp_val[e] : 0.f;
                    const int cnt = (int)min((int64_t)32, re - e0);
                    for (int u = 0; u < cnt; ++u)
                        acc = __fmaf_rn(__shfl_sync(0xffffffffu, qv, u), __shfl_sync(0xffffffffu, xv, u), acc);
                }
                if (lane == 0)
                    s_band_key[w] = ((unsigned long long)((acc == acc) ? ivf_f2o(acc) : 0u) << 32) |
                                    (unsigned long long)(0xFFFFFFFFu - (uint32_t)row_of(pos));
            }
            __syncthreads();
            const int need = kk - s_ncert;
            for (int w = threadIdx.x; w < nband; w += blockDim.x) {
                const unsigned long long key = s_band_key[w];
                int r = 0;
                for (int w2 = 0; w2 < nband; ++w2) r += s_band_key[w2] > key;
                if (r < need) emit((int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull)), ivf_o2f((uint32_t)(key >> 32)));
            }
            __syncthreads();
            if (PACKED) {
                for (int i = s_nout + threadIdx.x; i < a.k; i += blockDim.x) a.packed[(int64_t)q * a.k + i] = IVF_PACKED_PAD;
            } else if (threadIdx.x == 0) {
                a.sel_cnt[q] = s_nout;
            }
            return;
        }
        __syncthreads();  // band larger than the on-chip list: generic path below (s_keys still hold approximate keys)
        if (threadIdx.x == 0) s_nout = 0;
        __syncthreads();
    }
    if (kk > 0) {
        int gt;
        if (a.eps.rel > 0.f) {
            const uint32_t T = block_kth_largest_u32(s_keys, n, kk, false, 0u, s_hist, s_bc, &gt);
            const float t = ivf_o2f(T);
            const float e2 = 2.f * band_eps(a.eps, t, q);
            const float lo = t - e2, hi = t + e2;
            for (int j = threadIdx.x; j < a.d; j += blockDim.x) s_q[j] = a.q[(int64_t)q * a.d + j];
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned long long e = b[i];
                const float s = __uint_as_float((uint32_t)(e >> 32));
                uint32_t key;
                if (s < lo) key = 0u;                                  // certainly out
                else if (s > hi && !sorted_out) key = 0xFFFFFFFFu;     // certainly in
                else key = ivf_f2o(exact_score(a, s_q, (uint32_t)(e & 0xFFFFFFFFull)));
                s_keys[i] = key;
            }
            __syncthreads();
        }
        // exact selection on the decision keys (0 = out)
        const uint32_t T2 = block_kth_largest_u32(s_keys, n, kk, true, 0u, s_hist, s_bc, &gt);
        const int need_eq = kk - gt;
        int n_eq_local = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t key = s_keys[i];
            n_eq_local += (key == T2);
            if (key > T2) {  // strictly better than the k-th: in
                const unsigned long long e = b[i];
                const int id = row_of((uint32_t)(e & 0xFFFFFFFFull));
                if (sorted_out) {
                    int slot = atomicAdd(&s_nout, 1);
                    s_sorted[slot] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)id);
                } else {   // a "certainly in" key carries no score: the entry's own (approximate) one is reported
                    emit(id, PACKED ? (key == 0xFFFFFFFFu ? __uint_as_float((uint32_t)(e >> 32)) : ivf_o2f(key)) : 0.f);
                }
            }
        }
        if (n_eq_local) atomicAdd(&s_neq, n_eq_local);
        __syncthreads();
        const int neq = s_neq;
        // ties at the k-th score: the need_eq lowest ids win
        uint32_t Tid = 1u;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t key = s_keys[i];
            s_keys[i] = (key == T2) ? 0xFFFFFFFEu - (uint32_t)row_of((uint32_t)(b[i] & 0xFFFFFFFFull)) : 0u;
        }
        __syncthreads();
        if (neq > need_eq) {
            int gt2;
            Tid = block_kth_largest_u32(s_keys, n, need_eq, true, 0u, s_hist, s_bc, &gt2);
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t key = s_keys[i];
            if (key != 0u && key >= Tid) {
                const int id = (int)(0xFFFFFFFEu - key);
                if (sorted_out) {
                    int slot = atomicAdd(&s_nout, 1);
                    s_sorted[slot] = ((unsigned long long)T2 << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)id);
                } else {
                    emit(id, ivf_o2f(T2));
                }
            }
        }
    }
    __syncthreads();
    if (!sorted_out) {
        if (PACKED) {
            for (int i = s_nout + threadIdx.x; i < a.k; i += blockDim.x) a.packed[(int64_t)q * a.k + i] = IVF_PACKED_PAD;
        } else if (threadIdx.x == 0) {
            a.sel_cnt[q] = s_nout;
        }
        return;
    }
    int npad = 1;
    while (npad < kk) npad <<= 1;
    for (int i = kk + threadIdx.x; i < npad; i += blockDim.x) s_sorted[i] = 0ull;
    __syncthreads();
    if (npad > 1) block_bitonic_desc_u64(s_sorted, npad);
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) {
        if (i < kk) {
            unsigned long long key = s_sorted[i];
            a.I[(int64_t)q * a.k + i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
            if (a.D) a.D[(int64_t)q * a.k + i] = ivf_o2f((uint32_t)(key >> 32));
        } else {
            a.I[(int64_t)q * a.k + i] = -1;
            if (a.D) a.D[(int64_t)q * a.k + i] = -INFINITY;
        }
    }
}


// ======================================================================= list-order build

__global__ void gather_len_kernel(const int32_t *__restrict__ list_ids, int64_t n, const int64_t *__restrict__ row_off,
                                  int32_t *__restrict__ len) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int r = list_ids[p];
    len[p] = (int32_t)(row_off[r + 1] - row_off[r]);
}

// one warp per list position: copy the sparse row, write the scaled fp16 dense row
__global__ void build_list_order_kernel(const int32_t *__restrict__ list_ids, int64_t n,
                                        const int64_t *__restrict__ row_off, const uint16_t *__restrict__ row_idx,
                                        const float *__restrict__ row_val, const int64_t *__restrict__ sp_off,
                                        uint16_t *__restrict__ sp_idx, float *__restrict__ sp_val,
                                        __half *__restrict__ vec_h, int d, float scale) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n) return;
    const int r = list_ids[p];
    const int64_t b = row_off[r], e = row_off[r + 1], o = sp_off[p];
    __half *dst = vec_h ? vec_h + p * (int64_t)d : nullptr;
    if (dst)
        for (int j = lane; j < d; j += 32) dst[j] = __float2half_rn(0.f);
    __syncwarp();
    for (int64_t t = b + lane; t < e; t += 32) {
        uint16_t i = row_idx[t];
        float v = row_val[t];
        sp_idx[o + (t - b)] = i;
        sp_val[o + (t - b)] = v;
        if (dst) dst[i] = __float2half_rn(v * scale);
    }
}

__global__ void f32_to_f16_scaled_kernel(const float *__restrict__ x, int64_t n, __half *__restrict__ y, float scale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2half_rn(x[i] * scale);
}

// ======================================================================= host side

void ivf_reset(IvfIndex &ix) {
    ix.ntotal = 0;
    ix.nstored = 0;
    ix.nnz = 0;
    ix.dirty = true;
    ix.max_norm = 0.f;
    ix.max_list_len = 0;
    ix.h_list_off.clear();
}

void ivf_set_centroids(solo_handle *h, IvfIndex &ix, const float *h_cent, int nlist, int dim) {
    SOLO_REQUIRE(nlist > 0 && nlist <= IVF_MAX_NLIST, SOLO_EINVAL, "nlist must be in [1, %d] (got %d)", IVF_MAX_NLIST,
                 nlist);
    SOLO_REQUIRE(dim > 0 && dim <= 1536, SOLO_EINVAL, "dim must be in [1, 1536] (got %d)", dim);
    ivf_reset(ix);  // new centroids define a new index: stored vectors are dropped
    ix.nlist = nlist;
    ix.dim = dim;
    size_t nb = (size_t)nlist * dim;
    ix.cent.ensure(nb * sizeof(float));
    ix.cent_h.ensure(nb * sizeof(__half));
    ix.stats.ensure(4 * sizeof(int32_t));
    SOLO_CUDA(cudaMemcpyAsync(ix.cent.p, h_cent, nb * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    SOLO_CUDA(cudaMemsetAsync(ix.stats.p, 0, 4 * sizeof(int32_t), h->stream));
    bool neg = false;
    double mx = 0;
    for (int c = 0; c < nlist; ++c) {
        double ss = 0;
        for (int j = 0; j < dim; ++j) {
            float v = h_cent[(size_t)c * dim + j];
            neg |= v < 0.f;
            ss += (double)v * v;
        }
        mx = std::max(mx, std::sqrt(ss));
    }
    ix.cent_max_norm = (float)mx;
    ix.nonneg = !neg;
    ix.cent_nonneg = !neg;
    // fp16 copy holds c * 2^s: |c| * 2^s <= 2^15, small values pushed away from the subnormal range
    ix.cent_scale_log2 = 10;
    if (mx > 0) ix.cent_scale_log2 = std::max(-14, std::min(10, (int)std::floor(std::log2(32768.0 / mx))));
    f32_to_f16_scaled_kernel<<<div_up(nb, 256), 256, 0, h->stream>>>(ix.cent.as<float>(), (int64_t)nb,
                                                                      ix.cent_h.as<__half>(),
                                                                      ldexpf(1.f, ix.cent_scale_log2));
    SOLO_CUDA(cudaGetLastError());
    h->launches++;
    ivf_reset(ix);
    ix.coarse_items_nq = -1;
    ix.coarse_items_used = -1;
    SOLO_CUDA(cudaStreamSynchronize(h->stream));
    tc_make_centroid_map(ix);
}

// sparse-convert `n` dense device rows and append them to the row-ordered store; returns nothing,
// assignment happens in ivf_add_device
static void append_sparse_rows(solo_handle *h, IvfIndex &ix, const float *d_x, int64_t n, DevBuf &cnt, DevBuf &bad) {
    const int d = ix.dim;
    cnt.ensure(n * sizeof(int32_t));
    bad.ensure(n);
    int rows_per_block = 8;
    dense_count_kernel<<<div_up(n, rows_per_block), rows_per_block * 32, 0, h->stream>>>(
        d_x, n, d, cnt.as<int32_t>(), bad.as<uint8_t>(), ix.stats.as<int32_t>(), 0, nullptr);
    SOLO_CUDA(cudaGetLastError());
    ix.row_off.ensure_keep((ix.ntotal + n + 1) * sizeof(int64_t), (ix.ntotal + 1) * sizeof(int64_t), h->stream);
    if (ix.ntotal == 0) SOLO_CUDA(cudaMemsetAsync(ix.row_off.p, 0, sizeof(int64_t), h->stream));
    scan_counts(h, cnt.as<int32_t>(), n, ix.row_off.as<int64_t>() + ix.ntotal, ix.nnz);
    int64_t new_nnz = 0;
    SOLO_CUDA(cudaMemcpyAsync(&new_nnz, ix.row_off.as<int64_t>() + ix.ntotal + n, sizeof(int64_t),
                              cudaMemcpyDeviceToHost, h->stream));
    SOLO_CUDA(cudaStreamSynchronize(h->stream));
    ix.row_idx.ensure_keep(new_nnz * sizeof(uint16_t) + 64, ix.nnz * sizeof(uint16_t), h->stream);
    ix.row_val.ensure_keep(new_nnz * sizeof(float) + 64, ix.nnz * sizeof(float), h->stream);
    dense_fill_kernel<<<div_up(n, rows_per_block), rows_per_block * 32, 0, h->stream>>>(
        d_x, n, d, ix.row_off.as<int64_t>() + ix.ntotal, bad.as<uint8_t>(), ix.row_idx.as<uint16_t>(),
        ix.row_val.as<float>());
    SOLO_CUDA(cudaGetLastError());
    h->launches += 2;
    ix.nnz = new_nnz;
}

__global__ void placeholder_assign_kernel(const uint8_t *__restrict__ bad, int64_t n, int32_t *__restrict__ row_list) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) row_list[i] = bad[i] ? -1 : 0;
}

void ivf_add_device(solo_handle *h, IvfIndex &ix, const float *d_x, int64_t n, bool assign) {
    SOLO_REQUIRE(ix.nlist > 0, SOLO_ESTATE, "index has no centroids (train or set_centroids first)");
    if (n <= 0) return;
    SOLO_REQUIRE(ix.ntotal + n < (int64_t)0x7fffffff, SOLO_ECAPACITY, "more than 2^31 rows");
    DevBuf &cnt = h->scratch[0], &bad = h->scratch[1], &best = h->scratch[2];
    const int64_t row0 = ix.ntotal;
    append_sparse_rows(h, ix, d_x, n, cnt, bad);
    if (!assign) {  // training: rows are stored, lists are decided by the Lloyd iterations
        ix.row_list.ensure_keep((row0 + n) * sizeof(int32_t), row0 * sizeof(int32_t), h->stream);
        placeholder_assign_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(bad.as<uint8_t>(), n,
                                                                        ix.row_list.as<int32_t>() + row0);
        SOLO_CUDA(cudaGetLastError());
        h->launches++;
        ix.ntotal += n;
        ix.dirty = true;
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        return;
    }
    best.ensure((row0 + n) * sizeof(unsigned long long));
    SOLO_CUDA(cudaMemsetAsync(best.as<unsigned long long>() + row0, 0, n * sizeof(unsigned long long), h->stream));
    launch_coarse<1>(h, ix, ix.row_off.as<int64_t>(), ix.row_idx.as<uint16_t>(), ix.row_val.as<float>(), row0, n,
                     nullptr, best.as<unsigned long long>());
    ix.row_list.ensure_keep((row0 + n) * sizeof(int32_t), row0 * sizeof(int32_t), h->stream);
    decode_assign_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(best.as<unsigned long long>() + row0,
                                                                bad.as<uint8_t>(), n, ix.row_list.as<int32_t>() + row0);
    SOLO_CUDA(cudaGetLastError());
    h->launches += 2;
    ix.ntotal += n;
    ix.dirty = true;
    SOLO_CUDA(cudaStreamSynchronize(h->stream));
}

void ivf_finalize(solo_handle *h, IvfIndex &ix) {
    if (!ix.dirty) return;
    const int64_t n = ix.ntotal;
    const int nlist = ix.nlist;
    {   // statistics gathered while rows were added: sign and largest norm decide the fp16 scale
        int32_t st[4] = {0, 0, 0, 0};
        SOLO_CUDA(cudaMemcpyAsync(st, ix.stats.p, sizeof st, cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        if (st[0]) ix.nonneg = false;
        memcpy(&ix.max_norm, &st[1], 4);
        // fp16 copies hold x * 2^scale: keep |x| * 2^scale <= 2^15 and push small values away from
        // the fp16 subnormal range (unit-norm rows: scale 10)
        int s = 10;
        if (ix.max_norm > 0.f) s = std::min(10, (int)std::floor(std::log2(32768.0 / (double)ix.max_norm)));
        ix.scale_log2 = std::max(s, -14);
    }
    std::vector<int32_t> row_list(n);
    if (n) {
        SOLO_CUDA(cudaMemcpyAsync(row_list.data(), ix.row_list.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    }
    // stable bucket by list: insertion order inside every list, as Faiss `add` leaves them
    ix.h_list_off.assign(nlist + 1, 0);
    const bool sharded = (int)ix.owned.size() == nlist;
    if (sharded)  // rows of lists another GPU owns are not stored (their ids stay reserved)
        for (int64_t i = 0; i < n; ++i)
            if (row_list[i] >= 0 && !ix.owned[row_list[i]]) row_list[i] = -1;
    for (int64_t i = 0; i < n; ++i)
        if (row_list[i] >= 0) ix.h_list_off[row_list[i] + 1]++;
    ix.max_list_len = 0;
    for (int l = 0; l < nlist; ++l) {
        ix.max_list_len = std::max(ix.max_list_len, ix.h_list_off[l + 1]);
        ix.h_list_off[l + 1] += ix.h_list_off[l];
    }
    const int64_t ns = ix.h_list_off[nlist];
    std::vector<int32_t> ids(std::max<int64_t>(ns, 1));
    {
        std::vector<int64_t> cur(ix.h_list_off.begin(), ix.h_list_off.end() - 1);
        for (int64_t i = 0; i < n; ++i)
            if (row_list[i] >= 0) ids[cur[row_list[i]]++] = (int32_t)i;
    }
    ix.nstored = ns;
    ix.list_off.ensure((nlist + 1) * sizeof(int64_t));
    SOLO_CUDA(cudaMemcpyAsync(ix.list_off.p, ix.h_list_off.data(), (nlist + 1) * sizeof(int64_t),
                              cudaMemcpyHostToDevice, h->stream));
    ix.list_ids.ensure(std::max<int64_t>(ns, 1) * sizeof(int32_t));
    ix.sp_off.ensure((ns + 1) * sizeof(int64_t));
    ix.sp_idx.ensure(std::max<int64_t>(ix.nnz, 1) * sizeof(uint16_t) + 64);
    ix.sp_val.ensure(std::max<int64_t>(ix.nnz, 1) * sizeof(float) + 64);
    ix.vec_h.ensure(std::max<int64_t>(ns, 1) * ix.dim * sizeof(__half));
    if (ns) {
        SOLO_CUDA(cudaMemcpyAsync(ix.list_ids.p, ids.data(), ns * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        DevBuf &len = h->scratch[0];
        len.ensure(ns * sizeof(int32_t));
        gather_len_kernel<<<div_up(ns, 256), 256, 0, h->stream>>>(ix.list_ids.as<int32_t>(), ns,
                                                                  ix.row_off.as<int64_t>(), len.as<int32_t>());
        SOLO_CUDA(cudaGetLastError());
        scan_counts(h, len.as<int32_t>(), ns, ix.sp_off.as<int64_t>(), 0);
        build_list_order_kernel<<<div_up(ns, 8), 256, 0, h->stream>>>(
            ix.list_ids.as<int32_t>(), ns, ix.row_off.as<int64_t>(), ix.row_idx.as<uint16_t>(),
            ix.row_val.as<float>(), ix.sp_off.as<int64_t>(), ix.sp_idx.as<uint16_t>(), ix.sp_val.as<float>(),
            ix.vec_h.as<__half>(), ix.dim, ldexpf(1.f, ix.scale_log2));
        SOLO_CUDA(cudaGetLastError());
        h->launches += 2;
    } else {
        SOLO_CUDA(cudaMemsetAsync(ix.sp_off.p, 0, sizeof(int64_t), h->stream));
    }
    SOLO_CUDA(cudaStreamSynchronize(h->stream));
    tc_make_tensor_map(ix);  // vec_h may have moved
    ix.dirty = false;
}

// ----------------------------------------------------------------------- search

static void sparsify_queries(solo_handle *h, IvfIndex &ix, const float *d_q, int nq, DevBuf &q_off, DevBuf &q_idx,
                             DevBuf &q_val) {
    DevBuf &cnt = h->scratch[0], &bad = h->scratch[1];
    cnt.ensure((size_t)nq * sizeof(int32_t));
    bad.ensure(nq);
    DevBuf &qstats = h->scratch[3];
    qstats.ensure(4 * sizeof(int32_t));
    SOLO_CUDA(cudaMemsetAsync(qstats.p, 0, 4 * sizeof(int32_t), h->stream));
    DevBuf &qnorm = h->scratch[22];
    qnorm.ensure(std::max<size_t>((size_t)nq * sizeof(float), 16));
    dense_count_kernel<<<div_up(nq, 8), 256, 0, h->stream>>>(d_q, nq, ix.dim, cnt.as<int32_t>(), bad.as<uint8_t>(),
                                                            qstats.as<int32_t>(), 0, qnorm.as<float>());
    SOLO_CUDA(cudaGetLastError());
    q_off.ensure((size_t)(nq + 1) * sizeof(int64_t));
    scan_counts(h, cnt.as<int32_t>(), nq, q_off.as<int64_t>(), 0);
    // worst case every entry is non-zero
    q_idx.ensure((size_t)nq * ix.dim * sizeof(uint16_t) + 64);
    q_val.ensure((size_t)nq * ix.dim * sizeof(float) + 64);
    dense_fill_kernel<<<div_up(nq, 8), 256, 0, h->stream>>>(d_q, nq, ix.dim, q_off.as<int64_t>(), bad.as<uint8_t>(),
                                                           q_idx.as<uint16_t>(), q_val.as<float>());
    SOLO_CUDA(cudaGetLastError());
    h->launches += 2;
}

__global__ void fill_f32_kernel(float *p, int64_t n, float v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

void ivf_search(solo_handle *h, IvfIndex &ix, const IvfSearchArgs &a) {
    SOLO_REQUIRE(ix.nlist > 0, SOLO_ESTATE, "index has no centroids");
    ivf_finalize(h, ix);
    const int nq = a.nq, d = ix.dim, nlist = ix.nlist;
    if (nq <= 0) return;
    const int nprobe = std::min(a.nprobe, nlist);
    SOLO_REQUIRE(nprobe >= 1, SOLO_EINVAL, "nprobe must be >= 1");
    SOLO_REQUIRE(a.coarse_only || (a.k >= 1 && a.k <= IVF_MAX_K), SOLO_EINVAL, "k must be in [1, %d] (got %d)",
                 IVF_MAX_K, a.k);
    cudaStream_t st = h->stream;

    // scratch map: 4 q_off, 5 q_idx, 6 q_val, 7 coarse scores, 8 probes, 9 probe keys, 10 r0,
    //              11 gcnt/gcur, 12 goff, 13 gq, 14 tau, 15 cnt, 16 n0, 17 buf, 18 overflow
    DevBuf &q_off = h->scratch[4], &q_idx = h->scratch[5], &q_val = h->scratch[6];
    {
        StageTimer t(h, ST_COARSE, 0);
        sparsify_queries(h, ix, a.q, nq, q_off, q_idx, q_val);
    }
    // ---- engine choice: tcgen05 (fp16 inputs, fp32 accumulate, exact band re-rank) unless the
    // dimension is not a multiple of 16 or the CUDA-core engine is asked for (cross-check/debugging)
    static const bool env_exact = [] {
        const char *e = getenv("SOLO_SCAN_ENGINE");
        return e && strcmp(e, "exact") == 0;
    }();
    const bool want_tc = !env_exact && !h->opt_scan_exact && tc_scan_supported(ix);
    const bool use_tc = want_tc && ix.tmap_valid;
    const bool use_tc_coarse = want_tc && ix.tmap_cent_valid;
    EpsArgs ea;
    ea.rel = 0.f;
    ea.nonneg = 0;
    ea.qnorm = h->scratch[22].as<float>();
    ea.max_norm = ix.max_norm;
    int q_scale_log2 = 0;
    bool q_nonneg = false;
    DevBuf &qh = h->scratch[24], &tile_cnt = h->scratch[25], &tile_off = h->scratch[26], &items = h->scratch[27];
    DevBuf &qmask = h->scratch[28];
    if (use_tc || use_tc_coarse) {
        StageTimer t(h, ST_COARSE, 0);
        int32_t qst[4];
        SOLO_CUDA(cudaMemcpyAsync(qst, h->scratch[3].p, sizeof qst, cudaMemcpyDeviceToHost, st));
        SOLO_CUDA(cudaStreamSynchronize(st));
        float qmax;
        memcpy(&qmax, &qst[1], 4);
        q_scale_log2 = 10;
        if (qmax > 0.f) q_scale_log2 = std::max(-14, std::min(10, (int)std::floor(std::log2(32768.0 / (double)qmax))));
        q_nonneg = qst[0] == 0;
        qh.ensure((size_t)nq * d * sizeof(__half));
        qmask.ensure((size_t)nq * 8 * sizeof(uint32_t));
        tc_prepare_queries(h, ix, a.q, nq, q_scale_log2, qh.as<__half>(), qmask.as<uint32_t>());
    }

    DevBuf &probes = h->scratch[8], &pkeys = h->scratch[9];
    const int32_t *d_probes_c = a.given_probes;
    if (!a.given_probes) {
        probes.ensure((size_t)nq * nprobe * sizeof(int32_t));
        int32_t *d_probes = probes.as<int32_t>();
        d_probes_c = d_probes;
        DevBuf &coarse = h->scratch[7];
        const int pitch = (nlist + 31) & ~31;
        SelectArgs sel;
        memset(&sel, 0, sizeof sel);
        sel.pitch = pitch;
        sel.nlist = nlist;
        sel.nprobe = nprobe;
        sel.probes = a.sort_probes ? nullptr : d_probes;
        sel.eps = ea;
        if (use_tc_coarse) {
            sel.eps.rel = IVF_REL_EPS;
            sel.eps.nonneg = (ix.cent_nonneg && q_nonneg) ? 1 : 0;
            sel.eps.max_norm = ix.cent_max_norm;
        }
        sel.exact_all = a.sort_probes ? 1 : 0;
        sel.q_off = q_off.as<int64_t>();
        sel.q_idx = q_idx.as<uint16_t>();
        sel.q_val = q_val.as<float>();
        sel.cent = ix.cent.as<float>();
        sel.d = d;
        // lists that fit the unconditional first scan round (c0 scores, see below), with some slack
        sel.n_front = (ix.nstored > 0 && h->opt_front_probes)
                          ? (int)std::min<int64_t>(nprobe, std::max<int64_t>(1, 3 * (int64_t)h->opt_round0_scores * nlist / (4 * ix.nstored)))
                          : 0;
        const size_t sel_smem = (size_t)nlist * sizeof(uint32_t);
        SOLO_CUDA(cudaFuncSetAttribute(select_probes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
        // ---- compact path: no (Q, nlist) score matrix. A per-query threshold is estimated on a sample of the
        // centroids (the first NS: k-means starts from randomly drawn rows) so that about 1.75 nprobe scores pass;
        // the coarse pass keeps only those, the selection works on a few thousand pairs. A query whose threshold
        // turned out too tight (or whose band reaches below it) goes through the dense path, decided on the device.
        constexpr int NS = 512;
        const double expect = 1.75 * nprobe;
        const int m_rank = (int)std::ceil(NS * expect / nlist);
        int cap_c = 1024;
        while (cap_c < 2.2 * expect) cap_c <<= 1;
        static const bool env_dense = getenv("SOLO_COARSE_DENSE") != nullptr;
        const bool compact = use_tc_coarse && !a.sort_probes && !env_dense && h->opt_compact_probes && nlist >= 4 * NS &&
                             expect <= 0.5 * nlist && m_rank >= 16 && cap_c <= 8192;
        if (compact) {
            DevBuf &samp = h->scratch[29], &cbuf = h->scratch[30], &cmisc = h->scratch[31];
            samp.ensure((size_t)nq * NS * sizeof(float));
            cbuf.ensure((size_t)nq * cap_c * sizeof(unsigned long long));
            cmisc.ensure(((size_t)3 * nq + 4) * sizeof(int32_t));
            float *c_tau = cmisc.as<float>();
            int32_t *c_cnt = cmisc.as<int32_t>() + nq, *c_fail = cmisc.as<int32_t>() + 2 * (size_t)nq;
            int32_t *c_nfail = cmisc.as<int32_t>() + 3 * (size_t)nq;
            {
                StageTimer t(h, ST_COARSE, 3, 2.0 * nq * (double)nlist * d);
                launch_coarse_tc(h, ix, qh.as<__half>(), qmask.as<uint32_t>(), nq, q_scale_log2, samp.as<float>(), NS, NS);
                coarse_tau_kernel<<<nq, 128, NS * sizeof(uint32_t), st>>>(samp.as<float>(), NS, NS, m_rank, c_tau, c_cnt, c_nfail);
                SOLO_CUDA(cudaGetLastError());
                launch_coarse_tc(h, ix, qh.as<__half>(), qmask.as<uint32_t>(), nq, q_scale_log2, nullptr, 0, -1, c_tau,
                                 cbuf.as<unsigned long long>(), c_cnt, cap_c);
            }
            {
                StageTimer t(h, ST_PROBE_SELECT, 4);
                CompactSelectArgs ca;
                ca.cbuf = cbuf.as<unsigned long long>();
                ca.ccnt = c_cnt;
                ca.tau = c_tau;
                ca.cap = cap_c;
                ca.fail_list = c_fail;
                ca.n_fail = c_nfail;
                ca.s = sel;
                const size_t csm = (size_t)cap_c * 2 * sizeof(uint32_t);
                SOLO_CUDA(cudaFuncSetAttribute(select_probes_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
                select_probes_compact_kernel<<<nq, 256, csm, st>>>(ca);
                SOLO_CUDA(cudaGetLastError());
                // fall-back for the listed queries (usually none or a handful): dense rows + the dense selection
                coarse.ensure((size_t)nq * pitch * sizeof(float));
                launch_coarse_tc_listed(h, ix, qh.as<__half>(), qmask.as<uint32_t>(), q_scale_log2, coarse.as<float>(), pitch,
                                        c_fail, c_nfail);
                sel.scores = coarse.as<float>();
                SOLO_CUDA(cudaFuncSetAttribute(select_probes_listed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
                select_probes_listed_kernel<<<std::min(nq, 2 * kNumSMs), SEL_THREADS, sel_smem, st>>>(sel, c_fail, c_nfail);
                SOLO_CUDA(cudaGetLastError());
                h->compact_probe_batches++;
            }
        } else {
            coarse.ensure((size_t)nq * pitch * sizeof(float));
            {
                StageTimer t(h, ST_COARSE, 1, 2.0 * nq * (double)nlist * d);
                if (use_tc_coarse)
                    launch_coarse_tc(h, ix, qh.as<__half>(), qmask.as<uint32_t>(), nq, q_scale_log2, coarse.as<float>(), pitch);
                else
                    launch_coarse<0>(h, ix, q_off.as<int64_t>(), q_idx.as<uint16_t>(), q_val.as<float>(), 0, nq,
                                     coarse.as<float>(), nullptr, pitch);
            }
            {
                StageTimer t(h, ST_PROBE_SELECT, 1 + (a.sort_probes ? 1 : 0));
                unsigned long long *d_keys = nullptr;
                if (a.sort_probes) {
                    pkeys.ensure((size_t)nq * nprobe * sizeof(unsigned long long));
                    d_keys = pkeys.as<unsigned long long>();
                }
                sel.scores = coarse.as<float>();
                sel.probe_keys = d_keys;
                select_probes_kernel<<<nq, SEL_THREADS, sel_smem, st>>>(sel);
                SOLO_CUDA(cudaGetLastError());
                if (a.sort_probes) {
                    SOLO_REQUIRE(nprobe <= 4096, SOLO_ECAPACITY, "sorted probe output supports nprobe <= 4096");
                    int npad = 1;
                    while (npad < nprobe) npad <<= 1;
                    size_t sm2 = (size_t)npad * sizeof(unsigned long long);
                    SOLO_CUDA(cudaFuncSetAttribute(sort_probe_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
                    sort_probe_rows_kernel<<<nq, 512, sm2, st>>>(d_keys, nprobe, npad, d_probes);
                    SOLO_CUDA(cudaGetLastError());
                }
            }
        }
    }
    const int32_t *d_probes = d_probes_c;
    if (a.probes)
        SOLO_CUDA(cudaMemcpyAsync(a.probes, d_probes, (size_t)nq * nprobe * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (a.coarse_only) return;

    // ---- per-query candidate buffers
    const int cap = 32768;
    const int64_t c0 = std::max<int64_t>(h->opt_round0_scores, ix.max_list_len);  // at least one whole list
    SOLO_REQUIRE(ix.max_list_len <= IVF_ROUND0_SCORES, SOLO_ECAPACITY,
                 "an inverted list holds %lld vectors; the scan buffer supports lists up to %lld (use more lists)",
                 (long long)ix.max_list_len, (long long)IVF_ROUND0_SCORES);
    DevBuf &r0 = h->scratch[10], &gcnt = h->scratch[11], &goff = h->scratch[12], &gq = h->scratch[13];
    DevBuf &tau = h->scratch[14], &cnt = h->scratch[15], &n0 = h->scratch[16], &buf = h->scratch[17];
    DevBuf &ovf = h->scratch[18];
    r0.ensure((size_t)nq * sizeof(int32_t));
    gcnt.ensure((size_t)4 * nlist * sizeof(int32_t));
    goff.ensure((size_t)(2 * nlist + 2) * sizeof(int64_t));
    gq.ensure((size_t)nq * nprobe * sizeof(int32_t));
    tau.ensure((size_t)nq * sizeof(float));
    cnt.ensure((size_t)nq * sizeof(int32_t));
    n0.ensure((size_t)nq * sizeof(int32_t));
    buf.ensure((size_t)nq * cap * sizeof(unsigned long long));
    ovf.ensure(sizeof(int32_t) * 2);
    const int64_t npairs = (int64_t)nq * nprobe;
    {
        StageTimer t(h, ST_GROUP, 5);
        round0_split_kernel<<<div_up(nq, 128), 128, 0, st>>>(d_probes, nq, nprobe, ix.list_off.as<int64_t>(), c0,
                                                             r0.as<int32_t>());
        SOLO_CUDA(cudaMemsetAsync(gcnt.p, 0, (size_t)4 * nlist * sizeof(int32_t), st));
        group_count_kernel<<<div_up(npairs, 256), 256, 0, st>>>(d_probes, nq, nprobe, r0.as<int32_t>(), nlist,
                                                                ix.list_off.as<int64_t>(), gcnt.as<int32_t>());
        // one scan over [round0 lists | round1 lists]: goff[r*nlist + l]
        scan_counts_kernel<<<1, 1024, 0, st>>>(gcnt.as<int32_t>(), 2 * nlist, goff.as<int64_t>(), 0);
        group_fill_kernel<<<div_up(npairs, 256), 256, 0, st>>>(d_probes, nq, nprobe, r0.as<int32_t>(), nlist,
                                                               goff.as<int64_t>(), ix.list_off.as<int64_t>(),
                                                               gcnt.as<int32_t>() + 2 * nlist, gq.as<int32_t>());
        SOLO_CUDA(cudaGetLastError());
    }
    SOLO_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nq * sizeof(int32_t), st));
    SOLO_CUDA(cudaMemsetAsync(ovf.p, 0, 2 * sizeof(int32_t), st));
    fill_f32_kernel<<<div_up(nq, 256), 256, 0, st>>>(tau.as<float>(), nq, -INFINITY);
    h->launches++;

    if (use_tc) {
        ea.rel = IVF_REL_EPS;
        ea.nonneg = (ix.nonneg && q_nonneg) ? 1 : 0;
    }
    ix.last_eps_rel = ea.rel;
    ix.last_eps_nonneg = ea.nonneg;

    ScanArgs sa;
    sa.gq = gq.as<int32_t>();
    sa.list_off = ix.list_off.as<int64_t>();
    sa.sp_off = ix.sp_off.as<int64_t>();
    sa.sp_idx = ix.sp_idx.as<uint16_t>();
    sa.sp_val = ix.sp_val.as<float>();
    sa.q = a.q;
    sa.d = d;
    sa.tau = tau.as<float>();
    sa.buf = buf.as<unsigned long long>();
    sa.cnt = cnt.as<int32_t>();
    sa.cap = cap;
    const size_t en_smem = (size_t)EN_ENT * sizeof(uint2) + (EN_VCH + 4) * sizeof(int) + (size_t)EN_WARPS * d * sizeof(float);
    SOLO_CUDA(cudaFuncSetAttribute(scan_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)en_smem));
    const int ysplit = std::max(1, std::min(64, div_up(kNumSMs * 8, nlist)));
    // K4 shared memory: keys + the query (+ the sorted output rows); the common case runs with a small
    // key capacity and several CTAs per SM, the rare large query is deferred to a second, big launch
    const bool sorted_out = a.I != nullptr;
    auto tk_bytes = [&](int keys) {
        return (size_t)keys * sizeof(uint32_t) + (size_t)((d + 1) & ~1) * sizeof(float) +
               (sorted_out ? (size_t)IVF_MAX_K * sizeof(unsigned long long) : 0);
    };
    const int scap_thr = (int)std::min<int64_t>(cap, std::max<int64_t>(c0, a.k));   // round 0 appends <= c0 per query
    static const int env_scap = getenv("SOLO_K4_SCAP") ? atoi(getenv("SOLO_K4_SCAP")) : 12288;
    const int scap_fin = std::min(cap, std::max(1024, env_scap));
    const size_t tk_smem = tk_bytes(cap);
    SOLO_CUDA(cudaFuncSetAttribute(threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tk_smem));
    auto final_kern = a.packed ? final_topk_kernel<true> : final_topk_kernel<false>;
    SOLO_CUDA(cudaFuncSetAttribute(final_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tk_smem));
    // expected scanned vectors per query, for the flop figure of the stage
    const double scan_units = 2.0 * d * (double)nq * ((double)ix.nstored * nprobe / nlist);

    auto run_round = [&](int round) {
        sa.goff = goff.as<int64_t>() + (size_t)round * nlist;
        StageTimer t(h, ST_SCAN, 1, round == 0 ? scan_units : 0.0);
        if (use_tc) {
            launch_scan_tc(h, ix, sa.goff, sa.gq, qh.as<__half>(), qmask.as<uint32_t>(), q_scale_log2, sa.tau, sa.buf,
                           sa.cnt, cap, tile_cnt, tile_off, items, round == 0);
        } else {
            scan_exact_kernel<<<dim3(nlist, ysplit), EN_WARPS * 32, en_smem, st>>>(sa);
            SOLO_CUDA(cudaGetLastError());
        }
    };
    run_round(0);
    {
        StageTimer t(h, ST_TOPK, 1);
        threshold_kernel<<<nq, scap_thr <= 16384 ? 512 : TK_THREADS, tk_bytes(scap_thr), st>>>(
            buf.as<unsigned long long>(), cnt.as<int32_t>(), n0.as<int32_t>(), tau.as<float>(), cap, scap_thr, a.k, ea, 0);
        SOLO_CUDA(cudaGetLastError());
    }
    FinalArgs fa;
    memset(&fa, 0, sizeof fa);
    fa.buf = buf.as<unsigned long long>();
    fa.cnt = cnt.as<int32_t>();
    fa.list_ids = ix.list_ids.as<int32_t>();
    fa.cap = cap;
    fa.k = a.k;
    fa.overflow = ovf.as<int32_t>();
    fa.I = a.I;
    fa.D = a.D;
    fa.sel_ids = a.sel_ids;
    fa.sel_cnt = a.sel_cnt;
    fa.packed = a.packed;
    fa.tol_mode = -1;
    fa.eps = ea;
    fa.q = a.q;
    fa.d = d;
    fa.sp_off = ix.sp_off.as<int64_t>();
    fa.sp_idx = ix.sp_idx.as<uint16_t>();
    fa.sp_val = ix.sp_val.as<float>();
    if (a.I == nullptr && a.packed == nullptr) {
        SOLO_REQUIRE(a.sel_ids && a.sel_cnt, SOLO_EINVAL, "no output requested");
        fa.q_prec_mz = a.win_q_prec_mz;
        fa.lib_prec_mz32 = a.win_lib_prec_mz32;
        fa.lib_valid = a.win_lib_valid;
        fa.charge = a.win_charge;
        fa.tol = a.win_tol;
        fa.tol_mode = a.win_tol_mode;
    }
    for (int attempt = 0;; ++attempt) {
        run_round(1);
        {
            StageTimer t(h, ST_TOPK, 1);
            fa.scap = scap_fin;
            fa.min_cnt = -1;
            final_kern<<<nq, 512, tk_bytes(scap_fin), st>>>(fa);
            SOLO_CUDA(cudaGetLastError());
        }
        // A query whose candidate buffer overflowed was not finished: raise its threshold from
        // what the buffer holds and rescan (never truncate). One 8-byte read-back per batch:
        // [0] overflowed queries, [1] queries deferred to the large-capacity launch.
        int32_t n_flags[2] = {0, 0};
        SOLO_CUDA(cudaMemcpyAsync(n_flags, ovf.p, sizeof n_flags, cudaMemcpyDeviceToHost, st));
        SOLO_CUDA(cudaStreamSynchronize(st));
        if (n_flags[1] > 0) {
            SOLO_CUDA(cudaMemsetAsync(ovf.as<int32_t>() + 1, 0, sizeof(int32_t), st));
            StageTimer t(h, ST_TOPK, 1);
            fa.scap = cap;
            fa.min_cnt = scap_fin;
            final_kern<<<nq, TK_THREADS, tk_smem, st>>>(fa);
            SOLO_CUDA(cudaGetLastError());
            SOLO_CUDA(cudaMemcpyAsync(n_flags, ovf.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            SOLO_CUDA(cudaStreamSynchronize(st));
        }
        const int32_t n_over = n_flags[0];
        if (n_over == 0) break;
        SOLO_REQUIRE(attempt < 8, SOLO_ECAPACITY, "candidate buffer overflow persists for %d queries", n_over);
        SOLO_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int32_t), st));
        StageTimer t(h, ST_TOPK, 1);
        threshold_kernel<<<nq, TK_THREADS, tk_smem, st>>>(buf.as<unsigned long long>(), cnt.as<int32_t>(),
                                                          n0.as<int32_t>(), tau.as<float>(), cap, cap, a.k, ea, 1);
        SOLO_CUDA(cudaGetLastError());
    }
}

// ----------------------------------------------------------------------- mode B: merge of per-GPU top-k entries

// (parts, S, k) packed entries -> per query of the slice one contiguous run of the real ones + their count
__global__ void __launch_bounds__(256)
merge_pack_kernel(const unsigned long long *__restrict__ parts, int n_parts, int S, int k, int cap,
                  unsigned long long *__restrict__ buf, int32_t *__restrict__ cnt) {
    __shared__ int s_n;
    const int q = blockIdx.x;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int p = 0; p < n_parts; ++p) {
        const unsigned long long *row = parts + ((int64_t)p * S + q) * k;
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const unsigned long long e = row[i];
            if ((uint32_t)(e & 0xFFFFFFFFull) != 0xFFFFFFFFu) buf[(int64_t)q * cap + atomicAdd(&s_n, 1)] = e;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cnt[q] = s_n;
}

// Exact top-k of the union of the GPUs' local top-k entries for the queries [q_begin, q_begin + n): the same
// selection as K4 — k-th approximate score, entries above the band are in, the band is re-scored exactly, here
// from the insertion-order sparse rows every GPU keeps of the whole library — with the precursor window fused in.
void ivf_merge_select(solo_handle *h, IvfIndex &ix, const unsigned long long *d_parts, int n_parts, int S, int k,
                      const float *d_q_slice, const float *d_qnorm_slice, int n, const IvfSearchArgs &win,
                      int32_t *sel_ids, int32_t *sel_cnt) {
    if (n <= 0) return;
    const int cap = n_parts * k;
    SOLO_REQUIRE(cap <= 32768, SOLO_ECAPACITY, "%d parts x k = %d entries per query exceed the merge capacity 32768", n_parts, cap);
    DevBuf &buf = h->scratch[17], &cnt = h->scratch[15], &ovf = h->scratch[18];
    buf.ensure((size_t)n * cap * sizeof(unsigned long long));
    cnt.ensure((size_t)n * sizeof(int32_t));
    ovf.ensure(2 * sizeof(int32_t));
    StageTimer t(h, ST_TOPK, 2);
    SOLO_CUDA(cudaMemsetAsync(ovf.p, 0, 2 * sizeof(int32_t), h->stream));
    merge_pack_kernel<<<n, 256, 0, h->stream>>>(d_parts, n_parts, S, k, cap, buf.as<unsigned long long>(), cnt.as<int32_t>());
    SOLO_CUDA(cudaGetLastError());
    FinalArgs fa;
    memset(&fa, 0, sizeof fa);
    fa.buf = buf.as<unsigned long long>();
    fa.cnt = cnt.as<int32_t>();
    fa.list_ids = nullptr;          // entries carry library rows
    fa.cap = cap;
    fa.k = k;
    fa.scap = cap;
    fa.min_cnt = -1;
    fa.overflow = ovf.as<int32_t>();
    fa.sel_ids = sel_ids;
    fa.sel_cnt = sel_cnt;
    fa.q_prec_mz = win.win_q_prec_mz;
    fa.lib_prec_mz32 = win.win_lib_prec_mz32;
    fa.lib_valid = win.win_lib_valid;
    fa.charge = win.win_charge;
    fa.tol = win.win_tol;
    fa.tol_mode = win.win_tol_mode;
    fa.eps.rel = ix.last_eps_rel;
    fa.eps.nonneg = ix.last_eps_nonneg;
    fa.eps.qnorm = d_qnorm_slice;
    fa.eps.max_norm = ix.max_norm;
    fa.q = d_q_slice;
    fa.d = ix.dim;
    fa.sp_off = ix.row_off.as<int64_t>();
    fa.sp_idx = ix.row_idx.as<uint16_t>();
    fa.sp_val = ix.row_val.as<float>();
    const size_t smem = (size_t)cap * sizeof(uint32_t) + (size_t)((ix.dim + 1) & ~1) * sizeof(float);
    SOLO_CUDA(cudaFuncSetAttribute(final_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    final_topk_kernel<false><<<n, cap <= 16384 ? 512 : TK_THREADS, smem, h->stream>>>(fa);
    SOLO_CUDA(cudaGetLastError());
}

// ----------------------------------------------------------------------- k-means (train)

// one CTA per centroid: sum the sparse rows assigned to it (list order = deterministic), normalise
__global__ void __launch_bounds__(256)
centroid_update_kernel(const int64_t *__restrict__ list_off, const int32_t *__restrict__ list_ids,
                       const int64_t *__restrict__ row_off, const uint16_t *__restrict__ row_idx,
                       const float *__restrict__ row_val, int d, float *__restrict__ cent) {
    extern __shared__ double s_acc[];  // [d]
    __shared__ double s_red[8];
    const int c = blockIdx.x;
    const int64_t p0 = list_off[c], p1 = list_off[c + 1];
    if (p0 == p1) return;  // empty cluster keeps its previous centroid
    for (int j = threadIdx.x; j < d; j += blockDim.x) s_acc[j] = 0.0;
    __syncthreads();
    for (int64_t p = p0; p < p1; ++p) {
        const int r = list_ids[p];
        const int64_t b = row_off[r], e = row_off[r + 1];
        for (int64_t t = b + threadIdx.x; t < e; t += blockDim.x) s_acc[row_idx[t]] += (double)row_val[t];
        __syncthreads();
    }
    double ss = 0.0;
    for (int j = threadIdx.x; j < d; j += blockDim.x) ss += s_acc[j] * s_acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += s_red[w];
    const double inv = tot > 0.0 ? 1.0 / sqrt(tot) : 0.0;
    for (int j = threadIdx.x; j < d; j += blockDim.x) cent[(int64_t)c * d + j] = (float)(s_acc[j] * inv);
}

// selected sparse rows -> dense centroid rows
__global__ void densify_rows_kernel(const int32_t *__restrict__ rows, int nsel, const int64_t *__restrict__ row_off,
                                    const uint16_t *__restrict__ row_idx, const float *__restrict__ row_val, int d,
                                    float *__restrict__ out) {
    const int c = blockIdx.x;
    if (c >= nsel) return;
    for (int j = threadIdx.x; j < d; j += blockDim.x) out[(int64_t)c * d + j] = 0.f;
    __syncthreads();
    const int r = rows[c];
    for (int64_t t = row_off[r] + threadIdx.x; t < row_off[r + 1]; t += blockDim.x)
        out[(int64_t)c * d + row_idx[t]] = row_val[t];
}

// Spherical k-means (Faiss `train` for an inner-product IVF): the training rows are streamed to
// the device once (`fill(r0, m, dst)` puts rows [r0, r0+m) into the dense device chunk `dst`) and
// kept as sparse rows; every Lloyd iteration is one exact coarse-assignment pass plus one
// deterministic centroid update. Leaves only centroids behind, like Faiss.
// cent[dst] = cent[src] with every dimension scaled by 1 +- 1/1024 (signs from a hash of (copy, dimension)): the
// two lists share the source's vectors at the next assignment (Faiss' split_clusters uses the same perturbation)
__global__ void reseed_centroids_kernel(const int32_t *__restrict__ pairs, float *__restrict__ cent, int dim) {
    const int src = pairs[3 * blockIdx.x], dst = pairs[3 * blockIdx.x + 1], copy = pairs[3 * blockIdx.x + 2];
    for (int j = threadIdx.x; j < dim; j += blockDim.x) {
        uint32_t hsh = (uint32_t)(copy * 0x9E3779B1u) ^ (uint32_t)(j * 0x85EBCA77u);
        hsh ^= hsh >> 15;
        hsh *= 0x2C1B3C6Du;
        hsh ^= hsh >> 12;
        const float f = (hsh & 1u) ? 1.f + 1.f / 1024.f : 1.f - 1.f / 1024.f;
        cent[(int64_t)dst * dim + j] = __fmul_rn(cent[(int64_t)src * dim + j], f);
    }
}

void ivf_train_rows(solo_handle *h, IvfIndex &ix, int64_t n, int dim, int nlist, int iters, uint64_t seed,
                    const std::function<void(int64_t, int64_t, float *)> &fill) {
    SOLO_REQUIRE(n >= nlist, SOLO_EINVAL, "need at least nlist (%d) training rows, got %lld", nlist, (long long)n);
    SOLO_REQUIRE(nlist > 0 && nlist <= IVF_MAX_NLIST, SOLO_EINVAL, "nlist must be in [1, %d]", IVF_MAX_NLIST);
    SOLO_REQUIRE(dim > 0 && dim <= 1536, SOLO_EINVAL, "dim must be in [1, 1536] (got %d)", dim);
    ivf_reset(ix);
    ix.nlist = nlist;
    ix.dim = dim;
    ix.cent.ensure((size_t)nlist * dim * sizeof(float));
    ix.stats.ensure(4 * sizeof(int32_t));
    SOLO_CUDA(cudaMemsetAsync(ix.stats.p, 0, 4 * sizeof(int32_t), h->stream));
    DevBuf &xd = h->scratch[19];
    const int64_t chunk = 1 << 17;
    xd.ensure((size_t)std::min(n, chunk) * dim * sizeof(float));
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        int64_t m = std::min(chunk, n - r0);
        fill(r0, m, xd.as<float>());
        ivf_add_device(h, ix, xd.as<float>(), m, /*assign=*/false);
    }
    // initial centroids: nlist distinct storable rows, seeded partial Fisher-Yates
    std::vector<int32_t> row_list(n);
    SOLO_CUDA(cudaMemcpyAsync(row_list.data(), ix.row_list.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    SOLO_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<int32_t> good;
    good.reserve(n);
    for (int64_t i = 0; i < n; ++i)
        if (row_list[i] >= 0) good.push_back((int32_t)i);
    SOLO_REQUIRE((int64_t)good.size() >= nlist, SOLO_EINVAL, "not enough finite training rows");
    std::mt19937_64 rng(seed);
    for (int c = 0; c < nlist; ++c) {
        std::uniform_int_distribution<int64_t> dist(c, (int64_t)good.size() - 1);
        std::swap(good[c], good[dist(rng)]);
    }
    std::sort(good.begin(), good.begin() + nlist);
    DevBuf &sel = h->scratch[0];
    sel.ensure(nlist * sizeof(int32_t));
    SOLO_CUDA(cudaMemcpyAsync(sel.p, good.data(), nlist * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    densify_rows_kernel<<<nlist, 128, 0, h->stream>>>(sel.as<int32_t>(), nlist, ix.row_off.as<int64_t>(),
                                                      ix.row_idx.as<uint16_t>(), ix.row_val.as<float>(), dim,
                                                      ix.cent.as<float>());
    SOLO_CUDA(cudaGetLastError());
    h->launches++;
    auto lloyd_iteration = [&]() {
        DevBuf &best = h->scratch[2];
        best.ensure(n * sizeof(unsigned long long));
        SOLO_CUDA(cudaMemsetAsync(best.p, 0, n * sizeof(unsigned long long), h->stream));
        launch_coarse<1>(h, ix, ix.row_off.as<int64_t>(), ix.row_idx.as<uint16_t>(), ix.row_val.as<float>(), 0, n,
                         nullptr, best.as<unsigned long long>());
        decode_assign_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(best.as<unsigned long long>(), nullptr, n,
                                                                    ix.row_list.as<int32_t>());
        SOLO_CUDA(cudaGetLastError());
        ix.dirty = true;
        ivf_finalize(h, ix);
        centroid_update_kernel<<<nlist, 256, dim * sizeof(double), h->stream>>>(
            ix.list_off.as<int64_t>(), ix.list_ids.as<int32_t>(), ix.row_off.as<int64_t>(), ix.row_idx.as<uint16_t>(),
            ix.row_val.as<float>(), dim, ix.cent.as<float>());
        SOLO_CUDA(cudaGetLastError());
        h->launches += 3;
    };
    for (int it = 0; it < iters; ++it) lloyd_iteration();
    // ---- optional balancing rounds (solo_set_option("train_balance", R)): Lloyd on sparse high-dimensional data
    // leaves a heavy tail (C2: median 33, mean 92, max 1,251 vectors per list), which costs the list scan tiles
    // that are mostly empty. Like Faiss' split of empty clusters, the centroid of a list far below the mean is
    // re-seeded with a slightly perturbed copy of the centroid of a list far above it (several copies for very
    // long lists); one Lloyd iteration follows each round. Centroids are an input to parity, not a result.
    for (int round = 0; round < h->opt_train_balance && iters > 0; ++round) {
        std::vector<std::pair<int64_t, int>> sz(nlist);
        for (int l = 0; l < nlist; ++l) sz[l] = {ix.h_list_off[l + 1] - ix.h_list_off[l], l};
        std::sort(sz.begin(), sz.end());
        const double mean = (double)ix.nstored / nlist;
        std::vector<int32_t> pairs;  // (source list, re-seeded list, copy number)
        int lo_i = 0;
        for (int hi_i = nlist - 1; hi_i > lo_i && (double)sz[hi_i].first > 1.75 * mean; --hi_i) {
            const int copies = (int)std::min<double>(15.0, std::floor((double)sz[hi_i].first / (1.25 * mean))) - 1;
            for (int c = 0; c < std::max(copies, 1) && lo_i < hi_i && (double)sz[lo_i].first < mean / 3.0; ++c, ++lo_i) {
                pairs.push_back(sz[hi_i].second);
                pairs.push_back(sz[lo_i].second);
                pairs.push_back(c + 1);
            }
        }
        if (pairs.empty()) break;
        DevBuf &dp = h->scratch[0];
        dp.ensure(pairs.size() * sizeof(int32_t));
        SOLO_CUDA(cudaMemcpyAsync(dp.p, pairs.data(), pairs.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        reseed_centroids_kernel<<<(int)(pairs.size() / 3), 256, 0, h->stream>>>(dp.as<int32_t>(), ix.cent.as<float>(), dim);
        SOLO_CUDA(cudaGetLastError());
        lloyd_iteration();
        h->launches += 1;
    }
    std::vector<float> out((size_t)nlist * dim);
    SOLO_CUDA(cudaMemcpyAsync(out.data(), ix.cent.p, out.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    SOLO_CUDA(cudaStreamSynchronize(h->stream));
    ivf_reset(ix);
    ivf_set_centroids(h, ix, out.data(), nlist, dim);
}

void ivf_train(solo_handle *h, IvfIndex &ix, const float *h_x, int64_t n, int dim, int nlist, int iters,
               uint64_t seed) {
    ivf_train_rows(h, ix, n, dim, nlist, iters, seed, [&](int64_t r0, int64_t m, float *dst) {
        SOLO_CUDA(cudaMemcpyAsync(dst, h_x + r0 * dim, (size_t)m * dim * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    });
}

}  // namespace solo

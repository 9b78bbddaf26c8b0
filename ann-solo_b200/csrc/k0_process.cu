// K0: batched process_spectrum. Replaces the per-spectrum Python/numba
// flow of reference spectrum.py:57-119 — set_mz_range, validity checks (:14-36), remove_precursor_peak,
// filter_intensity(min_intensity, max_peaks), scale_intensity('root' | 'rank'), L2 norm (:39-54) — by
// one launch over a CSR batch of raw spectra: one CTA per spectrum, raw intensities staged in shared
// memory, output rows of at most max_peaks peaks (m/z ascending) in fixed-stride arrays.
//
// The five peak operations live in spectrum_utils (absent from the reference tree: parity unpinned,
// SURVEY.md §8c); their arithmetic is DEFINED by the oracle restatement oracle/solo_oracle.py
// process_spectrum_np, which this kernel reproduces bit for bit:
//   * comparisons against configuration scalars happen in the precision of the m/z array (NumPy
//     weak-scalar promotion: float32 arrays compare against float32-rounded scalars);
//   * filter_intensity keeps peaks with intensity > float32(min_intensity) * max and then the
//     max_peaks largest under (intensity, peak index) — a stable ascending argsort's tail;
//   * rank scaling gives max_peaks - (number of kept peaks greater under the same order);
//   * the norm is float32(sqrt(sum of float64 squares)) with NumPy's pairwise summation order.
//   * round(resolution, 'sum') (spectrum.py:84-89): NumPy's round — rint(x * 10^d) / 10^d in the m/z array's
//     precision —, peaks that land on the same value (adjacent, m/z is ascending) are merged: float32 sum
//     in peak order, the merged peak stands where the group's most intense member (first on ties) stood, so
//     that its annotation / fragment charge follows it.
#include "solo_common.cuh"

namespace solo {

constexpr int K0_THREADS = 128;
constexpr int K0_MAX_RAW = 8192;  // raw peaks per spectrum held in shared memory

// np.sum over float64 (pairwise_sum in NumPy's loops, n <= 128 here): < 8 sequential; otherwise eight
// interleaved accumulators over the multiple-of-8 prefix, combined as a tree, then the tail
__device__ double numpy_pairwise_sum(const double *a, int n) {
    if (n < 8) {
        double s = 0.0;   // NumPy starts from a[0]'s identity: res = 0. + a[0] ...
        for (int i = 0; i < n; ++i) s = __dadd_rn(s, a[i]);
        return s;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    double s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) s = __dadd_rn(s, a[i]);
    return s;
}

// np.round(x, d), d >= 0: rint(x * 10^d) / 10^d in the array's precision (PyArray_Round)
__device__ __forceinline__ float round_decimals(float x, float f) { return __fdiv_rn(rintf(__fmul_rn(x, f)), f); }
__device__ __forceinline__ double round_decimals(double x, double f) { return __ddiv_rn(rint(__dmul_rn(x, f)), f); }

template <typename T>
__global__ void __launch_bounds__(K0_THREADS)
k0_process_kernel(const ProcessArgs a) {
    __shared__ float s_int[K0_MAX_RAW];
    __shared__ uint8_t s_keep[K0_MAX_RAW];
    __shared__ int s_red[K0_THREADS / 32][3];
    __shared__ int s_count, s_first, s_last, s_base;
    __shared__ float s_max;
    __shared__ double s_sq[128];
    __shared__ float s_nrm;
    const int sp = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = a.off[sp];
    const int n = (int)(a.off[sp + 1] - b);
    const solo_process_params &P = a.p;
    const T *mz = reinterpret_cast<const T *>(a.mz) + b;
    const float *inten = a.inten + b;
    const int stride = P.max_peaks;
    if (tid == 0) {
        a.out_cnt[sp] = 0;
        a.out_valid[sp] = 0;
    }
    if (n > K0_MAX_RAW) {
        if (tid == 0) atomicAdd(a.err, 1);
        return;
    }
    const T lo = (T)P.min_mz, hi = (T)P.max_mz, min_range = (T)P.min_mz_range;
    const bool do_round = P.resolution >= 0;
    T rf = (T)1;
    if (do_round) {
        double f = 1.0;
        for (int d = 0; d < P.resolution; ++d) f *= 10.0;   // exact for d <= 22
        rf = (T)f;
    }
    // m/z of raw peak i as every step after set_mz_range sees it
    auto MZ = [&](int i) -> T { return do_round ? round_decimals(mz[i], rf) : mz[i]; };

    // count / first / last of the peaks currently kept; every thread returns the same verdict
    auto valid_now = [&]() -> bool {
        int cnt = 0, first = n, last = -1;
        for (int i = tid; i < n; i += K0_THREADS)
            if (s_keep[i]) {
                ++cnt;
                first = min(first, i);
                last = max(last, i);
            }
        for (int o = 16; o; o >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        }
        __syncthreads();
        if (lane == 0) {
            s_red[warp][0] = cnt;
            s_red[warp][1] = first;
            s_red[warp][2] = last;
        }
        __syncthreads();
        if (tid == 0) {
            int c = 0, f = n, l = -1;
            for (int w = 0; w < K0_THREADS / 32; ++w) {
                c += s_red[w][0];
                f = min(f, s_red[w][1]);
                l = max(l, s_red[w][2]);
            }
            s_count = c;
            s_first = f;
            s_last = l;
        }
        __syncthreads();
        if (s_count < P.min_peaks || s_count == 0) return false;
        return (T)(MZ(s_last) - MZ(s_first)) >= min_range;   // spectrum.py:14-36, in the array's precision
    };

    // set_mz_range (spectrum.py:79)
    for (int i = tid; i < n; i += K0_THREADS) {
        s_int[i] = inten[i];
        s_keep[i] = mz[i] >= lo && mz[i] <= hi;
    }
    __syncthreads();
    if (!valid_now()) return;
    // round(resolution, 'sum') (spectrum.py:84-89): the kept peaks are the contiguous index range
    // [s_first, s_last] (m/z ascending); the thread of a group's first peak merges the group
    if (do_round) {
        const int f0 = s_first, l0 = s_last;
        __syncthreads();
        for (int i = f0 + tid; i <= l0; i += K0_THREADS) {
            const T r = MZ(i);
            if (i > f0 && MZ(i - 1) == r) continue;
            float sum = s_int[i], top = s_int[i];   // np.add.at on zeros: 0 + x is x
            int best = i, e = i + 1;
            while (e <= l0 && MZ(e) == r) {
                sum = __fadd_rn(sum, s_int[e]);
                if (s_int[e] > top) {
                    top = s_int[e];
                    best = e;
                }
                ++e;
            }
            if (e > i + 1) {
                for (int j = i; j < e; ++j) s_keep[j] = 0;
                s_keep[best] = 1;
                s_int[best] = sum;
            }
        }
        __syncthreads();
        if (!valid_now()) return;
    }
    // remove_precursor_peak(tol, 'Da', isotope = 2) (spectrum.py:90-96)
    if (P.remove_precursor) {
        const int z = a.prec_charge[sp];
        const double neutral = __dmul_rn(__dsub_rn(a.prec_mz[sp], 1.0072766), (double)z);
        const T tol = (T)P.remove_precursor_tolerance;
        for (int c = z; c >= 1; --c)
            for (int iso = 0; iso < 3; ++iso) {
                const T target = (T)__dadd_rn(__ddiv_rn(__dadd_rn(neutral, (double)iso), (double)c), 1.0072766);
                for (int i = tid; i < n; i += K0_THREADS) {
                    T d = MZ(i) - target;
                    if (d < 0) d = -d;
                    if (s_keep[i] && d <= tol) s_keep[i] = 0;
                }
            }
        __syncthreads();
        if (!valid_now()) return;
    }
    // filter_intensity(min_intensity, max_peaks) (spectrum.py:97-99)
    {
        float mx = 0.f;
        for (int i = tid; i < n; i += K0_THREADS)
            if (s_keep[i]) mx = fmaxf(mx, s_int[i]);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_red[warp][0] = __float_as_int(mx);
        __syncthreads();
        if (tid == 0) {
            float m = 0.f;
            for (int w = 0; w < K0_THREADS / 32; ++w) m = fmaxf(m, __int_as_float(s_red[w][0]));
            s_max = m;
        }
        __syncthreads();
        const float thr = __fmul_rn((float)P.min_intensity, s_max);
        for (int i = tid; i < n; i += K0_THREADS)
            if (s_keep[i] && s_int[i] <= thr) s_keep[i] = 0;
        __syncthreads();
    }
    // rank under (intensity, index): number of kept peaks that are greater; keep the max_peaks greatest.
    // Survivors are exactly the peaks with fewer than max_peaks greater ones, and every peak greater than a
    // survivor survives too, so a survivor's count is also its rank among the survivors: it is parked in
    // the flag byte (1 + rank <= 128) for the scaling step. Flags only move between non-zero values during
    // the pass (255 = dropped, cleared after the barrier), so concurrent readers still see "candidate".
    for (int i = tid; i < n; i += K0_THREADS) {
        if (!s_keep[i]) continue;
        const float v = s_int[i];
        int greater = 0;
        for (int j = 0; j < n; ++j)
            greater += (s_keep[j] != 0) && (s_int[j] > v || (s_int[j] == v && j > i));
        s_keep[i] = greater >= P.max_peaks ? 255 : (uint8_t)(1 + greater);
    }
    __syncthreads();
    for (int i = tid; i < n; i += K0_THREADS)
        if (s_keep[i] == 255) s_keep[i] = 0;
    __syncthreads();
    if (!valid_now()) return;
    const int kept = s_count;   // <= max_peaks
    // ordered compaction (m/z ascending = index ascending) + intensity scaling (spectrum.py:104-110)
    if (tid == 0) s_base = 0;
    __syncthreads();
    T *o_mz = reinterpret_cast<T *>(a.out_mz) + (int64_t)sp * stride;
    float *o_int = a.out_int + (int64_t)sp * stride;
    int32_t *o_idx = a.out_idx + (int64_t)sp * stride;
    for (int i0 = 0; i0 < n; i0 += K0_THREADS) {
        const int i = i0 + tid;
        const bool k = i < n && s_keep[i];
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_red[warp][0] = __popc(m);
        __syncthreads();
        int pos = s_base + __popc(m & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += s_red[w][0];
        if (k) {
            float v = s_int[i];
            if (P.scaling == SOLO_SCALING_ROOT) v = __fsqrt_rn(v);
            else if (P.scaling == SOLO_SCALING_RANK) v = (float)(P.max_peaks - ((int)s_keep[i] - 1));
            o_mz[pos] = MZ(i);
            o_int[pos] = v;
            o_idx[pos] = i;
            s_sq[pos] = __dmul_rn((double)v, (double)v);
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < K0_THREADS / 32; ++w) t += s_red[w][0];
            s_base += t;
        }
        __syncthreads();
    }
    // L2 norm (spectrum.py:39-54, :112)
    if (tid == 0) s_nrm = (float)sqrt(numpy_pairwise_sum(s_sq, kept));
    __syncthreads();
    for (int p = tid; p < kept; p += K0_THREADS) o_int[p] = __fdiv_rn(o_int[p], s_nrm);
    if (tid == 0) {
        a.out_cnt[sp] = kept;
        a.out_valid[sp] = 1;
    }
}

void launch_process(solo_handle *h, const ProcessArgs &a, int mz_is_f64) {
    if (a.n <= 0) return;
    if (mz_is_f64) k0_process_kernel<double><<<a.n, K0_THREADS, 0, h->stream>>>(a);
    else k0_process_kernel<float><<<a.n, K0_THREADS, 0, h->stream>>>(a);
    SOLO_CUDA(cudaGetLastError());
    h->launches++;
}

}  // namespace solo

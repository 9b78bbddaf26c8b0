// K1 — feature-hashed spectrum vectorisation, batched over CSR peak arrays.
//
// Replaces reference src/ann_solo/spectrum.py:166-214 (spectrum_to_vector), :147-163 (hash_idx)
// and :123-143 (get_dim). One warp per spectrum; the hashed vector is accumulated in shared
// memory in peak order (collisions add, exactly as `vector[h] += intensity` does), normalised
// and streamed out with coalesced stores. HBM-bound: 8*P + 8 bytes in, 4*hash_len (+2*hash_len
// for the fp16 copy) bytes out per spectrum.
#include "solo_common.cuh"

namespace solo {

constexpr int K1_WARPS = 8;

// NumPy's floating floor-divide (npy_divmod) as evaluated by `(mz - min_bound) // bin_size`
// at spectrum.py:207, in the precision T of the m/z array.
template <typename T>
__device__ __forceinline__ T npy_floor_divide(T a, T b) {
    if (b == T(0)) return a / b;
    T mod = fmod(a, b);  // exact in CUDA for float and double
    T div = (a - mod) / b;
    if (mod != T(0)) {
        if ((b < T(0)) != (mod < T(0))) div -= T(1);
    }
    T fd;
    if (div != T(0)) {
        fd = floor(div);
        if (div - fd > T(0.5)) fd += T(1);
    } else {
        fd = copysign(T(0), a / b);
    }
    return fd;
}

// MurmurHash3_x86_32 of the decimal string of `bin`, seed 42 (spectrum.py:163). Only reached for
// bins outside the precomputed LUT (m/z outside [min_mz, max_mz]); the hot path is the LUT.
__device__ uint32_t murmur3_decimal(long long bin, uint32_t seed) {
    char buf[24];
    int len = 0;
    unsigned long long v = bin < 0 ? (unsigned long long)(-(bin + 1)) + 1ull : (unsigned long long)bin;
    char tmp[24];
    int t = 0;
    do {
        tmp[t++] = (char)('0' + (v % 10ull));
        v /= 10ull;
    } while (v);
    if (bin < 0) buf[len++] = '-';
    while (t) buf[len++] = tmp[--t];
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t h = seed;
    int nblocks = len / 4;
    for (int i = 0; i < nblocks; ++i) {
        uint32_t k = (uint32_t)(uint8_t)buf[4 * i] | ((uint32_t)(uint8_t)buf[4 * i + 1] << 8) |
                     ((uint32_t)(uint8_t)buf[4 * i + 2] << 16) | ((uint32_t)(uint8_t)buf[4 * i + 3] << 24);
        k *= c1;
        k = (k << 15) | (k >> 17);
        k *= c2;
        h ^= k;
        h = (h << 13) | (h >> 19);
        h = h * 5u + 0xe6546b64u;
    }
    uint32_t k = 0;
    int rem = len & 3;
    if (rem == 3) k ^= (uint32_t)(uint8_t)buf[4 * nblocks + 2] << 16;
    if (rem >= 2) k ^= (uint32_t)(uint8_t)buf[4 * nblocks + 1] << 8;
    if (rem >= 1) {
        k ^= (uint32_t)(uint8_t)buf[4 * nblocks];
        k *= c1;
        k = (k << 15) | (k >> 17);
        k *= c2;
        h ^= k;
    }
    h ^= (uint32_t)len;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

template <typename T>
__global__ void __launch_bounds__(K1_WARPS * 32)
k1_vectorize_kernel(const T *__restrict__ mz, const float *__restrict__ inten, const int64_t *__restrict__ off,
                    int64_t n, T min_bound, T bin_size, const uint16_t *__restrict__ lut, int64_t n_lut,
                    int hash_len, int norm, float *__restrict__ out, __half *__restrict__ out_h, float h_scale) {
    extern __shared__ float s_rows[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * K1_WARPS + warp;
    if (s >= n) return;  // warp-uniform; only __syncwarp below
    float *row = s_rows + (size_t)warp * hash_len;
    for (int j = lane; j < hash_len; j += 32) row[j] = 0.f;
    __syncwarp();
    const int64_t beg = off[s], end = off[s + 1];
    for (int64_t base = beg; base < end; base += 32) {
        const int64_t p = base + lane;
        int slot = -1;
        float val = 0.f;
        if (p < end) {
            T m = mz[p];
            T fd = npy_floor_divide<T>(m - min_bound, bin_size);
            long long bin = (long long)floor(fd);  // math.floor(...)
            if (bin >= 0 && bin < n_lut) slot = lut[bin];
            else slot = (int)(murmur3_decimal(bin, 42u) % (uint32_t)hash_len);
            val = inten[p];
        }
        const int cnt = (int)min((int64_t)32, end - base);
        // peak order matters for fp32 collisions: same slot -> same lane -> program order.
        for (int j = 0; j < cnt; ++j) {
            int sj = __shfl_sync(0xffffffffu, slot, j);
            float vj = __shfl_sync(0xffffffffu, val, j);
            if ((sj & 31) == lane) row[sj] += vj;
        }
        __syncwarp();
    }
    float inv_is_div = 1.f;
    if (norm) {
        // ||v||: lane-strided partial sums of squares in double, xor-butterfly, sqrt in double,
        // rounded to float — the summation order the oracle defines (oracle/solo_oracle.cpp).
        double ss = 0.0;
        for (int j = lane; j < hash_len; j += 32) {
            double v = (double)row[j];
            ss = __dadd_rn(ss, __dmul_rn(v, v));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss = __dadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, o));
        inv_is_div = (float)sqrt(ss);
    }
    float *o = out ? out + s * (int64_t)hash_len : nullptr;
    __half *oh = out_h ? out_h + s * (int64_t)hash_len : nullptr;
    for (int j = lane; j < hash_len; j += 32) {
        float v = row[j];
        if (norm) v = __fdiv_rn(v, inv_is_div);  // 0/0 -> NaN like NumPy
        if (o) o[j] = v;
        if (oh) oh[j] = __float2half_rn(v * h_scale);
    }
}

void launch_vectorize(solo_handle *h, const void *d_mz, int mz_is_f64, const float *d_int, const int64_t *d_off,
                      int64_t n, int64_t n_peaks, int norm, float *d_out, __half *d_out_h, int scale_log2) {
    if (n <= 0) return;
    SOLO_REQUIRE(h->lut.p != nullptr, SOLO_ESTATE, "vectorizer not configured");
    size_t smem = (size_t)K1_WARPS * h->hash_len * sizeof(float);
    int grid = div_up(n, K1_WARPS);
    float h_scale = ldexpf(1.f, scale_log2);
    double bytes = 8.0 * (double)n_peaks + 8.0 * n + (d_out ? 4.0 : 0.0) * n * h->hash_len +
                   (d_out_h ? 2.0 : 0.0) * n * h->hash_len;
    StageTimer t(h, ST_VECTORIZE, 1, bytes);
    if (mz_is_f64) {
        auto k = k1_vectorize_kernel<double>;
        SOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, K1_WARPS * 32, smem, h->stream>>>((const double *)d_mz, d_int, d_off, n, h->min_bound,
                                                    h->bin_size, h->lut.as<uint16_t>(), h->n_bins + 2,
                                                    h->hash_len, norm, d_out, d_out_h, h_scale);
    } else {
        auto k = k1_vectorize_kernel<float>;
        SOLO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, K1_WARPS * 32, smem, h->stream>>>((const float *)d_mz, d_int, d_off, n, (float)h->min_bound,
                                                    (float)h->bin_size, h->lut.as<uint16_t>(), h->n_bins + 2,
                                                    h->hash_len, norm, d_out, d_out_h, h_scale);
    }
    SOLO_CUDA(cudaGetLastError());
}

}  // namespace solo

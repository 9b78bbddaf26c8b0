// Pipeline model of the K3 list scan: which producer / barrier discipline keeps the tensor pipe fed?
// One CTA per SM. Shared memory like K3: a resident list chunk (13 k-blocks x 96 rows x 128 B, junk), a ring of
// 16 KB query stages. Tiles of 13 k-blocks; the MMA warp issues 4 x (M128, N96, K16) per k-block into one of two
// accumulators, an "epilogue" thread hands the accumulator back after `epi_delay` cycles.
//   producer 0: 128 threads, 8 cp.async(16 B, zfill) per k-block + cp.async.mbarrier.arrive.noinc  (K3 today)
//   producer 1: NP threads, chunks of the NEXT batch prefetched into registers (predicated ld.global.nc.v4),
//               st.shared.v4 when the batch slot frees, fence.proxy.async, one arrival per warp
//   producer 2: no data at all (barrier traffic only)
//   batch: k-blocks per full/empty barrier pair (the unit of hand-over); ring = 4 stages = 4 / batch slots
//   mma 0: no MMA issued (barriers + commits only); 1: MMAs issued
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo ubench_k3.cu -o ubench_k3
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t bytes) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

constexpr int NKB = 13, NB = 96, A_BYTES = 16384, STAGES = 4;
struct Bars {
    unsigned long long full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base, pad;
};

template <int BATCH, int NP, int D>
__global__ void __launch_bounds__(64 + 64 + NP, 1)
k3_model(const unsigned char *q, const int *rows, const uint32_t *masks, int nrows_total, int tiles, int producer, int do_mma, int epi_delay, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    unsigned char *sA = sm, *sB = sm + STAGES * A_BYTES;
    __shared__ Bars bars;
    constexpr int SLOTS = STAGES / BATCH;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int n_arrive = producer == 0 ? 128 : (producer == 1 ? NP / 32 : 1);
    for (int i = t; i < (STAGES * A_BYTES + NKB * NB * 128) / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (t == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&bars.full[i]), n_arrive); mbar_init(smem_u32(&bars.empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars.tmem_full[i]), 1); mbar_init(smem_u32(&bars.tmem_empty[i]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars.tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars.tmem_base;
    const long long t0 = clock64();
    // batches of a tile: k-blocks [kb0, kb0 + n) with n = BATCH except the last one
    if (warp == 0) {
        // ---- "epilogue": hand the accumulator back after epi_delay cycles
        if (lane == 0) {
            for (int tile = 0; tile < tiles; ++tile) {
                const int buf = tile & 1;
                mbar_wait(smem_u32(&bars.tmem_full[buf]), (tile >> 1) & 1);
                const long long s = clock64();
                while (clock64() - s < epi_delay) {}
                mbar_arrive(smem_u32(&bars.tmem_empty[buf]));
            }
            out[blockIdx.x * 4 + 0] = clock64() - t0;
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        const uint32_t idesc = make_idesc_f16(128, NB);
        uint32_t g = 0;  // running batch counter
        long long waited = 0;
        for (int tile = 0; tile < tiles; ++tile) {
            const int buf = tile & 1;
            mbar_wait(smem_u32(&bars.tmem_empty[buf]), ((tile >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem + (uint32_t)(buf * 256);
#pragma unroll
            for (int kb0 = 0; kb0 < NKB; kb0 += BATCH) {
                const int n = kb0 + BATCH <= NKB ? BATCH : NKB - kb0;
                const uint32_t slot = g % SLOTS, ph = (g / SLOTS) & 1u;
                const long long w0 = clock64();
                mbar_wait(smem_u32(&bars.full[slot]), ph);
                waited += clock64() - w0;
                tc_fence_after();
                if (elect_one()) {
                    if (do_mma) {
#pragma unroll
                        for (int j = 0; j < BATCH; ++j) {
                            if (j < n) {
                                const int kb = kb0 + j;
                                uint64_t adesc = desc_hi | (uint64_t)(((a_base + (slot * BATCH + j) * A_BYTES) >> 4) & 0x3FFFu);
                                uint64_t bdesc = desc_hi | (uint64_t)(((b_base + (uint32_t)kb * NB * 128) >> 4) & 0x3FFFu);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    tc_mma_f16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                                    adesc += 2;
                                    bdesc += 2;
                                }
                            }
                        }
                    }
                    tc_commit(smem_u32(&bars.empty[slot]));
                    if (kb0 + n == NKB) tc_commit(smem_u32(&bars.tmem_full[buf]));
                }
                __syncwarp();
                ++g;
            }
        }
        if (lane == 0) { out[blockIdx.x * 4 + 1] = clock64() - t0; out[blockIdx.x * 4 + 2] = waited; }
    } else if (warp >= 4) {
        // ---- producers
        const int p = t - 128;
        uint32_t g = 0;
        if (producer == 0 && p < 128) {
            // K3 today: per k-block barrier semantics emulated at batch granularity: BATCH must be 1 for a faithful copy
            const int chunk = p & 7, rbase = p >> 3;
            const uint32_t dst_off = (uint32_t)(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));
            for (int tile = 0; tile < tiles; ++tile) {
                const int *r = rows + (((size_t)blockIdx.x * tiles + tile) * 128) % nrows_total;
                const unsigned char *src[8];
                uint32_t nz[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int qq = r[rbase + 16 * i];
                    src[i] = q + (size_t)qq * 1664 + chunk * 16;
                    nz[i] = masks[(size_t)qq * 8 + chunk];
                }
#pragma unroll
                for (int kb0 = 0; kb0 < NKB; kb0 += BATCH) {
                    const int n = kb0 + BATCH <= NKB ? BATCH : NKB - kb0;
                    const uint32_t slot = g % SLOTS, ph = (g / SLOTS) & 1u;
                    mbar_wait(smem_u32(&bars.empty[slot]), ph ^ 1u);
#pragma unroll
                    for (int j = 0; j < BATCH; ++j)
                        if (j < n) {
                            const uint32_t dst0 = smem_u32(sA) + (slot * BATCH + j) * A_BYTES + dst_off;
#pragma unroll
                            for (int i = 0; i < 8; ++i) cp_async16_zfill(dst0 + i * 16 * 128, src[i] + (kb0 + j) * 128, (nz[i] >> (kb0 + j)) & 1u ? 16u : 0u);
                        }
                    cp_async_arrive_noinc(smem_u32(&bars.full[slot]));
                    ++g;
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else if (producer == 1 && p < NP) {
            // register prefetch: thread p owns chunk (p & 7) of rows (p >> 3) + (NP/8) * i, i < 1024 / NP
            constexpr int NR = 1024 / NP, RSTEP = NP / 8;
            const int chunk = p & 7, rbase = p >> 3;
            const uint32_t dst_off = (uint32_t)(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));
            uint4 v[D][BATCH][NR];
            const unsigned char *src[NR];
            uint32_t nz[NR];
            constexpr int NBT = (NKB + BATCH - 1) / BATCH;
            const int total = tiles * NBT;
            auto fetch_g = [&](int gg, uint4 (&vv)[BATCH][NR]) {
                if (gg >= total) return;
                const int tile = gg / NBT, kb0 = (gg % NBT) * BATCH;
                if (kb0 == 0) {
                    const int *r = rows + (((size_t)blockIdx.x * tiles + tile) * 128) % nrows_total;
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const int qq = r[rbase + RSTEP * i];
                        src[i] = q + (size_t)qq * 1664 + chunk * 16;
                        nz[i] = masks[(size_t)qq * 8 + chunk];
                    }
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j)
                    if (kb0 + j < NKB) {
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            vv[j][i] = make_uint4(0, 0, 0, 0);
                            if ((nz[i] >> (kb0 + j)) & 1u) vv[j][i] = __ldg(reinterpret_cast<const uint4 *>(src[i] + (kb0 + j) * 128));
                        }
                    }
            };
#pragma unroll
            for (int d = 0; d < D - 1; ++d) fetch_g(d, v[d]);
            for (int g0 = 0; g0 < total; g0 += D) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const int gg = g0 + d;
                    if (gg < total) {
                        fetch_g(gg + D - 1, v[(d + D - 1) % D]);
                        const int kb0 = (gg % NBT) * BATCH;
                        const uint32_t slot = gg % SLOTS, ph = (gg / SLOTS) & 1u;
                        mbar_wait(smem_u32(&bars.empty[slot]), ph ^ 1u);
#pragma unroll
                        for (int j = 0; j < BATCH; ++j)
                            if (kb0 + j < NKB) {
                                unsigned char *dst0 = sA + (slot * BATCH + j) * A_BYTES + dst_off;
#pragma unroll
                                for (int i = 0; i < NR; ++i) *reinterpret_cast<uint4 *>(dst0 + i * RSTEP * 128) = v[d][j][i];
                            }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars.full[slot]));
                    }
                }
            }
        } else if (producer == 2 && p < 32) {
            for (int tile = 0; tile < tiles; ++tile)
                for (int kb0 = 0; kb0 < NKB; kb0 += BATCH) {
                    const uint32_t slot = g % SLOTS, ph = (g / SLOTS) & 1u;
                    mbar_wait(smem_u32(&bars.empty[slot]), ph ^ 1u);
                    if (lane == 0) mbar_arrive(smem_u32(&bars.full[slot]));
                    __syncwarp();
                    ++g;
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

// ---- swapped operands: the list chunk (128 rows) is the A operand, resident in TENSOR MEMORY (columns 0..415, junk
// here), the gathered query rows are the B operand (N rows x 128 B per k-block stage, deep ring: no resident chunk in
// shared memory). NACC accumulators of N columns behind the A columns.
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
struct BarsS {
    unsigned long long full[16], empty[16], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base, pad;
};
template <int N, int NACC, int NSTAGES, int BATCH>
__global__ void __launch_bounds__(256, 1)
k3_swapped(const unsigned char *q, const int *rows, const uint32_t *masks, int nrows_total, int tiles, int producer, int do_mma, int epi_delay, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    constexpr int ST_BYTES = N * 128;
    constexpr int SLOTS = NSTAGES / BATCH;
    __shared__ BarsS bars;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < NSTAGES * ST_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (t == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(smem_u32(&bars.full[i]), producer == 2 ? 1 : 128); mbar_init(smem_u32(&bars.empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars.tmem_full[i]), 1); mbar_init(smem_u32(&bars.tmem_empty[i]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars.tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars.tmem_base;
    const long long t0 = clock64();
    if (warp == 0) {
        if (lane == 0) {
            for (int tile = 0; tile < tiles; ++tile) {
                const int buf = tile % NACC;
                mbar_wait(smem_u32(&bars.tmem_full[buf]), (tile / NACC) & 1);
                const long long s = clock64();
                while (clock64() - s < epi_delay) {}
                mbar_arrive(smem_u32(&bars.tmem_empty[buf]));
            }
        }
    } else if (warp == 1) {
        const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t b_base = smem_u32(sm);
        const uint32_t idesc = make_idesc_f16(128, N);
        uint32_t g = 0;
        long long waited = 0, waited_acc = 0;
        for (int tile = 0; tile < tiles; ++tile) {
            const int buf = tile % NACC;
            const long long w1 = clock64();
            mbar_wait(smem_u32(&bars.tmem_empty[buf]), ((tile / NACC) & 1) ^ 1);
            waited_acc += clock64() - w1;
            tc_fence_after();
            const uint32_t tmem_d = tmem + 416u + (uint32_t)(buf * N);
#pragma unroll
            for (int kb0 = 0; kb0 < NKB; kb0 += BATCH) {
                const int n = kb0 + BATCH <= NKB ? BATCH : NKB - kb0;
                const uint32_t slot = g % SLOTS, ph = (g / SLOTS) & 1u;
                const long long w0 = clock64();
                mbar_wait(smem_u32(&bars.full[slot]), ph);
                waited += clock64() - w0;
                tc_fence_after();
                if (elect_one()) {
                    if (do_mma) {
#pragma unroll
                        for (int j = 0; j < BATCH; ++j) {
                            if (j < n) {
                                const int kb = kb0 + j;
                                uint64_t bdesc = desc_hi | (uint64_t)(((b_base + (slot * BATCH + j) * ST_BYTES) >> 4) & 0x3FFFu);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    tc_mma_f16_ts(tmem_d, tmem + (uint32_t)(kb * 32 + k * 8), bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                                    bdesc += 2;
                                }
                            }
                        }
                    }
                    tc_commit(smem_u32(&bars.empty[slot]));
                    if (kb0 + n == NKB) tc_commit(smem_u32(&bars.tmem_full[buf]));
                }
                __syncwarp();
                ++g;
            }
        }
        if (lane == 0) { out[blockIdx.x * 4 + 1] = clock64() - t0; out[blockIdx.x * 4 + 2] = waited; out[blockIdx.x * 4 + 3] = waited_acc; }
    } else if (warp >= 4) {
        const int p = t - 128;
        uint32_t g = 0;
        constexpr int NR = N / 16;  // rows per thread
        const int chunk = p & 7, rbase = p >> 3;
        const uint32_t dst_off = (uint32_t)(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));
        for (int tile = 0; tile < tiles; ++tile) {
            const int *r = rows + (((size_t)blockIdx.x * tiles + tile) * 128) % nrows_total;
            const unsigned char *src[NR];
            uint32_t nz[NR];
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int qq = r[rbase + 16 * i];
                src[i] = q + (size_t)qq * 1664 + chunk * 16;
                nz[i] = masks[(size_t)qq * 8 + chunk];
            }
#pragma unroll
            for (int kb0 = 0; kb0 < NKB; kb0 += BATCH) {
                const int n = kb0 + BATCH <= NKB ? BATCH : NKB - kb0;
                const uint32_t slot = g % SLOTS, ph = (g / SLOTS) & 1u;
                mbar_wait(smem_u32(&bars.empty[slot]), ph ^ 1u);
                if (producer == 0) {
#pragma unroll
                    for (int j = 0; j < BATCH; ++j)
                        if (j < n) {
                            const uint32_t dst0 = smem_u32(sm) + (slot * BATCH + j) * ST_BYTES + dst_off;
#pragma unroll
                            for (int i = 0; i < NR; ++i) cp_async16_zfill(dst0 + i * 16 * 128, src[i] + (kb0 + j) * 128, (nz[i] >> (kb0 + j)) & 1u ? 16u : 0u);
                        }
                    cp_async_arrive_noinc(smem_u32(&bars.full[slot]));
                } else if (p == 0) {
                    mbar_arrive(smem_u32(&bars.full[slot]));
                }
                ++g;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

int main() {
    CK(cudaSetDevice(0));
    const int NQ = 8192;
    unsigned char *d_q;
    CK(cudaMalloc(&d_q, (size_t)NQ * 1664));
    CK(cudaMemset(d_q, 0, (size_t)NQ * 1664));
    const int NR = 1 << 20;
    std::vector<int> hr(NR);
    uint32_t x = 12345;
    for (int i = 0; i < NR; ++i) { x = x * 1664525u + 1013904223u; hr[i] = (x >> 8) % NQ; }
    int *d_rows;
    CK(cudaMalloc(&d_rows, NR * 4));
    CK(cudaMemcpy(d_rows, hr.data(), NR * 4, cudaMemcpyHostToDevice));
    std::vector<uint32_t> hm((size_t)NQ * 8);
    for (auto &m : hm) {  // each (chunk, k-block) bit set with probability 0.4, like real query vectors
        m = 0;
        for (int b = 0; b < 13; ++b) { x = x * 1664525u + 1013904223u; if ((x >> 8) % 10 < 4) m |= 1u << b; }
    }
    uint32_t *d_m;
    CK(cudaMalloc(&d_m, hm.size() * 4));
    CK(cudaMemcpy(d_m, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice));
    long long *d_out;
    CK(cudaMalloc(&d_out, 148 * 4 * 8));
    const int smem = STAGES * A_BYTES + NKB * NB * 128 + 2048;
    const int tiles = 300;
    auto run = [&](auto kern, int nthreads, const char *name, int producer, int do_mma, int epi) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaMemset(d_out, 0, 148 * 4 * 8));
        kern<<<148, nthreads, smem>>>(d_q, d_rows, d_m, NR, tiles, producer, do_mma, epi, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
        long long h[4];
        CK(cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost));
        printf("%-44s producer %d mma %d epi %4d : %7.1f cycles per k-block (MMA warp waits %5.1f on full)\n", name, producer, do_mma, epi,
               (double)h[1] / (tiles * NKB), (double)h[2] / (tiles * NKB));
        fflush(stdout);
    };
    auto runs = [&](auto kern, int smem_b, const char *name, int producer, int do_mma, int epi) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
        CK(cudaMemset(d_out, 0, 148 * 4 * 8));
        kern<<<148, 256, smem_b>>>(d_q, d_rows, d_m, NR, tiles, producer, do_mma, epi, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
        long long h[4];
        CK(cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost));
        printf("%-50s producer %d mma %d epi %4d : %7.1f cycles per k-block (MMA warp waits %5.1f on full, %5.1f on the accumulator)\n", name, producer, do_mma, epi,
               (double)h[1] / (tiles * NKB), (double)h[2] / (tiles * NKB), (double)h[3] / (tiles * NKB));
        fflush(stdout);
    };
    for (int epi : {0, 400}) {
        for (int prod : {2, 0}) {
            for (int mma : {0, 1}) {
                if (prod == 2 && mma == 0) continue;
                runs(k3_swapped<96, 1, 16, 1>, 16 * 96 * 128 + 2048, "swapped N=96, 1 acc, 16 stages, batch 1", prod, mma, epi);
                runs(k3_swapped<96, 1, 16, 2>, 16 * 96 * 128 + 2048, "swapped N=96, 1 acc, 16 stages, batch 2", prod, mma, epi);
                runs(k3_swapped<96, 1, 8, 2>, 8 * 96 * 128 + 2048, "swapped N=96, 1 acc, 8 stages, batch 2", prod, mma, epi);
                runs(k3_swapped<96, 1, 16, 4>, 16 * 96 * 128 + 2048, "swapped N=96, 1 acc, 16 stages, batch 4", prod, mma, epi);
                runs(k3_swapped<48, 2, 16, 2>, 16 * 48 * 128 + 2048, "swapped N=48, 2 acc, 16 stages, batch 2", prod, mma, epi);
                runs(k3_swapped<48, 2, 16, 4>, 16 * 48 * 128 + 2048, "swapped N=48, 2 acc, 16 stages, batch 4", prod, mma, epi);
                runs(k3_swapped<32, 2, 16, 4>, 16 * 32 * 128 + 2048, "swapped N=32, 2 acc, 16 stages, batch 4", prod, mma, epi);
            }
        }
    }
    return 0;
}

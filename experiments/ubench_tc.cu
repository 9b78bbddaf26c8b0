// Micro-benchmarks that decide the K3 (list scan) design: issue cost of tcgen05.mma / tcgen05.commit patterns,
// and the rate of the three ways to gather 128 query rows x 128 B into a swizzled stage
// (cp.async 16 B, TMA tile::gather4). One CTA per SM, junk operands, clock64 around the measured loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo ubench_tc.cu -o ubench_tc -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// gives up after ~1e9 cycles (a wrong tensor map must not hang the box)
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 1000000000ll) return false;
    return true;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t bytes) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

struct Bars { unsigned long long done[8]; unsigned long long full[16]; uint32_t tmem_base; uint32_t pad; };

// ---- test 1: MMA issue patterns. One elected thread issues groups of PER MMAs (M=128, N=n, fully unrolled,
// descriptors = base + immediate) and one commit per group. NACC independent TMEM accumulators are cycled per MMA.
// wait_mode 1: before each group the issuer waits for the commit of the group `lag` groups earlier.
template <int PER, int NACC>
__global__ void __launch_bounds__(128, 1) mma_issue_kernel(int n, int groups, int wait_mode, int lag, int do_commit, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ Bars bars;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (64 * 1024 + 256 * 128 * 2) / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars.done[i]), 1);
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
    if (threadIdx.x == 32) {
        const uint64_t a0 = make_desc_sw128(smem_u32(sm)), b0 = make_desc_sw128(smem_u32(sm + 64 * 1024));
        const uint32_t idesc = make_idesc_f16(128, n);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            if (wait_mode == 1 && g >= lag) {
                const int w = g - lag;
                mbar_wait(smem_u32(&bars.done[w & 7]), (w >> 3) & 1);
                tc_fence_after();
            }
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                // A: 4 stages of 16 KB, 4 K-steps of 32 B inside each; B: one 32 KB tile, 4 K-steps
                const uint64_t ad = a0 + (uint64_t)(((i >> 2) & 3) * (16384 >> 4) + (i & 3) * 2);
                const uint64_t bd = b0 + (uint64_t)((i & 3) * 2);
                tc_mma_f16(tmem + (uint32_t)((i % NACC) * (512 / NACC)), ad, bd, idesc, i >= NACC ? 1u : 0u);
            }
            if (do_commit) tc_commit(smem_u32(&bars.done[g & 7]));
        }
        if (!do_commit) tc_commit(smem_u32(&bars.done[(groups - 1) & 7]));
        const long long t1 = clock64();
        const int w = groups - 1;
        mbar_wait(smem_u32(&bars.done[w & 7]), do_commit ? (w >> 3) & 1 : 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

// ---- test 2: gather rates. `stages` stages of 128 rows x 128 B; mode 0: 128 threads x 8 cp.async 16 B +
// arrive.noinc (like K3 today); mode 1: one thread, 32 TMA gather4 per stage; mode 2: mode 0 with `nthr` of the
// 128 threads doing all copies (nthr = 32: one warp, 32 copies per lane). A consumer thread waits for each
// stage and frees it right away (plain arrive on an empty barrier), so this is the producer-side rate.
struct GBars { unsigned long long full[16]; unsigned long long empty[16]; };
__global__ void __launch_bounds__(544, 1) gather_kernel(const __grid_constant__ CUtensorMap map, const __half *q, const int *rows, int nrows_total, int iters, int stages, int mode, int nthr, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ GBars bars;
    const int prod_threads = mode == 1 ? 1 : nthr;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) {
            mbar_init(smem_u32(&bars.full[i]), mode == 1 ? 1 : prod_threads);
            mbar_init(smem_u32(&bars.empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int t = threadIdx.x;
    const long long t0 = clock64();
    if (t < 512) {
        if (mode == 1) {
            if (t == 0) {
                uint32_t stage = 0, phase = 0;
                for (int it = 0; it < iters; ++it) {
                    const int *r = rows + (((size_t)blockIdx.x * iters + it) * 128) % nrows_total;
                    const int kb = it % 12;
                    if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                    mbar_expect_tx(smem_u32(&bars.full[stage]), 16384);
                    const uint32_t dst = smem_u32(sm) + stage * 16384;
#pragma unroll 4
                    for (int j = 0; j < 32; ++j)
                        tma_gather4(dst + j * 512, &map, kb * 64, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3], smem_u32(&bars.full[stage]));
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (t < prod_threads) {
            const int per = 1024 / prod_threads;  // 16-byte copies per thread per stage
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < iters; ++it) {
                const int *r = rows + (((size_t)blockIdx.x * iters + it) * 128) % nrows_total;
                const int kb = it % 12;
                if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                const uint32_t dst = smem_u32(sm) + stage * 16384;
#pragma unroll 8
                for (int j = 0; j < per; ++j) {
                    const int c = j * prod_threads + t;  // chunk id 0..1023: row = c >> 3, chunk-in-row = c & 7
                    const int row = c >> 3, ch = c & 7;
                    cp_async16(dst + row * 128 + ((ch ^ (row & 7)) << 4), reinterpret_cast<const unsigned char *>(q) + (size_t)r[row] * 1600 + kb * 128 + ch * 16, 16);
                }
                cp_async_arrive_noinc(smem_u32(&bars.full[stage]));
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
    } else if (t == 512) {
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < iters; ++it) {
            if (!mbar_wait_bounded(smem_u32(&bars.full[stage]), phase)) { if (blockIdx.x == 0) out[2] = -1; break; }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars.empty[stage])) : "memory");
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (t == 0 && blockIdx.x == 0 && mode == 1 && iters == 1) {  // correctness probe of gather4: checksum of the stage
        long long s = 0;
        for (int i = 0; i < 16384 / 2; ++i) s += (long long)reinterpret_cast<const unsigned short *>(sm)[i] * ((i % 977) + 1);
        out[1] = s;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    long long *d_out;
    CK(cudaMalloc(&d_out, 64));
    long long h[2];
    const int smem = 64 * 1024 + 256 * 128 * 2 + 2048;
    printf("== MMA issue patterns (cycles per MMA: issue-loop / incl. drain), 148 CTAs, M=128, K=16 per MMA, unrolled issue\n");
    auto run = [&](auto kern, const char *name, int per, int n, int wm, int lag, int do_commit) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int groups = 4096 / per;
        kern<<<148, 128, smem>>>(n, groups, wm, lag, do_commit, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("%-22s N=%3d wait=%d lag=%d commit=%d : %7.1f / %7.1f cycles per MMA (floor %.0f)\n", name, n, wm, lag, do_commit,
               (double)h[0] / (groups * per), (double)h[1] / (groups * per), 128.0 * n / 256.0);
    };
    for (int n : {96}) {
        run(mma_issue_kernel<16, 1>, "16/group, 1 acc", 16, n, 0, 1, 0);
        run(mma_issue_kernel<16, 1>, "16/group, 1 acc", 16, n, 0, 1, 1);
        if (n <= 256) run(mma_issue_kernel<16, 2>, "16/group, 2 acc", 16, n, 0, 1, 0);
        if (n <= 128) run(mma_issue_kernel<16, 4>, "16/group, 4 acc", 16, n, 0, 1, 0);
        if (n <= 128) run(mma_issue_kernel<16, 4>, "16/group, 4 acc", 16, n, 0, 1, 1);
        run(mma_issue_kernel<8, 1>, "8/group, 1 acc", 8, n, 0, 1, 1);
        run(mma_issue_kernel<8, 2>, "8/group, 2 acc", 8, n, 0, 1, 1);
        run(mma_issue_kernel<8, 1>, "8/group, 1 acc", 8, n, 1, 1, 1);
        run(mma_issue_kernel<8, 1>, "8/group, 1 acc", 8, n, 1, 2, 1);
        run(mma_issue_kernel<4, 1>, "4/group, 1 acc", 4, n, 0, 1, 1);
        run(mma_issue_kernel<4, 1>, "4/group, 1 acc", 4, n, 1, 2, 1);
        run(mma_issue_kernel<32, 1>, "32/group, 1 acc", 32, n, 0, 1, 1);
        run(mma_issue_kernel<32, 2>, "32/group, 2 acc", 32, n, 0, 1, 1);
        run(mma_issue_kernel<32, 1>, "32/group, 1 acc", 32, n, 1, 2, 1);
    }
    // gather
    const int NQ = 8192, DIM = 800;
    __half *d_q;
    CK(cudaMalloc(&d_q, (size_t)NQ * DIM * 2));
    std::vector<__half> hq((size_t)NQ * DIM);
    for (size_t i = 0; i < hq.size(); ++i) hq[i] = __float2half((float)((i * 2654435761u >> 16) & 1023) / 1024.f);
    CK(cudaMemcpy(d_q, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice));
    const int NR = 1 << 20;
    std::vector<int> hr(NR);
    uint32_t x = 12345;
    for (int i = 0; i < NR; ++i) { x = x * 1664525u + 1013904223u; hr[i] = (x >> 8) % NQ; }
    int *d_rows;
    CK(cudaMalloc(&d_rows, NR * 4));
    CK(cudaMemcpy(d_rows, hr.data(), NR * 4, cudaMemcpyHostToDevice));
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
    PFN_encodeTiled enc = (PFN_encodeTiled)fnp;
    const int gsm = 12 * 16384 + 2048;
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
    // reference checksum of a correctly gathered + swizzled stage (block 0, iteration 0, kb 0)
    long long want = 0;
    {
        std::vector<unsigned short> st(8192, 0);
        for (int row = 0; row < 128; ++row)
            for (int ch = 0; ch < 8; ++ch)
                for (int e = 0; e < 8; ++e) {
                    const int dst = (row * 128 + ((ch ^ (row & 7)) << 4)) / 2 + e;
                    st[dst] = reinterpret_cast<unsigned short *>(hq.data())[(size_t)hr[row] * DIM + ch * 8 + e];
                }
        for (int i = 0; i < 8192; ++i) want += (long long)st[i] * ((i % 977) + 1);
    }
    {
        CUtensorMap dummy;
        memset(&dummy, 0, sizeof dummy);
        printf("== cp.async 16 B gather\n");
        for (int nthr : {512, 256, 128})
            for (int stages : {4, 8, 12}) {
                gather_kernel<<<148, 544, gsm>>>(dummy, d_q, d_rows, NR, 2000, stages, 0, nthr, d_out);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
                printf("   cp.async, %3d producer threads, %2d stages: %.1f cycles per 16 KB stage (148 CTAs)\n", nthr, stages, (double)h[0] / 2000);
            }
    }
    for (int boxrows : {1}) {
        CUtensorMap map;
        cuuint64_t gdim[2] = {(cuuint64_t)DIM, (cuuint64_t)NQ};
        cuuint64_t gstr[1] = {(cuuint64_t)DIM * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_q, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== gather4 with tensor-map box rows = %d (encode rc %d)\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        CK(cudaMemset(d_out, 0, 64));
        gather_kernel<<<1, 544, gsm>>>(map, d_q, d_rows, NR, 1, 4, 1, 128, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("   gather4 launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        { long long h3[3]; CK(cudaMemcpy(h3, d_out, 24, cudaMemcpyDeviceToHost)); if (h3[2] == -1) { printf("   gather4 never completed its bytes (timeout)\n"); continue; } }
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("   stage checksum %lld, expected %lld -> %s\n", h[1], want, h[1] == want ? "LAYOUT OK" : "layout differs");
        if (h[1] != want) continue;
        for (int stages : {4, 8, 12}) {
            gather_kernel<<<148, 544, gsm>>>(map, d_q, d_rows, NR, 2000, stages, 1, 128, d_out);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
            printf("   gather4, %2d stages: %.1f cycles per 16 KB stage (148 CTAs)\n", stages, (double)h[0] / 2000);
        }
    }
    return 0;
}

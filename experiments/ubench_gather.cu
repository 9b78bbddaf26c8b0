// Which mechanism moves 128 query rows x 128 B (one k-block of a gathered query tile) from L2 into a
// SWIZZLE_128B stage fastest, and is the limit per SM or chip-wide?  One CTA per SM (grid given), a consumer
// thread frees each stage right away, so the number is the producer-side rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo ubench_gather.cu -o ubench_gather -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    int spins = 0;
    while (!mbar_try(bar, parity))
        if (++spins > 20000000) return false;
    return true;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t bytes) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap *map, int c0, int r0, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

struct GBars { unsigned long long full[16]; unsigned long long empty[16]; };

// modes:
//  0  cp.async 16 B, nthr producer threads, row ids of the tile in shared memory (no dependent global load)
//  1  ld.global.nc.v4 + st.shared.v4 (+ fence.proxy.async + arrive), 128 threads
//  2  cp.async.bulk 128 B per row (no swizzle; rate only), one per thread, 128 threads
//  3  TMA 2D tile, box {64, 1} per row (SWIZZLE_128B), one per thread, 128 threads
//  4  TMA gather4, 32 lanes of one warp issue one each
//  5  TMA 2D tile of 128 consecutive rows (one instruction per stage) — the contiguous-load ceiling
//  6  TMA 2D tile, fully out of bounds (zero fill, no memory traffic), one instruction per stage
//  7  sparse expansion: mode 6 zero fill + `nnz` 2-byte st.shared per row per stage (values from registers)
//  8  cp.async 16 B, only `nnz`/8 of the chunks of each row (others left alone): thread-op scaling
__global__ void __launch_bounds__(544, 1) gather_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map128, const __half *q, const int *rows, int nrows_total, int iters, int stages, int mode, int nthr, int nnz, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ GBars bars;
    __shared__ int srow[2][128];
    const int t = threadIdx.x;
    int n_arrive = nthr;
    if (mode == 2 || mode == 3 || mode == 4 || mode == 5 || mode == 6) n_arrive = 1;
    if (mode == 7) n_arrive = 1 + 128;
    if (t == 0) {
        for (int i = 0; i < 16; ++i) {
            mbar_init(smem_u32(&bars.full[i]), n_arrive);
            mbar_init(smem_u32(&bars.empty[i]), 1);
            mbar_init(smem_u32(reinterpret_cast<unsigned long long *>(sm + 12 * 16384) + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 128) { srow[0][t] = rows[((size_t)blockIdx.x * 128 + t) % nrows_total]; srow[1][t] = rows[((size_t)blockIdx.x * 128 + 4096 + t) % nrows_total]; }
    __syncthreads();
    const long long t0 = clock64();
    if (t < 512) {
        uint32_t stage = 0, phase = 0;
        if (mode == 0 || mode == 8) {
            if (t < nthr) {
                const int per = 1024 / nthr;
                for (int it = 0; it < iters; ++it) {
                    const int *r = srow[(it / 13) & 1];
                    const int kb = it % 12;
                    if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                    const uint32_t dst = smem_u32(sm) + stage * 16384;
#pragma unroll 8
                    for (int j = 0; j < per; ++j) {
                        const int c = j * nthr + t;
                        const int row = c >> 3, ch = c & 7;
                        if (mode == 8 && ch >= nnz) continue;
                        cp_async16(dst + row * 128 + ((ch ^ (row & 7)) << 4), reinterpret_cast<const unsigned char *>(q) + (size_t)r[row] * 1600 + kb * 128 + ch * 16, 16);
                    }
                    cp_async_arrive_noinc(smem_u32(&bars.full[stage]));
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
            }
        } else if (mode == 1) {
            if (t < 128) {
                for (int it = 0; it < iters; ++it) {
                    const int *r = srow[(it / 13) & 1];
                    const int kb = it % 12;
                    uint4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = j * 128 + t;
                        const int row = c >> 3, ch = c & 7;
                        v[j] = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(q) + (size_t)r[row] * 1600 + kb * 128 + ch * 16));
                    }
                    if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                    unsigned char *dst = sm + stage * 16384;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = j * 128 + t;
                        const int row = c >> 3, ch = c & 7;
                        *reinterpret_cast<uint4 *>(dst + row * 128 + ((ch ^ (row & 7)) << 4)) = v[j];
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(smem_u32(&bars.full[stage]));
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (mode == 2 || mode == 3) {
            if (t < 128) {
                for (int it = 0; it < iters; ++it) {
                    const int *r = srow[(it / 13) & 1];
                    const int kb = it % 12;
                    if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                    const uint32_t dst = smem_u32(sm) + stage * 16384 + t * 128;
                    if (t == 0) mbar_expect_tx(smem_u32(&bars.full[stage]), 16384);
                    __syncwarp();
                    // (expect_tx by thread 0 may race with completions of other warps' copies; the barrier's tx-count
                    //  may go negative transiently, which is allowed)
                    if (mode == 2) bulk_g2s(dst, reinterpret_cast<const unsigned char *>(q) + (size_t)r[t] * 1600 + kb * 128, 128, smem_u32(&bars.full[stage]));
                    else tma_2d(dst, &map1, kb * 64, r[t], smem_u32(&bars.full[stage]));
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (mode == 4) {
            if (t < 32) {
                for (int it = 0; it < iters; ++it) {
                    const int *r = srow[(it / 13) & 1];
                    const int kb = it % 12;
                    if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                    if (t == 0) mbar_expect_tx(smem_u32(&bars.full[stage]), 16384);
                    __syncwarp();
                    tma_gather4(smem_u32(sm) + stage * 16384 + t * 512, &map1, kb * 64, r[4 * t], r[4 * t + 1], r[4 * t + 2], r[4 * t + 3], smem_u32(&bars.full[stage]));
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (mode == 5 || mode == 6) {
            if (t == 0) {
                for (int it = 0; it < iters; ++it) {
                    const int kb = it % 12;
                    if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                    mbar_expect_tx(smem_u32(&bars.full[stage]), 16384);
                    const int r0 = mode == 6 ? (1 << 20) : (int)(((size_t)blockIdx.x * 977 + it * 131) % (8192 - 128));
                    tma_2d(smem_u32(sm) + stage * 16384, &map128, kb * 64, r0, smem_u32(&bars.full[stage]));
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (mode == 7) {
            // thread 0 keeps the zero fill two stages ahead on its own barriers (empty -> zero -> scatter -> full)
            // simplified for rate: zero fill and scatter target the same stage, the scatter waits for the zeros via
            // a second barrier array placed in dynamic smem behind the stages
            unsigned long long *zbar = reinterpret_cast<unsigned long long *>(sm + 12 * 16384);
            if (t == 128) {
                for (int it = 0; it < iters; ++it) {
                    if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                    mbar_expect_tx(smem_u32(&zbar[stage]), 16384);
                    tma_2d(smem_u32(sm) + stage * 16384, &map128, 0, 1 << 20, smem_u32(&zbar[stage]));
                    mbar_arrive(smem_u32(&bars.full[stage]));
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
                }
            } else if (t < 128) {
                const unsigned short val = (unsigned short)(0x3c00 + t);
                for (int it = 0; it < iters; ++it) {
                    if (!mbar_wait_bounded(smem_u32(&zbar[stage]), phase)) break;
                    unsigned char *dst = sm + stage * 16384 + t * 128;
                    for (int j = 0; j < nnz; ++j) {
                        const int col = (t * 7 + j * 13 + it) & 63;  // element within the 64-wide k-block
                        const int ch = col >> 3;
                        *reinterpret_cast<unsigned short *>(dst + (((ch ^ (t & 7)) << 4) | ((col & 7) << 1))) = val;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(smem_u32(&bars.full[stage]));
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (t == 512) {
        uint32_t stage = 0, phase = 0;
        int it = 0;
        for (; it < iters; ++it) {
            if (!mbar_wait_bounded(smem_u32(&bars.full[stage]), phase)) { if (blockIdx.x == 0) out[2] = -1 - it; break; }
            mbar_arrive(smem_u32(&bars.empty[stage]));
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = it; }
    }
    __syncthreads();
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    long long *d_out;
    CK(cudaMalloc(&d_out, 64));
    long long h[3];
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
    const int gsm = 12 * 16384 + 2048 + 256;
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
    CUtensorMap map1, map128;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)DIM, (cuuint64_t)NQ};
        cuuint64_t gstr[1] = {(cuuint64_t)DIM * 2};
        cuuint32_t estr[2] = {1, 1};
        cuuint32_t box1[2] = {64, 1}, box128[2] = {64, 128};
        CUresult r1 = enc(&map1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_q, gdim, gstr, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&map128, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_q, gdim, gstr, box128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("tensor maps: %d %d\n", (int)r1, (int)r2);
    }
    auto run = [&](const char *name, int grid, int mode, int nthr, int nnz, int stages) {
        CK(cudaMemset(d_out, 0, 64));
        const int iters = 2000;
        gather_kernel<<<grid, 544, gsm>>>(map1, map128, d_q, d_rows, NR, iters, stages, mode, nthr, nnz, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
        CK(cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost));
        printf("%-52s grid %3d stages %2d : %8.1f cycles per 16 KB stage%s\n", name, grid, stages, (double)h[0] / (h[1] > 0 ? h[1] : 1), h[2] < 0 ? "  (TIMEOUT)" : "");
        fflush(stdout);
    };
    for (int grid : {1, 37, 148}) {
        run("0 cp.async 16 B, 128 thr, ids in smem", grid, 0, 128, 8, 8);
        run("0 cp.async 16 B, 256 thr, ids in smem", grid, 0, 256, 8, 8);
        run("1 ldg.v4 + sts.v4, 128 thr", grid, 1, 128, 8, 8);
        run("2 cp.async.bulk 128 B per row, 128 thr", grid, 2, 128, 8, 8);
        run("3 TMA 2D box {64,1} per row, 128 thr", grid, 3, 128, 8, 8);
        run("4 TMA gather4, 32 lanes", grid, 4, 128, 8, 8);
        run("5 TMA 2D box {64,128} contiguous", grid, 5, 128, 8, 8);
        run("6 TMA 2D out of bounds (zero fill)", grid, 6, 128, 8, 8);
        run("7 zero fill + 4 scattered 2 B stores per row", grid, 7, 128, 4, 8);
        run("7 zero fill + 8 scattered 2 B stores per row", grid, 7, 128, 8, 8);
        run("8 cp.async 16 B, 4 of 8 chunks per row", grid, 8, 128, 4, 8);
        run("8 cp.async 16 B, 2 of 8 chunks per row", grid, 8, 128, 2, 8);
    }
    return 0;
}

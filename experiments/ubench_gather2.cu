// Second gather micro-benchmark: what is the UNIT of the per-SM gather cost (warp instruction, 128-byte line,
// sector, thread op, barrier arrival), and how fast is a sparse expansion (scatter + clear, no zero fill)?
// One CTA per SM, a consumer thread frees each stage at once: the numbers are producer-side rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo ubench_gather2.cu -o ubench_gather2
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async4_ca(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

struct GBars { unsigned long long full[16]; unsigned long long empty[16]; };

// modes (128 producer threads unless stated; a "stage" is always the data of 128 rows):
//  0  cp.async.cg 16 B, 128 B per row (8 lanes per row, 4 rows per warp instruction), row stride `stride` bytes
//  1  no copies, barrier traffic only (floor)
//  2  cp.async.cg 16 B, 32 B per row (2 lanes per row, 16 rows per warp instruction)
//  3  cp.async.cg 16 B, 256 B per row (16 lanes per row, 2 rows per warp instruction), 32 KB per stage
//  4  mode 2 with ld.global.nc.v4 + st.shared.v4
//  5  sparse expansion: per row `nnz` st.shared.u16 of values + `nnz` clears of the previous tenant's positions
//  6  cp.async.ca 4 B per lane: 1 row (128 B) per warp instruction, 128 warp instructions per stage
//  7  mode 0 with cp.async.ca (L1 allocating)
//  8  mode 0, one arrival per WARP (lane 0 after the warp's copies via cp.async.wait_group + __syncwarp): barrier cost
//  9  mode 0 with 64 B per row (4 lanes per row, 8 rows per warp instruction)
__global__ void __launch_bounds__(544, 1) gather_kernel(const unsigned char *q, const int *rows, int nrows_total, int iters, int stages, int mode, int nthr, int nnz, int stride, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ GBars bars;
    __shared__ int srow[2][128];
    const int t = threadIdx.x;
    const int stage_bytes = mode == 3 ? 32768 : 16384;
    int n_arrive = nthr;
    if (mode == 8) n_arrive = nthr / 32;
    if (t == 0) {
        for (int i = 0; i < 16; ++i) {
            mbar_init(smem_u32(&bars.full[i]), n_arrive);
            mbar_init(smem_u32(&bars.empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 128) { srow[0][t] = rows[((size_t)blockIdx.x * 128 + t) % nrows_total]; srow[1][t] = rows[((size_t)blockIdx.x * 128 + 4096 + t) % nrows_total]; }
    for (int i = t; i < stages * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0u;
    __syncthreads();
    const long long t0 = clock64();
    if (t < nthr) {
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < iters; ++it) {
            const int *r = srow[(it / 13) & 1];
            const int kb = it % 12;
            if (mode == 4) {
                uint4 v[2];
                const int c0 = t, c1 = 128 + t;  // 256 chunks of 16 B: row = c >> 1
                v[0] = __ldg(reinterpret_cast<const uint4 *>(q + (size_t)r[c0 >> 1] * stride + kb * 32 + (c0 & 1) * 16));
                v[1] = __ldg(reinterpret_cast<const uint4 *>(q + (size_t)r[c1 >> 1] * stride + kb * 32 + (c1 & 1) * 16));
                if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                unsigned char *dst = sm + stage * stage_bytes;
                *reinterpret_cast<uint4 *>(dst + c0 * 16) = v[0];
                *reinterpret_cast<uint4 *>(dst + c1 * 16) = v[1];
                mbar_arrive(smem_u32(&bars.full[stage]));
            } else {
                if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) break;
                const uint32_t dst = smem_u32(sm) + stage * stage_bytes;
                if (mode == 0 || mode == 7 || mode == 8) {
                    const int per = 1024 / nthr;
#pragma unroll 8
                    for (int j = 0; j < per; ++j) {
                        const int c = j * nthr + t;
                        const int row = c >> 3, ch = c & 7;
                        const unsigned char *src = q + (size_t)r[row] * stride + kb * 128 + ch * 16;
                        if (mode == 7) cp_async16_ca(dst + row * 128 + ((ch ^ (row & 7)) << 4), src);
                        else cp_async16(dst + row * 128 + ((ch ^ (row & 7)) << 4), src);
                    }
                } else if (mode == 9) {
                    const int per = 512 / nthr;
#pragma unroll 4
                    for (int j = 0; j < per; ++j) {
                        const int c = j * nthr + t;
                        const int row = c >> 2, ch = c & 3;
                        cp_async16(dst + row * 128 + ((ch ^ (row & 7)) << 4), q + (size_t)r[row] * stride + kb * 64 + ch * 16);
                    }
                } else if (mode == 2) {
                    const int per = 256 / nthr;
#pragma unroll 2
                    for (int j = 0; j < per; ++j) {
                        const int c = j * nthr + t;
                        cp_async16(dst + c * 16, q + (size_t)r[c >> 1] * stride + kb * 32 + (c & 1) * 16);
                    }
                } else if (mode == 3) {
                    const int per = 2048 / nthr;
#pragma unroll 8
                    for (int j = 0; j < per; ++j) {
                        const int c = j * nthr + t;
                        cp_async16(dst + c * 16, q + (size_t)r[c >> 4] * stride + (kb & 3) * 256 + (c & 15) * 16);
                    }
                } else if (mode == 6) {
                    const int warp = t >> 5, lane = t & 31, nw = nthr >> 5;
#pragma unroll 8
                    for (int row = warp; row < 128; row += nw)
                        cp_async4_ca(dst + row * 128 + lane * 4, q + (size_t)r[row] * stride + kb * 128 + lane * 4);
                } else if (mode == 5) {
                    // row t: clear the previous tenant's nnz positions, write this tile's (positions are a hash of
                    // (row, j, tenant): same spread over banks as real data; swizzled like the MMA operand)
                    unsigned char *drow = sm + stage * stage_bytes + t * 128;
                    const int prev = it - stages;
                    for (int j = 0; j < nnz; ++j) {
                        if (prev >= 0) {
                            const int col = (t * 7 + j * 13 + prev * 5) & 63;
                            *reinterpret_cast<unsigned short *>(drow + ((((col >> 3) ^ (t & 7)) << 4) | ((col & 7) << 1))) = 0;
                        }
                    }
                    for (int j = 0; j < nnz; ++j) {
                        const int col = (t * 7 + j * 13 + it * 5) & 63;
                        *reinterpret_cast<unsigned short *>(drow + ((((col >> 3) ^ (t & 7)) << 4) | ((col & 7) << 1))) = (unsigned short)(0x3c00 + j);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                if (mode == 5) mbar_arrive(smem_u32(&bars.full[stage]));
                else if (mode == 8) {
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncwarp();
                    if ((t & 31) == 0) mbar_arrive(smem_u32(&bars.full[stage]));
                } else cp_async_arrive_noinc(smem_u32(&bars.full[stage]));
            }
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
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

template <int LAG>
__device__ __forceinline__ void wait_lag() { asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory"); }
__device__ __forceinline__ void wait_lag_rt(int lag) {
    switch (lag) {
        case 0: wait_lag<0>(); break;
        case 1: wait_lag<1>(); break;
        case 2: wait_lag<2>(); break;
        case 3: wait_lag<3>(); break;
        case 4: wait_lag<4>(); break;
        case 5: wait_lag<5>(); break;
        default: wait_lag<6>(); break;
    }
}
// one barrier arrival per WARP (an elected lane), cp.async completion tracked with commit groups lagging `lag` stages
//  10 cp.async.cg 16 B x 8 per row    11 barrier only    12 ldg.v4 x 8 prefetched + sts    13 scatter + clear
//  14 cp.async.cg 16 B x 2 per row (32 B rows)    15 cp.async.cg 16 B x 16 per row (256 B rows, 32 KB stage)
__global__ void __launch_bounds__(544, 1) gather2_kernel(const unsigned char *q, const int *rows, int nrows_total, int iters, int stages, int mode, int nthr, int nnz, int stride, int lag, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ GBars bars;
    __shared__ int srow[2][128];
    const int t = threadIdx.x, lane = t & 31;
    const int stage_bytes = mode == 15 ? 32768 : 16384;
    if (t == 0) {
        for (int i = 0; i < 16; ++i) {
            mbar_init(smem_u32(&bars.full[i]), nthr / 32);
            mbar_init(smem_u32(&bars.empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 128) { srow[0][t] = rows[((size_t)blockIdx.x * 128 + t) % nrows_total]; srow[1][t] = rows[((size_t)blockIdx.x * 128 + 4096 + t) % nrows_total]; }
    for (int i = t; i < stages * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0u;
    __syncthreads();
    const long long t0 = clock64();
    if (t < nthr) {
        uint32_t stage = 0, phase = 0;
        const bool asyncm = mode == 10 || mode == 14 || mode == 15;
        uint4 v[8];
        if (mode == 12) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = j * 128 + t;
                v[j] = __ldg(reinterpret_cast<const uint4 *>(q + (size_t)srow[0][c >> 3] * stride + (c & 7) * 16));
            }
        }
        for (int it = 0; it < iters; ++it) {
            const int *r = srow[(it >> 4) & 1];
            const int kb = it & 7;
            if (lane == 0) { if (!mbar_wait_bounded(smem_u32(&bars.empty[stage]), phase ^ 1u)) iters = 0; }
            __syncwarp();
            const uint32_t dst = smem_u32(sm) + stage * stage_bytes;
            if (mode == 10) {
                const int per = 1024 / nthr;
#pragma unroll 8
                for (int j = 0; j < per; ++j) {
                    const int c = j * nthr + t;
                    const int row = c >> 3, ch = c & 7;
                    cp_async16(dst + row * 128 + ((ch ^ (row & 7)) << 4), q + (size_t)r[row] * stride + kb * 128 + ch * 16);
                }
            } else if (mode == 14) {
                const int per = 256 / nthr;
#pragma unroll 2
                for (int j = 0; j < per; ++j) {
                    const int c = j * nthr + t;
                    cp_async16(dst + c * 16, q + (size_t)r[c >> 1] * stride + kb * 32 + (c & 1) * 16);
                }
            } else if (mode == 15) {
                const int per = 2048 / nthr;
#pragma unroll 8
                for (int j = 0; j < per; ++j) {
                    const int c = j * nthr + t;
                    cp_async16(dst + c * 16, q + (size_t)r[c >> 4] * stride + (kb & 3) * 256 + (c & 15) * 16);
                }
            } else if (mode == 12) {
                unsigned char *d = sm + stage * stage_bytes;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = j * 128 + t;
                    const int row = c >> 3, ch = c & 7;
                    *reinterpret_cast<uint4 *>(d + row * 128 + ((ch ^ (row & 7)) << 4)) = v[j];
                }
                const int *rn = srow[((it + 1) >> 4) & 1];
                const int kbn = (it + 1) & 7;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = j * 128 + t;
                    v[j] = __ldg(reinterpret_cast<const uint4 *>(q + (size_t)rn[c >> 3] * stride + kbn * 128 + (c & 7) * 16));
                }
            } else if (mode == 13) {
                unsigned char *drow = sm + stage * stage_bytes + t * 128;
                const int prev = it - stages;
                for (int j = 0; j < nnz; ++j) {
                    if (prev >= 0) {
                        const int col = (t * 7 + j * 13 + prev * 5) & 63;
                        *reinterpret_cast<unsigned short *>(drow + ((((col >> 3) ^ (t & 7)) << 4) | ((col & 7) << 1))) = 0;
                    }
                }
                for (int j = 0; j < nnz; ++j) {
                    const int col = (t * 7 + j * 13 + it * 5) & 63;
                    *reinterpret_cast<unsigned short *>(drow + ((((col >> 3) ^ (t & 7)) << 4) | ((col & 7) << 1))) = (unsigned short)(0x3c00 + j);
                }
            }
            if (asyncm) {
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (it >= lag) {
                    wait_lag_rt(lag);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars.full[(it - lag) % stages]));
                }
            } else {
                if (mode != 11) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars.full[stage]));
            }
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
        }
        if (asyncm) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                for (int it = max(iters - lag, 0); it < iters; ++it) mbar_arrive(smem_u32(&bars.full[it % stages]));
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

int main() {
    CK(cudaSetDevice(0));
    long long *d_out;
    CK(cudaMalloc(&d_out, 64));
    long long h[3];
    const int NQ = 8192;
    unsigned char *d_q;
    CK(cudaMalloc(&d_q, (size_t)NQ * 2048));
    CK(cudaMemset(d_q, 1, (size_t)NQ * 2048));
    const int NR = 1 << 20;
    std::vector<int> hr(NR);
    uint32_t x = 12345;
    for (int i = 0; i < NR; ++i) { x = x * 1664525u + 1013904223u; hr[i] = (x >> 8) % NQ; }
    int *d_rows;
    CK(cudaMalloc(&d_rows, NR * 4));
    CK(cudaMemcpy(d_rows, hr.data(), NR * 4, cudaMemcpyHostToDevice));
    const int gsm = 8 * 16384 + 2048;
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
    auto run = [&](const char *name, int grid, int mode, int nthr, int nnz, int stages, int stride) {
        CK(cudaMemset(d_out, 0, 64));
        const int iters = 2000;
        gather_kernel<<<grid, 544, gsm>>>(d_q, d_rows, NR, iters, stages, mode, nthr, nnz, stride, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
        CK(cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost));
        printf("%-64s thr %3d stride %4d grid %3d stages %d : %8.1f cycles per 128 rows%s\n", name, nthr, stride, grid, stages, (double)h[0] / (h[1] > 0 ? h[1] : 1), h[2] < 0 ? "  (TIMEOUT)" : "");
        fflush(stdout);
    };
    for (int grid : {1, 148}) {
        run("0 cp.async.cg 16 B x 8 per row", grid, 0, 128, 0, 8, 1600);
        run("0 cp.async.cg 16 B x 8 per row, aligned rows", grid, 0, 128, 0, 8, 1664);
        run("0 cp.async.cg 16 B x 8 per row, aligned rows", grid, 0, 256, 0, 8, 1664);
        run("0 cp.async.cg 16 B x 8 per row, 2 KB rows", grid, 0, 128, 0, 8, 2048);
        run("1 barrier traffic only", grid, 1, 128, 0, 8, 1600);
        run("1 barrier traffic only", grid, 1, 256, 0, 8, 1600);
        run("2 cp.async.cg 16 B x 2 per row (32 B rows)", grid, 2, 128, 0, 8, 416);
        run("2 cp.async.cg 16 B x 2 per row (32 B rows)", grid, 2, 256, 0, 8, 416);
        run("2 cp.async.cg 16 B x 2 per row (32 B of 1664 B rows)", grid, 2, 128, 0, 8, 1664);
        run("3 cp.async.cg 16 B x 16 per row (256 B rows)", grid, 3, 128, 0, 4, 256);
        run("3 cp.async.cg 16 B x 16 per row (256 B rows)", grid, 3, 256, 0, 4, 256);
        run("3 cp.async.cg 16 B x 16 per row (256 B of 1024 B rows)", grid, 3, 256, 0, 4, 1024);
        run("4 ldg.v4 + sts.v4, 32 B rows", grid, 4, 128, 0, 8, 416);
        run("5 scatter + clear, 4 per row", grid, 5, 128, 4, 8, 0);
        run("5 scatter + clear, 8 per row", grid, 5, 128, 8, 8, 0);
        run("5 scatter + clear, 4 per row, 4 stages", grid, 5, 128, 4, 4, 0);
        run("6 cp.async.ca 4 B, one row per warp instruction", grid, 6, 128, 0, 8, 1664);
        run("6 cp.async.ca 4 B, one row per warp instruction", grid, 6, 256, 0, 8, 1664);
        run("7 cp.async.ca 16 B x 8 per row", grid, 7, 128, 0, 8, 1664);
        run("8 cp.async.cg 16 B x 8 per row, one arrival per warp", grid, 8, 128, 0, 8, 1664);
        run("8 cp.async.cg 16 B x 8 per row, one arrival per warp", grid, 8, 256, 0, 8, 1664);
        run("9 cp.async.cg 16 B x 4 per row (64 B)", grid, 9, 128, 0, 8, 1664);
    }
    CK(cudaFuncSetAttribute(gather2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
    auto run2 = [&](const char *name, int grid, int mode, int nthr, int nnz, int stages, int stride, int lag) {
        CK(cudaMemset(d_out, 0, 64));
        const int iters = 2000;
        gather2_kernel<<<grid, 544, gsm>>>(d_q, d_rows, NR, iters, stages, mode, nthr, nnz, stride, lag, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
        CK(cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost));
        printf("%-64s thr %3d stride %4d lag %d grid %3d stages %d : %8.1f cycles per 128 rows%s\n", name, nthr, stride, lag, grid, stages, (double)h[0] / (h[1] > 0 ? h[1] : 1), h[2] < 0 ? "  (TIMEOUT)" : "");
        fflush(stdout);
    };
    for (int grid : {1, 148}) {
        run2("11 barrier only, one arrival per warp", grid, 11, 128, 0, 8, 0, 0);
        run2("11 barrier only, one arrival per warp", grid, 11, 256, 0, 8, 0, 0);
        for (int lag : {1, 2, 3, 4, 6}) run2("10 cp.async.cg 16 B x 8 per row, warp arrival", grid, 10, 128, 0, 8, 1664, lag);
        for (int lag : {2, 4, 6}) run2("10 cp.async.cg 16 B x 8 per row, warp arrival", grid, 10, 256, 0, 8, 1664, lag);
        for (int lag : {2, 4, 6}) run2("10 cp.async.cg 16 B x 8 per row, warp arrival", grid, 10, 512, 0, 8, 1664, lag);
        run2("10 cp.async.cg 16 B x 8 per row, warp arrival, 1600", grid, 10, 128, 0, 8, 1600, 4);
        run2("12 ldg.v4 x 8 prefetched + sts, warp arrival", grid, 12, 128, 0, 8, 1664, 0);
        run2("13 scatter + clear 4 per row, warp arrival", grid, 13, 128, 4, 8, 0, 0);
        run2("13 scatter + clear 8 per row, warp arrival", grid, 13, 128, 8, 8, 0, 0);
        run2("13 scatter + clear 4 per row, warp arrival, 4 stages", grid, 13, 128, 4, 4, 0, 0);
        for (int lag : {2, 4}) run2("14 cp.async.cg 16 B x 2 per row (32 B rows), warp arrival", grid, 14, 128, 0, 8, 416, lag);
        for (int lag : {1, 2, 3}) run2("15 cp.async.cg 16 B x 16 per row (256 B rows), warp arrival", grid, 15, 256, 0, 4, 256, lag);
    }
    return 0;
}

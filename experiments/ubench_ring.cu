// Minimal mbarrier ring: NP producer warps (lane 0 arrives on full[s], count NP) and one consumer thread that
// frees the stage at once. No data. Per-iteration timestamps of producer warp 0 and of the consumer are traced.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo ubench_ring.cu -o ubench_ring
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_test(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWL2:\nmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD2;\nbra WL2;\nWD2:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
struct Bars { unsigned long long full[16]; unsigned long long empty[16]; };
// mode 0: try_wait loops; 1: test_wait (polling) loops; 2: all 32 producer lanes arrive (count NP * 32)
__global__ void __launch_bounds__(544, 1) ring_kernel(int iters, int stages, int np, int mode, int work, long long *out, int *trace) {
    __shared__ Bars bars;
    __shared__ volatile int sink;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) {
        for (int i = 0; i < 16; ++i) {
            mbar_init(smem_u32(&bars.full[i]), mode == 2 ? np * 32 : np);
            mbar_init(smem_u32(&bars.empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp < np) {
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < iters; ++it) {
            if (mode == 2 || lane == 0) {
                if (mode == 1) mbar_wait_test(smem_u32(&bars.empty[stage]), phase ^ 1u);
                else mbar_wait(smem_u32(&bars.empty[stage]), phase ^ 1u);
            }
            __syncwarp();
            for (int w = 0; w < work; ++w) sink = w;  // stand-in for producer work (smem stores)
            __syncwarp();
            if (mode == 2 || lane == 0) mbar_arrive(smem_u32(&bars.full[stage]));
            if (warp == 0 && lane == 0 && it < 256) trace[it] = (int)(clock64() - t0);
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
        }
    } else if (t == 512) {
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < iters; ++it) {
            if (mode == 1) mbar_wait_test(smem_u32(&bars.full[stage]), phase);
            else mbar_wait(smem_u32(&bars.full[stage]), phase);
            mbar_arrive(smem_u32(&bars.empty[stage]));
            if (it < 256) trace[256 + it] = (int)(clock64() - t0);
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
        }
        out[0] = clock64() - t0;
    }
    __syncthreads();
}
int main() {
    long long *d_out; int *d_tr;
    CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_tr, 512 * 4));
    int tr[512]; long long h;
    for (int mode : {0, 1, 2}) for (int np : {1, 4, 8}) for (int stages : {2, 4, 8}) for (int work : {0, 8}) {
        const int iters = 4000;
        ring_kernel<<<1, 544>>>(iters, stages, np, mode, work, d_out, d_tr);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(tr, d_tr, sizeof(tr), cudaMemcpyDeviceToHost));
        printf("mode %d producers %d stages %d work %d : %7.1f cycles per stage | producer deltas:", mode, np, stages, work, (double)h / iters);
        for (int i = 200; i < 212; ++i) printf(" %d", tr[i] - tr[i - 1]);
        printf(" | consumer deltas:");
        for (int i = 200; i < 212; ++i) printf(" %d", tr[256 + i] - tr[256 + i - 1]);
        printf("\n");
    }
    return 0;
}

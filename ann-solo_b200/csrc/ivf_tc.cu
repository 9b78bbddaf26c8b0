// K3 — IVF list scan on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
// One work tile = (inverted list l, block of 128 queries of the list's query group). The list's
// vectors (fp16, list order, contiguous rows) are the B operand, streamed by TMA
// (cp.async.bulk.tensor.2d, SWIZZLE_128B); the 128 gathered query rows (fp16) are the A operand,
// written by four producer warps with 16-byte cp.async into the same swizzled K-major layout.
// A single thread issues tcgen05.mma (M=128, N=list length rounded up to 16, K=16 per
// instruction) into one of two TMEM accumulator buffers; four epilogue warps read the other
// buffer with tcgen05.ld (one query row per thread), compare each score with the query's running
// threshold and append (score, position) pairs to the query's candidate buffer — scores that do
// not pass never leave the SM. Products of fp16 inputs are exact in fp32; the scores are
// approximations of the oracle's fp32 fmaf chain with a proven error bound (IVF_REL_EPS), and
// K4 (ivf.cu) re-scores the band around the k-th score exactly, so the top-k is bit-exact.
#include <cuda.h>

#include "ivf.cuh"

namespace solo {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 64;      // fp16 elements per k-block = 128 bytes = one swizzle atom row
constexpr int TC_STAGES = 4;
constexpr int TC_BOX = 64;     // rows per TMA box
constexpr int TC_A_BYTES = TC_BM * 128;
constexpr int TC_B_BYTES = TC_BN * 128;
constexpr int TC_LAG = 2;      // A-producer signal lag (cp.async groups in flight)
constexpr int TC_THREADS = 320;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4,
// leading byte offset (unused for swizzled K-major, set to 1), stride byte offset = 1024 B between
// 8-row groups, descriptor version 1, layout type 2 (128-byte swizzle).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

struct TcScanArgs {
    const int64_t *goff;      // [nlist+1] query-group offsets of this round
    const int32_t *gq;        // grouped query ids
    const int64_t *tile_off;  // [nlist+1] exclusive scan of tiles per list
    const int64_t *list_off;  // [nlist+1]
    const __half *qh;         // (nq, dim) fp16 scaled queries
    int nlist;
    int dim;
    float inv_scale;          // 2^-(scale_index + scale_query)
    const float *tau;         // [nq]
    unsigned long long *buf;  // [nq][cap]
    int32_t *cnt;             // [nq]
    int cap;
};

struct __align__(8) TcBarriers {
    unsigned long long full[TC_STAGES];
    unsigned long long empty[TC_STAGES];
    unsigned long long tmem_full[2];
    unsigned long long tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ int find_list(const int64_t *tile_off, int nlist, int64_t tile) {
    int lo = 0, hi = nlist - 1;  // largest l with tile_off[l] <= tile
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (tile_off[mid] <= tile) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_vec, TcScanArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    // SWIZZLE_128B atoms need a 1024-byte aligned base: align by hand (1 KB of slack is allocated)
    unsigned char *tc_smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    // [A stages][B stages][barriers]
    unsigned char *sA = tc_smem;
    unsigned char *sB = tc_smem + TC_STAGES * TC_A_BYTES;
    TcBarriers *bars = reinterpret_cast<TcBarriers *>(sB + TC_STAGES * TC_B_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = a.tile_off[a.nlist];
    const int num_kb = (a.dim + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 128 + 1);  // 128 A-producer threads + the TMA thread
            mbar_init(smem_u32(&bars->empty[s]), 1);       // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->tmem_full[b]), 1);   // tcgen05.commit
            mbar_init(smem_u32(&bars->tmem_empty[b]), 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {  // TMEM: 2 accumulator buffers x 256 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 4) {
        // ================= epilogue: TMEM -> registers -> threshold -> append =================
        uint32_t unit = 0;
        const int row = warp * 32 + lane;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int l = find_list(a.tile_off, a.nlist, tile);
            const int qb = (int)(tile - a.tile_off[l]);
            const int64_t g0 = a.goff[l];
            const int G = (int)(a.goff[l + 1] - g0);
            const int64_t p0 = a.list_off[l];
            const int len = (int)(a.list_off[l + 1] - p0);
            const int gi = qb * TC_BM + row;
            const int q = gi < G ? a.gq[g0 + gi] : -1;
            const float thr = q >= 0 ? a.tau[q] : INFINITY;
            unsigned long long *qbuf = q >= 0 ? a.buf + (int64_t)q * a.cap : nullptr;
            for (int n0 = 0; n0 < len; n0 += TC_BN, ++unit) {
                const int buf = unit & 1;
                const int nt = min(TC_BN, len - n0);
                mbar_wait(smem_u32(&bars->tmem_full[buf]), (unit >> 1) & 1);
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * TC_BN);
                for (int c0 = 0; c0 < nt; c0 += 32) {
                    uint32_t r[32];
                    tc_ld32(tbase + (uint32_t)c0, r);
                    tc_wait_ld();
                    if (c0 + 32 >= nt) {  // last chunk of this accumulator: hand the buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[buf]));
                    }
                    if (q >= 0) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float s = __uint_as_float(r[j]) * a.inv_scale;
                            r[j] = __float_as_uint(s);
                            if (c0 + j < nt && s >= thr) mask |= 1u << j;
                        }
                        if (mask) {
                            const int n = __popc(mask);
                            int slot = atomicAdd(&a.cnt[q], n);
                            const uint32_t pbase = (uint32_t)(p0 + n0 + c0);
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if ((mask >> j) & 1u) {
                                    if (slot < a.cap)
                                        qbuf[slot] = ((unsigned long long)r[j] << 32) | (unsigned long long)(pbase + j);
                                    ++slot;
                                }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ================= TMA producer: list vectors (B operand) =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int l = find_list(a.tile_off, a.nlist, tile);
                const int64_t p0 = a.list_off[l];
                const int len = (int)(a.list_off[l + 1] - p0);
                for (int n0 = 0; n0 < len; n0 += TC_BN) {
                    const int nt = min(TC_BN, len - n0);
                    const int nbox = (nt + TC_BOX - 1) / TC_BOX;
                    for (int kb = 0; kb < num_kb; ++kb, ++it) {
                        const int s = it % TC_STAGES;
                        mbar_wait(smem_u32(&bars->empty[s]), ((it / TC_STAGES) & 1) ^ 1);
                        const uint32_t fb = smem_u32(&bars->full[s]);
                        mbar_expect_tx(fb, (uint32_t)(nbox * TC_BOX * 128));
                        for (int j = 0; j < nbox; ++j)
                            tma_load_2d(smem_u32(sB + s * TC_B_BYTES + j * TC_BOX * 128), &tmap_vec, kb * TC_BK,
                                        (int)(p0 + n0 + j * TC_BOX), fb);
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t it = 0, unit = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int l = find_list(a.tile_off, a.nlist, tile);
                const int len = (int)(a.list_off[l + 1] - a.list_off[l]);
                for (int n0 = 0; n0 < len; n0 += TC_BN, ++unit) {
                    const int buf = unit & 1;
                    const int nt = min(TC_BN, len - n0);
                    const uint32_t idesc = make_idesc_f16((nt + 15) & ~15);
                    mbar_wait(smem_u32(&bars->tmem_empty[buf]), ((unit >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TC_BN);
                    for (int kb = 0; kb < num_kb; ++kb, ++it) {
                        const int s = it % TC_STAGES;
                        mbar_wait(smem_u32(&bars->full[s]), (it / TC_STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sA + s * TC_A_BYTES);
                        const uint32_t b_addr = smem_u32(sB + s * TC_B_BYTES);
                        const int ksteps = min(TC_BK, a.dim - kb * TC_BK) / 16;
                        for (int k = 0; k < ksteps; ++k) {
                            tc_mma_f16(tmem_d, make_desc_sw128(a_addr + k * 32), make_desc_sw128(b_addr + k * 32), idesc,
                                       (kb | k) != 0 ? 1u : 0u);
                        }
                        tc_commit(smem_u32(&bars->empty[s]));  // frees the stage when these MMAs retire
                    }
                    tc_commit(smem_u32(&bars->tmem_full[buf]));
                }
            }
        }
    } else {
        // ================= A producers: gather 128 query rows per k-block =================
        const int p = threadIdx.x - 6 * 32;  // 0..127
        const int chunk = p & 7;             // 16-byte chunk inside the 128-byte k-block row
        const int rbase = p >> 3;            // rows rbase, rbase+16, ..., rbase+112
        uint32_t it = 0;
        uint32_t pending[TC_LAG + 1];
        int npend = 0;
        const size_t row_bytes = (size_t)a.dim * sizeof(__half);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int l = find_list(a.tile_off, a.nlist, tile);
            const int qb = (int)(tile - a.tile_off[l]);
            const int64_t g0 = a.goff[l];
            const int G = (int)(a.goff[l + 1] - g0);
            const int len = (int)(a.list_off[l + 1] - a.list_off[l]);
            const unsigned char *src[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int gi = qb * TC_BM + rbase + 16 * i;
                int q = a.gq[g0 + (gi < G ? gi : 0)];  // padding rows replay a valid query; the epilogue skips them
                src[i] = reinterpret_cast<const unsigned char *>(a.qh) + (size_t)q * row_bytes + chunk * 16;
            }
            for (int n0 = 0; n0 < len; n0 += TC_BN) {
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(smem_u32(&bars->empty[s]), ((it / TC_STAGES) & 1) ^ 1);
                    const bool valid = kb * TC_BK + chunk * 8 < a.dim;  // last k-block may be partial
                    if (valid) {
                        const uint32_t dst0 = smem_u32(sA + s * TC_A_BYTES);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rbase + 16 * i;
                            cp_async16(dst0 + r * 128 + ((chunk ^ (r & 7)) << 4), src[i] + (size_t)kb * 128);
                        }
                    }
                    cp_async_commit();
                    pending[npend++] = smem_u32(&bars->full[s]);
                    if (npend > TC_LAG) {
                        cp_async_wait<TC_LAG>();
                        fence_proxy_async();
                        mbar_arrive(pending[0]);
#pragma unroll
                        for (int j = 0; j < TC_LAG; ++j) pending[j] = pending[j + 1];
                        --npend;
                    }
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (int j = 0; j < npend; ++j) mbar_arrive(pending[j]);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

__global__ void tc_tile_count_kernel(const int64_t *__restrict__ goff, const int64_t *__restrict__ list_off, int nlist,
                                     int32_t *__restrict__ cnt) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    int64_t G = goff[l + 1] - goff[l];
    int64_t len = list_off[l + 1] - list_off[l];
    cnt[l] = (len > 0 && G > 0) ? (int32_t)((G + TC_BM - 1) / TC_BM) : 0;
}

// ---------------------------------------------------------------- host side

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SOLO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        SOLO_REQUIRE(qres == cudaDriverEntryPointSuccess && p, SOLO_ECUDA, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

bool tc_scan_supported(const IvfIndex &ix) { return ix.dim % 16 == 0 && ix.dim >= 16; }

// (re)build the TMA descriptor of the list-ordered fp16 vectors; called from ivf_finalize
void tc_make_tensor_map(IvfIndex &ix) {
    ix.tmap_valid = false;
    if (!tc_scan_supported(ix) || ix.nstored == 0) return;
    static_assert(sizeof(CUtensorMap) <= sizeof(ix.tmap_storage), "tensor map storage too small");
    CUtensorMap *m = reinterpret_cast<CUtensorMap *>(ix.tmap_storage);
    cuuint64_t gdim[2] = {(cuuint64_t)ix.dim, (cuuint64_t)ix.nstored};
    cuuint64_t gstr[1] = {(cuuint64_t)ix.dim * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BOX};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ix.vec_h.p, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SOLO_REQUIRE(r == CUDA_SUCCESS, SOLO_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    ix.tmap_valid = true;
}

void launch_scan_tc(solo_handle *h, IvfIndex &ix, const int64_t *goff, const int32_t *gq, const __half *qh,
                    int q_scale_log2, const float *tau, unsigned long long *buf, int32_t *cnt, int cap,
                    DevBuf &tile_cnt, DevBuf &tile_off) {
    SOLO_REQUIRE(ix.tmap_valid, SOLO_ESTATE, "tensor map missing");
    const int nlist = ix.nlist;
    tile_cnt.ensure((size_t)nlist * sizeof(int32_t));
    tile_off.ensure((size_t)(nlist + 1) * sizeof(int64_t));
    tc_tile_count_kernel<<<div_up(nlist, 256), 256, 0, h->stream>>>(goff, ix.list_off.as<int64_t>(), nlist,
                                                                    tile_cnt.as<int32_t>());
    scan_counts_i32(h, tile_cnt.as<int32_t>(), nlist, tile_off.as<int64_t>());
    TcScanArgs a;
    a.goff = goff;
    a.gq = gq;
    a.tile_off = tile_off.as<int64_t>();
    a.list_off = ix.list_off.as<int64_t>();
    a.qh = qh;
    a.nlist = nlist;
    a.dim = ix.dim;
    a.inv_scale = ldexpf(1.f, -(ix.scale_log2 + q_scale_log2));
    a.tau = tau;
    a.buf = buf;
    a.cnt = cnt;
    a.cap = cap;
    const size_t smem = (size_t)TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + sizeof(TcBarriers) + 1024;
    SOLO_CUDA(cudaFuncSetAttribute(scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUtensorMap map;
    memcpy(&map, ix.tmap_storage, sizeof map);
    scan_tc_kernel<<<kNumSMs, TC_THREADS, smem, h->stream>>>(map, a);
    SOLO_CUDA(cudaGetLastError());
    h->launches += 2;
}

}  // namespace solo

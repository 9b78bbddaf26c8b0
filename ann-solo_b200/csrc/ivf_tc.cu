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

#include <algorithm>
#include <vector>

#include "ivf.cuh"

// unroll factors of the two straight-line loops of scan_tc_kernel (the kernel is sensitive to its code size)
#ifndef K3_UNROLL_MMA
#define K3_UNROLL_MMA 1
#endif
#ifndef K3_UNROLL_PROD
#define K3_UNROLL_PROD 13
#endif

namespace solo {

constexpr int K3_UM = K3_UNROLL_MMA, K3_UP = K3_UNROLL_PROD;
constexpr int TC_BM = 128;       // queries per accumulator tile (MMA M)
constexpr int TC_BK = 64;        // fp16 elements per k-block = 128 bytes = one swizzle atom row
constexpr int TC_MAX_STAGES = 12; // A (query) stages: as many as fit next to the resident list chunk
constexpr int TC_BOX = 32;       // rows per TMA box
constexpr int TC_A_BYTES = TC_BM * 128;
constexpr int TC_MAX_KB = 24;    // dim <= 1536
#ifndef K3_PRODUCERS
#define K3_PRODUCERS 128
#endif
constexpr int TC_PRODUCERS = K3_PRODUCERS;  // A-producer threads (warps 10..)
constexpr int TC_EPI_WARPS = 8;   // two sets of four: set s drains accumulator buffer s
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32 + TC_PRODUCERS;
constexpr int TC_SMEM_MAX = 232448;  // 227 KB opt-in limit per CTA

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
// one probe of the barrier phase; the predicate can be consumed later (latency overlaps other issue work)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// two probes in flight together (each costs ~90 cycles until its predicate can be read)
__device__ __forceinline__ void mbar_try_wait2(uint32_t bar0, uint32_t par0, uint32_t bar1, uint32_t par1, bool &ok0, bool &ok1) {
    uint32_t r0, r1;
    asm volatile(
        "{\n"
        ".reg .pred p0, p1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p0, [%2], %3;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p1, [%4], %5;\n"
        "selp.u32 %0, 1, 0, p0;\n"
        "selp.u32 %1, 1, 0, p1;\n"
        "}\n"
        : "=r"(r0), "=r"(r1)
        : "r"(bar0), "r"(par0), "r"(bar1), "r"(par1)
        : "memory");
    ok0 = r0 != 0;
    ok1 = r1 != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// 16-byte asynchronous copy; src_bytes == 0 writes zeros without touching global memory
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// this thread's arrival on `bar` is triggered when all its earlier cp.async copies have landed (the
// barrier's expected count includes it: .noinc). No blocking wait in the producer, every stage can be in flight.
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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

// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

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
    const TcItem *items;
    const int64_t *item_off;  // [nlist+1]; item_off[nlist] = number of items
    const int32_t *gq;        // grouped query ids
    const __half *qh;         // (nq, dim) fp16 scaled queries
    const uint32_t *qmask;    // (nq, 8): bit kb of word c = 16-byte chunk c of k-block kb is non-zero
    int nlist;
    int dim;
    int nb;                   // chunk capacity NB (multiple of 32)
    int stages;               // A stages in the ring
    int kbb;                  // k-blocks per MMA issue batch
    int wide;                 // 1: list vectors are streamed per (tile, k-block) next to the query stage (chunks of up to 256 rows)
    float inv_scale;          // 2^-(scale_index + scale_query)
    const float *tau;         // [nq]
    unsigned long long *buf;  // [nq][cap]
    int32_t *cnt;             // [nq]
    int cap;
    float *dense_out;         // MODE 1: (nq, dense_ld) all scores, row = query (gq == null: identity)
    int dense_ld;
    unsigned long long *prof; // optional (SOLO_TC_PROF=1): per-CTA wait-cycle counters, 8 per CTA
    int debug;                // timing experiments only (SOLO_TC_DEBUG): 1 = skip query copies, 2 = skip score handling
    int a_col0;               // scan_ts_kernel: first tensor-memory column of the resident list chunk (the accumulator sits in front)
};

// wait on an mbarrier; when profiling, add the cycles spent to *acc
__device__ __forceinline__ void mbar_wait_prof(uint32_t bar, uint32_t parity, bool on, unsigned long long &acc) {
    if (on) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += (unsigned long long)(clock64() - t0);
    } else {
        mbar_wait(bar, parity);
    }
}

struct __align__(8) TcBarriers {
    unsigned long long full_a[TC_MAX_STAGES];
    unsigned long long empty_a[TC_MAX_STAGES];
    unsigned long long full_b[TC_MAX_KB];
    unsigned long long empty_b[TC_MAX_KB];
    unsigned long long tmem_full[2];
    unsigned long long tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

// Walks the (item, query block) tiles of one CTA; the next item's descriptor is fetched one item ahead.
struct TileCursor {
    const TcItem *items;
    int n_items, stride;
    int item, qb, nqb;
    int bm;  // queries per tile (128; 256 for CTA pairs)
    TcItem cur, nxt;
    bool valid;
    __device__ __forceinline__ void load_next() {
        const int ni = item + stride;
        if (ni < n_items) nxt = items[ni];
    }
    __device__ __forceinline__ void init(const TcItem *it, int n, int first, int step, int tile_m = TC_BM) {
        bm = tile_m;
        items = it;
        n_items = n;
        stride = step;
        item = first;
        qb = 0;
        valid = item < n_items;
        if (valid) cur = items[item];
        nqb = valid ? (cur.G + bm - 1) / bm : 0;
        load_next();
    }
    __device__ __forceinline__ bool first_qb() const { return qb == 0; }
    __device__ __forceinline__ bool last_qb() const { return qb == nqb - 1; }
    __device__ __forceinline__ void advance() {
        if (++qb < nqb) return;
        qb = 0;
        item += stride;
        valid = item < n_items;
        cur = nxt;
        nqb = (cur.G + bm - 1) / bm;
        if (valid) load_next();
    }
};

// Persistent, warp-specialised. Shared memory: [A stages][B: num_kb k-blocks x NB rows x 128 B][barriers].
// The list chunk (B operand) stays resident while all query blocks of the item stream through the
// A stages; its k-block slots are released one by one during the item's last query block so the
// next item's vectors arrive behind the last reader.
// MODE 0: list scan (threshold + append). MODE 1: dense output (coarse quantizer: the "list" is the
// centroid table, the query group is every query).
// NUM_KB > 0 fixes the number of k-blocks at compile time (dim 800 -> 13) so that the gather loop
// unrolls into copies with immediate offsets; NUM_KB == 0 is the generic kernel.
// FLAGS & 4: the query group of every item is the identity (coarse quantizer): an A tile is 128 CONSECUTIVE rows
// of the fp16 query matrix and comes in by TMA (tmap_q, one box per k-block) instead of the 16-byte gather.
template <int MODE, int NUM_KB, int EPI, int FLAGS>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_vec, const __grid_constant__ CUtensorMap tmap_q, TcScanArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    // SWIZZLE_128B atoms need a 1024-byte aligned base: align by hand (1 KB of slack is allocated)
    unsigned char *tc_smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    const int num_kb = NUM_KB > 0 ? NUM_KB : (a.dim + TC_BK - 1) / TC_BK;
    const int b_kb_bytes = a.nb * 128;
    unsigned char *sA = tc_smem;
    const int n_stages = a.stages;
    unsigned char *sB = tc_smem + n_stages * TC_A_BYTES;
    // wide mode: one B buffer of nb rows per stage; resident mode: one per k-block
    const bool wide = a.wide != 0;
    TcBarriers *bars = reinterpret_cast<TcBarriers *>(sB + (size_t)(wide ? n_stages : num_kb) * b_kb_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = (int)a.item_off[a.nlist];
    const bool prof_on = a.prof != nullptr;
    unsigned long long pw0 = 0, pw1 = 0, pw2 = 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) {
            mbar_init(smem_u32(&bars->full_a[s]), (FLAGS & 4) ? 1 : TC_PRODUCERS);  // the A-producer threads (or the TMA issuer)
            mbar_init(smem_u32(&bars->empty_a[s]), 1);            // tcgen05.commit
        }
        for (int kb = 0; kb < TC_MAX_KB; ++kb) {
            mbar_init(smem_u32(&bars->full_b[kb]), 1);   // TMA thread (expect_tx)
            mbar_init(smem_u32(&bars->empty_b[kb]), 1);  // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->tmem_full[b]), 1);   // tcgen05.commit
            mbar_init(smem_u32(&bars->tmem_empty[b]), 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_EPI_WARPS) {  // TMEM: 2 accumulator buffers x 256 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < TC_EPI_WARPS) {
        // ================= epilogue: TMEM -> registers -> threshold -> append =================
        // Two sets of four warps; set s owns accumulator buffer s, i.e. every other tile, so each set
        // has two tile periods for its global-memory round trips. Warp w reads TMEM lanes 32 (w % 4)..
        const int set = warp >> 2;
        uint32_t unit = (uint32_t)set;
        const int row = (warp & 3) * 32 + lane;
        const long long t_begin = prof_on ? clock64() : 0;
        const float scale = 1.f / a.inv_scale;  // exact power of two
        TileCursor tc;
        tc.init(a.items, n_items, blockIdx.x, gridDim.x);
        if (set == 1 && tc.valid) tc.advance();
        // the row's query and threshold are fetched one (own) tile ahead
        int q = -1;
        float thr = INFINITY;
        if (tc.valid) {
            const int gi = tc.qb * TC_BM + row;
            q = gi < tc.cur.G ? (a.gq ? a.gq[tc.cur.g0 + gi] : tc.cur.g0 + gi) : -1;
            if (MODE == 0 && q >= 0) thr = a.tau[q] * scale;
        }
        while (tc.valid) {
            const TcItem it = tc.cur;
            TileCursor nx = tc;
            nx.advance();
            if (nx.valid) nx.advance();
            int q_next = -1;
            if (nx.valid) {
                const int gi = nx.qb * TC_BM + row;
                q_next = gi < nx.cur.G ? (a.gq ? a.gq[nx.cur.g0 + gi] : nx.cur.g0 + gi) : -1;
            }
            const int buf = set;
            mbar_wait_prof(smem_u32(&bars->tmem_full[buf]), (unit >> 1) & 1, prof_on, pw0);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * 256);
            if (MODE == 1) {
                for (int c0 = 0; c0 < it.nv; c0 += 32) {
                    uint32_t r[32];
                    tc_ld32(tbase + (uint32_t)c0, r);
                    tc_wait_ld();
                    if (c0 + 32 >= it.nv) {  // last chunk of this accumulator: hand the buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[buf]));
                    }
                    if (q < 0) continue;
                    // the row pitch is a multiple of 32 floats and p0 + c0 a multiple of 32: aligned, in-row
                    float4 *dst = reinterpret_cast<float4 *>(a.dense_out + (size_t)q * a.dense_ld + it.p0 + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(r[4 * j]) * a.inv_scale,
                                             __uint_as_float(r[4 * j + 1]) * a.inv_scale,
                                             __uint_as_float(r[4 * j + 2]) * a.inv_scale,
                                             __uint_as_float(r[4 * j + 3]) * a.inv_scale);
                }
            } else {
                // groups of up to 128 columns: read them all, release the accumulator, then one
                // atomic reservation per group and the (rare) stores
                for (int g0 = 0; g0 < it.nv; g0 += 32 * EPI) {
                    uint32_t r[EPI][32];
#pragma unroll
                    for (int c = 0; c < EPI; ++c)
                        if (g0 + 32 * c < it.nv) tc_ld32(tbase + (uint32_t)(g0 + 32 * c), r[c]);
                    tc_wait_ld();
                    if (g0 + 32 * EPI >= it.nv) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[buf]));
                    }
                    if (q < 0 || (a.debug & 2)) continue;
                    uint32_t mask[EPI];
                    int n = 0;
#pragma unroll
                    for (int c = 0; c < EPI; ++c) {
                        mask[c] = 0;
                        if (g0 + 32 * c < it.nv) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)  // thr is pre-scaled: raw accumulators are compared
                                if (__uint_as_float(r[c][j]) >= thr) mask[c] |= 1u << j;
                            const int left = it.nv - (g0 + 32 * c);  // valid columns in this chunk
                            if (left < 32) mask[c] &= (1u << left) - 1u;
                            n += __popc(mask[c]);
                        }
                    }
                    if (n) {
                        int slot = atomicAdd(&a.cnt[q], n);
                        unsigned long long *qbuf = a.buf + (int64_t)q * a.cap;
#pragma unroll
                        for (int c = 0; c < EPI; ++c) {
                            if (mask[c]) {
                                const uint32_t pbase = (uint32_t)(it.p0 + g0 + 32 * c);
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    if ((mask[c] >> j) & 1u) {
                                        if (slot < a.cap)
                                            qbuf[slot] = ((unsigned long long)__float_as_uint(__uint_as_float(r[c][j]) * a.inv_scale) << 32) |
                                                         (unsigned long long)(pbase + j);
                                        ++slot;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            q = q_next;
            thr = (MODE == 0 && q >= 0) ? a.tau[q] * scale : INFINITY;  // lands while waiting for the next accumulator
            tc = nx;
            unit += 2;
        }
        if (prof_on && threadIdx.x == 0) {
            a.prof[blockIdx.x * 8 + 0] = pw0;                                          // epilogue: wait tmem_full
            a.prof[blockIdx.x * 8 + 7] = (unsigned long long)(clock64() - t_begin);    // total
        }
    } else if (warp == TC_EPI_WARPS) {
        // ================= TMA producer: the item's list vectors (B operand, resident) =================
        // The whole warp walks the loop (uniform control flow), one elected lane issues the copies.
        uint32_t n_local = 0;
        TcItem it, nxt;
        int item = blockIdx.x;
        if (wide) {
            // the chunk's k-block is loaded again for every query block, into the stage the gather fills
            // (it comes from L2 after the first query block); full_b[stage] carries the transaction count
            uint32_t stage = 0, phase = 0;
            TileCursor tc;
            tc.init(a.items, n_items, blockIdx.x, gridDim.x);
            while (tc.valid) {
                const int nbox = (tc.cur.nv + TC_BOX - 1) / TC_BOX;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_prof(smem_u32(&bars->empty_a[stage]), phase ^ 1u, prof_on, pw0);
                    if (elect_one()) {
                        const uint32_t fb = smem_u32(&bars->full_b[stage]);
                        mbar_expect_tx(fb, (uint32_t)(nbox * TC_BOX * 128));
                        for (int j = 0; j < nbox; ++j)
                            tma_load_2d(smem_u32(sB + (size_t)stage * b_kb_bytes + j * TC_BOX * 128), &tmap_vec, kb * TC_BK,
                                        tc.cur.p0 + j * TC_BOX, fb);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc.advance();
            }
            item = n_items;  // skip the resident-mode loop below
        }
        if (item < n_items) it = a.items[item];
        for (; item < n_items; item += gridDim.x, ++n_local) {
            if (item + (int)gridDim.x < n_items) nxt = a.items[item + gridDim.x];
            const int nbox = (it.nv + TC_BOX - 1) / TC_BOX;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_prof(smem_u32(&bars->empty_b[kb]), (n_local & 1) ^ 1, prof_on, pw0);
                if (elect_one()) {
                    const uint32_t fb = smem_u32(&bars->full_b[kb]);
                    mbar_expect_tx(fb, (uint32_t)(nbox * TC_BOX * 128));
                    for (int j = 0; j < nbox; ++j)
                        tma_load_2d(smem_u32(sB + (size_t)kb * b_kb_bytes + j * TC_BOX * 128), &tmap_vec, kb * TC_BK,
                                    it.p0 + j * TC_BOX, fb);
                }
                __syncwarp();
            }
            it = nxt;
        }
        if (prof_on && lane == 0) a.prof[blockIdx.x * 8 + 1] = pw0;  // TMA: wait empty_b
    } else if (warp == TC_EPI_WARPS + 1) {
        // ================= MMA issuer =================
        // Uniform control flow for the whole warp; one elected lane issues tcgen05.mma / commit.
        const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        uint32_t stage = 0, phase = 0, unit = 0, n_local = 0;
        unsigned long long pq0 = 0, pq1 = 0, pq2 = 0;
        bool ready = false, ready1 = false;  // outcome of the early probes of the next batch's query stages
        uint32_t g = 0;      // running k-block counter (straight-line path)
        const int kbb = max(1, min(a.kbb, n_stages / 2));  // k-blocks per issue batch
        const int last_ksteps = (a.dim - (num_kb - 1) * TC_BK) / 16;
        TileCursor tc;
        tc.init(a.items, n_items, blockIdx.x, gridDim.x);
        while (tc.valid) {
            const uint32_t idesc = make_idesc_f16((a.debug & 4) ? 48 : (a.debug & 8) ? 16 : (tc.cur.nv + 15) & ~15);
            const int buf = unit & 1;
            mbar_wait_prof(smem_u32(&bars->tmem_empty[buf]), ((unit >> 1) & 1) ^ 1, prof_on, pw0);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 256);
            const bool first_qb = tc.first_qb(), last_qb = tc.last_qb();
            // k-blocks are issued in batches of KBB = 2: the first tcgen05.mma after a tcgen05.commit costs
            // ~290 cycles of issue time (measured), the following ones ~15, so the commits of a batch go
            // out together after its last MMA. This warp is the serial bottleneck of the kernel: with
            // NUM_KB fixed and a four-stage ring everything below unrolls into straight-line code
            // (stage = g & 3, phase = (g >> 2) & 1 for the running k-block counter g).
            if (NUM_KB > 0 && n_stages == 4 && !(a.debug & 256)) {
                constexpr int KBB = 2;
                constexpr int NKB = NUM_KB > 0 ? NUM_KB : 1;
#pragma unroll K3_UM
                for (int kb0 = 0; kb0 < NKB; kb0 += KBB) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int nkb = kb0 + KBB <= NKB ? KBB : NKB - kb0;
#pragma unroll
                    for (int j = 0; j < KBB; ++j) {
                        if (j < nkb) {
                            const uint32_t gj = g + (uint32_t)j;
                            if (wide) mbar_wait(smem_u32(&bars->full_b[gj & 3u]), (gj >> 2) & 1u);
                            else if (first_qb) mbar_wait(smem_u32(&bars->full_b[kb0 + j]), n_local & 1);
                            // both query stages of the batch were probed behind the previous batch's commits (below):
                            // when the gather is ahead nothing is waited for here
                            if (!(j == 0 ? ready : ready1)) mbar_wait(smem_u32(&bars->full_a[gj & 3u]), (gj >> 2) & 1u);
                        }
                    }
                    const uint32_t gn = g + (uint32_t)nkb;
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < KBB; ++j) {
                            if (j < nkb) {
                                const int kb = kb0 + j;
                                const uint32_t st = (g + (uint32_t)j) & 3u;
                                uint64_t adesc = desc_hi | (uint64_t)(((a_base + st * TC_A_BYTES) >> 4) & 0x3FFFu);
                                uint64_t bdesc = desc_hi | (uint64_t)(((b_base + (wide ? st : (uint32_t)kb) * (uint32_t)b_kb_bytes) >> 4) & 0x3FFFu);
                                const int ksteps = kb == NKB - 1 ? last_ksteps : TC_BK / 16;
                                tc_mma_f16(tmem_d, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
                                for (int k = 1; k < ksteps; ++k) {
                                    adesc += 2;  // 32 bytes (16 fp16 along K) inside the swizzle atom
                                    bdesc += 2;
                                    tc_mma_f16(tmem_d, adesc, bdesc, idesc, 1u);
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < KBB; ++j) {
                            if (j < nkb) {
                                tc_commit(smem_u32(&bars->empty_a[(g + (uint32_t)j) & 3u]));  // frees the A stage when the batch retires
                                if (last_qb && !wide) tc_commit(smem_u32(&bars->empty_b[kb0 + j]));  // last reader of this B slot
                            }
                        }
                        if (kb0 + nkb == NKB) tc_commit(smem_u32(&bars->tmem_full[buf]));
                    }
                    __syncwarp();
                    // the next batch's stages, probed while the pipe works on this one
                    mbar_try_wait2(smem_u32(&bars->full_a[gn & 3u]), (gn >> 2) & 1u, smem_u32(&bars->full_a[(gn + 1u) & 3u]),
                                   ((gn + 1u) >> 2) & 1u, ready, ready1);
                    g = gn;
                }
                stage = g & 3u;
                phase = (g >> 2) & 1u;
            } else {
            for (int kb0 = 0; kb0 < num_kb; kb0 += kbb) {
                const int nkb = min(kbb, num_kb - kb0);
                bool ready_next;
                {
                    uint32_t st = stage, ph = phase;
                    for (int j = 0; j < nkb; ++j) {
                        if (wide) mbar_wait_prof(smem_u32(&bars->full_b[st]), ph, prof_on, pw1);
                        else if (first_qb) mbar_wait_prof(smem_u32(&bars->full_b[kb0 + j]), n_local & 1, prof_on, pw1);
                        if (j > 0 || !ready) mbar_wait_prof(smem_u32(&bars->full_a[st]), ph, prof_on, pw2);
                        if (++st == (uint32_t)n_stages) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                    // probe the next batch's first stage now; the answer is needed one batch later
                    ready_next = mbar_try_wait(smem_u32(&bars->full_a[st]), ph);
                }
                tc_fence_after();
                if (elect_one()) {
                    uint32_t st = stage;
                    for (int j = 0; j < nkb; ++j) {
                        const int kb = kb0 + j;
                        uint64_t adesc = desc_hi | (uint64_t)(((a_base + st * TC_A_BYTES) >> 4) & 0x3FFFu);
                        uint64_t bdesc = desc_hi | (uint64_t)(((b_base + (wide ? st : (uint32_t)kb) * (uint32_t)b_kb_bytes) >> 4) & 0x3FFFu);
                        const int ksteps = kb == num_kb - 1 ? last_ksteps : TC_BK / 16;
                        tc_mma_f16(tmem_d, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
                        for (int k = 1; k < ksteps; ++k) {
                            adesc += 2;
                            bdesc += 2;
                            tc_mma_f16(tmem_d, adesc, bdesc, idesc, 1u);
                        }
                        if (++st == (uint32_t)n_stages) st = 0;
                    }
                    st = stage;
                    for (int j = 0; j < nkb; ++j) {
                        tc_commit(smem_u32(&bars->empty_a[st]));
                        if (last_qb && !wide) tc_commit(smem_u32(&bars->empty_b[kb0 + j]));
                        if (++st == (uint32_t)n_stages) st = 0;
                    }
                    if (kb0 + nkb == num_kb) tc_commit(smem_u32(&bars->tmem_full[buf]));
                }
                __syncwarp();
                for (int j = 0; j < nkb; ++j)
                    if (++stage == (uint32_t)n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                ready = ready_next;
            }
            g = 0;  // unused on this path
            }
            if (last_qb) ++n_local;
            ++unit;
            tc.advance();
        }
        if (prof_on && lane == 0) {
            a.prof[blockIdx.x * 8 + 2] = pw0;  // MMA: wait tmem_empty
            a.prof[blockIdx.x * 8 + 3] = pw1;  // MMA: wait full_b
            a.prof[blockIdx.x * 8 + 4] = pw2;  // MMA: wait full_a
        }
        if (prof_on) {  // the elected lane's issue-side breakdown (debug: reuses the TMA / producer slots)
            pq0 = __reduce_max_sync(0xffffffffu, (unsigned)(pq0 >> 10));
            pq1 = __reduce_max_sync(0xffffffffu, (unsigned)(pq1 >> 10));
            pq2 = __reduce_max_sync(0xffffffffu, (unsigned)(pq2 >> 10));
            if (lane == 0 && (a.debug & 64)) {
                a.prof[blockIdx.x * 8 + 1] = pq0 << 10;
                a.prof[blockIdx.x * 8 + 5] = pq1 << 10;
                a.prof[blockIdx.x * 8 + 6] = pq2 << 10;
            }
        }
    } else if (FLAGS & 4) {
        // ================= A by TMA: 128 consecutive query rows per k-block =================
        // (rows behind the last query and the columns behind `dim` are zero-filled by the out-of-bounds rule)
        if (warp == TC_EPI_WARPS + 2) {
            uint32_t stage = 0, phase = 0;
            const uint32_t a_base = smem_u32(sA);
            TileCursor tc;
            tc.init(a.items, n_items, blockIdx.x, gridDim.x);
            while (tc.valid) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_prof(smem_u32(&bars->empty_a[stage]), phase ^ 1u, prof_on, pw0);
                    if (elect_one()) {
                        const uint32_t fa = smem_u32(&bars->full_a[stage]);
                        mbar_expect_tx(fa, (uint32_t)TC_A_BYTES);
                        tma_load_2d(a_base + stage * TC_A_BYTES, &tmap_q, kb * TC_BK, tc.cur.g0 + tc.qb * TC_BM, fa);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc.advance();
            }
            if (prof_on && lane == 0) a.prof[blockIdx.x * 8 + 5] = pw0;
        }
    } else {
        // ================= A producers: gather 128 query rows per k-block =================
        // Thread p copies 16-byte chunk `chunk` of rows rbase + TC_PROD_ROWS_STEP * i; a chunk whose
        // eight fp16 values are all zero is zero-filled without touching L2. The rows' query ids and
        // non-zero masks are fetched one tile ahead.
        const int p = threadIdx.x - (TC_EPI_WARPS + 2) * 32;  // 0..TC_PRODUCERS-1
        const int chunk = p & 7;             // 16-byte chunk inside the 128-byte k-block row
        const int rbase = p >> 3;
        constexpr int RSTEP = TC_PRODUCERS / 8;  // row step between a thread's rows
        constexpr int NR = TC_BM / RSTEP;        // rows per thread
        const uint32_t dst_off = (uint32_t)(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));
        const uint32_t a_base = smem_u32(sA);
        uint32_t stage = 0, phase = 0;
        const size_t row_bytes = (size_t)a.dim * sizeof(__half);
        const unsigned char *qh_c = reinterpret_cast<const unsigned char *>(a.qh) + chunk * 16;
        TileCursor tc;
        tc.init(a.items, n_items, blockIdx.x, gridDim.x);
        const unsigned char *src[NR];
        uint32_t nz[NR];
        int qn[NR];
        if (tc.valid) {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int gi = rbase + RSTEP * i;
                const bool real = gi < tc.cur.G;
                const int q = a.gq ? a.gq[tc.cur.g0 + (real ? gi : 0)] : tc.cur.g0 + (real ? gi : 0);
                src[i] = qh_c + (size_t)q * row_bytes;
                nz[i] = real ? a.qmask[(size_t)q * 8 + chunk] : 0u;  // padding rows are zero-filled
            }
        }
        while (tc.valid) {
            TileCursor nx = tc;
            nx.advance();
            const unsigned char *src_n[NR];
            uint32_t nz_n[NR];
            // row groups (RSTEP rows each) that hold at least one query of this tile
            const int n_rowgroups = (a.debug & 32) ? NR : (min(tc.cur.G - tc.qb * TC_BM, TC_BM) + RSTEP - 1) / RSTEP;
            if (!(FLAGS & 1)) {  // no prefetch: fetch this tile's rows now
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int gi = tc.qb * TC_BM + rbase + RSTEP * i;
                    const bool real = gi < tc.cur.G;
                    const int q = a.gq ? a.gq[tc.cur.g0 + (real ? gi : 0)] : tc.cur.g0 + (real ? gi : 0);
                    src[i] = qh_c + (size_t)q * row_bytes;
                    nz[i] = real ? a.qmask[(size_t)q * 8 + chunk] : 0u;
                }
            }
            auto produce = [&](int kb) {
                if ((FLAGS & 1) && kb == 0 && nx.valid) {  // next tile's query ids
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const int gi = nx.qb * TC_BM + rbase + RSTEP * i;
                        const bool real = gi < nx.cur.G;
                        const int q = a.gq ? a.gq[nx.cur.g0 + (real ? gi : 0)] : nx.cur.g0 + (real ? gi : 0);
                        qn[i] = real ? q : -1 - q;  // negative: padding row (zero-filled)
                    }
                }
                if ((FLAGS & 1) && kb == num_kb / 2 && nx.valid) {  // ... and its masks / row pointers
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const bool real = qn[i] >= 0;
                        const int q = real ? qn[i] : -1 - qn[i];
                        src_n[i] = qh_c + (size_t)q * row_bytes;
                        nz_n[i] = real ? a.qmask[(size_t)q * 8 + chunk] : 0u;
                    }
                }
                mbar_wait_prof(smem_u32(&bars->empty_a[stage]), phase ^ 1u, prof_on, pw0);
                const uint32_t dst0 = a_base + stage * TC_A_BYTES + dst_off;
                if (!(a.debug & 1)) {
                    // only the row groups that hold queries of this tile are written: the rows behind the
                    // group's end keep whatever the stage held (an accumulator row depends on its own A row
                    // only, and the epilogue never looks at those rows), which saves the gather's LSU slots —
                    // the resource this kernel is bound by — for short query groups (round 0, tail tiles)
#pragma unroll
                    for (int i = 0; i < NR; ++i)
                        if (i < n_rowgroups)
                            cp_async16_zfill(dst0 + i * RSTEP * 128, src[i] + (size_t)kb * 128, (nz[i] >> kb) & 1u ? 16u : 0u);
                }
                cp_async_arrive_noinc(smem_u32(&bars->full_a[stage]));
                if (++stage == (uint32_t)n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            };
            if (NUM_KB > 0) {
#pragma unroll K3_UP
                for (int kb = 0; kb < NUM_KB; ++kb) produce(kb);
            } else {
                for (int kb = 0; kb < num_kb; ++kb) produce(kb);
            }
            if (FLAGS & 1) {
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    src[i] = src_n[i];
                    nz[i] = nz_n[i];
                }
            }
            tc = nx;
        }
        cp_async_wait<0>();
        if (prof_on && p == 0) {
            a.prof[blockIdx.x * 8 + 5] = pw0;  // A producer: wait empty_a
            a.prof[blockIdx.x * 8 + 6] = pw1;  // A producer: wait cp.async data
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}


// ======================================================================= CTA-pair variant (cta_group::2)
//
// Two CTAs of a cluster (one TPC) share every work item: the pair's tile is 256 queries (each CTA
// gathers and drains its own 128) and the list chunk is split between them (CTA r keeps rows
// [r N/2, (r+1) N/2) resident), so the resident chunk costs half the shared memory per CTA and the
// ring of query stages — what hides the gather's L2 latency — roughly doubles. Only the leader
// (rank 0) issues tcgen05.mma.cta_group::2 (M = 256); its commits are multicast to both CTAs'
// barriers. Cross-CTA signals: both CTAs' TMA loads complete on the leader's full_b; the peer's
// MMA warp relays "my query stage is full" to the leader's full_peer; the peer's epilogue warps
// arrive on the leader's tmem_empty.

struct __align__(8) Tc2Barriers {
    unsigned long long full_a[TC_MAX_STAGES];
    unsigned long long full_peer[TC_MAX_STAGES];
    unsigned long long empty_a[TC_MAX_STAGES];
    unsigned long long full_b[TC_MAX_KB];
    unsigned long long empty_b[TC_MAX_KB];
    unsigned long long tmem_full[2];
    unsigned long long tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {  // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int TC2_BOX = 16;  // rows per TMA box of the pair kernel's tensor map

template <int NUM_KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
scan_tc2_kernel(const __grid_constant__ CUtensorMap tmap_vec, TcScanArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char *tc_smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    const int num_kb = NUM_KB > 0 ? NUM_KB : (a.dim + TC_BK - 1) / TC_BK;
    const int half_nb = a.nb / 2;             // list rows resident in this CTA
    const int b_kb_bytes = half_nb * 128;
    const int n_stages = a.stages;
    unsigned char *sA = tc_smem;
    unsigned char *sB = tc_smem + n_stages * TC_A_BYTES;
    Tc2Barriers *bars = reinterpret_cast<Tc2Barriers *>(sB + (size_t)num_kb * b_kb_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_items = (int)a.item_off[a.nlist];
    const int first_item = blockIdx.x >> 1, item_step = gridDim.x >> 1;
    constexpr int BM2 = 2 * TC_BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) {
            mbar_init(smem_u32(&bars->full_a[s]), TC_PRODUCERS);
            mbar_init(smem_u32(&bars->full_peer[s]), 1);
            mbar_init(smem_u32(&bars->empty_a[s]), 1);
        }
        for (int kb = 0; kb < TC_MAX_KB; ++kb) {
            mbar_init(smem_u32(&bars->full_b[kb]), 1);
            mbar_init(smem_u32(&bars->empty_b[kb]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->tmem_full[b]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[b]), 8);  // four epilogue warps of each CTA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();  // barriers of both CTAs exist before any remote signal
    if (warp == TC_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < TC_EPI_WARPS) {
        // ================= epilogue (this CTA's 128 query rows of every 256-query tile) =================
        const int set = warp >> 2;
        uint32_t unit = (uint32_t)set;
        const int row = (int)rank * TC_BM + (warp & 3) * 32 + lane;  // row inside the 256-query tile
        const float scale = 1.f / a.inv_scale;
        TileCursor tc;
        tc.init(a.items, n_items, first_item, item_step, BM2);
        if (set == 1 && tc.valid) tc.advance();
        int q = -1;
        float thr = INFINITY;
        if (tc.valid) {
            const int gi = tc.qb * BM2 + row;
            q = gi < tc.cur.G ? a.gq[tc.cur.g0 + gi] : -1;
            if (q >= 0) thr = a.tau[q] * scale;
        }
        const uint32_t leader_empty0 = map_to_cta(smem_u32(&bars->tmem_empty[0]), 0);
        const uint32_t leader_empty1 = map_to_cta(smem_u32(&bars->tmem_empty[1]), 0);
        while (tc.valid) {
            const TcItem it = tc.cur;
            TileCursor nx = tc;
            nx.advance();
            if (nx.valid) nx.advance();
            int q_next = -1;
            if (nx.valid) {
                const int gi = nx.qb * BM2 + row;
                q_next = gi < nx.cur.G ? a.gq[nx.cur.g0 + gi] : -1;
            }
            const int buf = set;
            mbar_wait(smem_u32(&bars->tmem_full[buf]), (unit >> 1) & 1);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * 256);
            for (int c0 = 0; c0 < it.nv; c0 += 32) {
                uint32_t r[32];
                tc_ld32(tbase + (uint32_t)c0, r);
                tc_wait_ld();
                if (c0 + 32 >= it.nv) {  // last chunk of this accumulator: hand the buffer back to the leader's MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(buf ? leader_empty1 : leader_empty0);
                }
                if (q < 0) continue;
                uint32_t mask = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (__uint_as_float(r[j]) >= thr) mask |= 1u << j;
                const int left = it.nv - c0;
                if (left < 32) mask &= (1u << left) - 1u;
                if (mask) {
                    int slot = atomicAdd(&a.cnt[q], __popc(mask));
                    unsigned long long *qbuf = a.buf + (int64_t)q * a.cap;
                    const uint32_t pbase = (uint32_t)(it.p0 + c0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if ((mask >> j) & 1u) {
                            if (slot < a.cap)
                                qbuf[slot] = ((unsigned long long)__float_as_uint(__uint_as_float(r[j]) * a.inv_scale) << 32) |
                                             (unsigned long long)(pbase + j);
                            ++slot;
                        }
                    }
                }
            }
            q = q_next;
            thr = q >= 0 ? a.tau[q] * scale : INFINITY;
            tc = nx;
            unit += 2;
        }
    } else if (warp == TC_EPI_WARPS) {
        // ================= TMA: this CTA's half of the item's list chunk =================
        uint32_t n_local = 0;
        TcItem it, nxt;
        int item = first_item;
        if (item < n_items) it = a.items[item];
        for (; item < n_items; item += item_step, ++n_local) {
            if (item + item_step < n_items) nxt = a.items[item + item_step];
            const int npad = (it.nv + 15) & ~15;         // MMA N
            const int half = npad >> 1;                   // rows per CTA
            const int nbox = (half + TC2_BOX - 1) / TC2_BOX;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(smem_u32(&bars->empty_b[kb]), (n_local & 1) ^ 1);
                if (elect_one()) {
                    const uint32_t fb = map_to_cta(smem_u32(&bars->full_b[kb]), 0);  // the leader's barrier
                    if (leader) mbar_expect_tx(smem_u32(&bars->full_b[kb]), (uint32_t)(2 * nbox * TC2_BOX * 128));
                    for (int j = 0; j < nbox; ++j)
                        tma_load_2d_pair(smem_u32(sB + (size_t)kb * b_kb_bytes + j * TC2_BOX * 128), &tmap_vec, kb * TC_BK,
                                         it.p0 + (int)rank * half + j * TC2_BOX, fb);
                }
                __syncwarp();
            }
            it = nxt;
        }
    } else if (warp == TC_EPI_WARPS + 1) {
        uint32_t stage = 0, phase = 0, unit = 0, n_local = 0;
        TileCursor tc;
        tc.init(a.items, n_items, first_item, item_step, BM2);
        if (leader) {
            // ================= MMA issuer (leader CTA) =================
            const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            const int last_ksteps = (a.dim - (num_kb - 1) * TC_BK) / 16;
            while (tc.valid) {
                const int npad = (tc.cur.nv + 15) & ~15;
                // instruction descriptor: D=f32, A=B=f16, K-major, M=256 (pair), N=npad
                const uint32_t idesc = (1u << 4) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(BM2 >> 4) << 24);
                const int buf = unit & 1;
                mbar_wait(smem_u32(&bars->tmem_empty[buf]), ((unit >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 256);
                const bool first_qb = tc.first_qb(), last_qb = tc.last_qb();
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (first_qb) mbar_wait(smem_u32(&bars->full_b[kb]), n_local & 1);
                    mbar_wait(smem_u32(&bars->full_a[stage]), phase);
                    mbar_wait(smem_u32(&bars->full_peer[stage]), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        uint64_t adesc = desc_hi | (uint64_t)(((a_base + stage * TC_A_BYTES) >> 4) & 0x3FFFu);
                        uint64_t bdesc = desc_hi | (uint64_t)(((b_base + (uint32_t)kb * (uint32_t)b_kb_bytes) >> 4) & 0x3FFFu);
                        const int ksteps = kb == num_kb - 1 ? last_ksteps : TC_BK / 16;
                        tc_mma_f16_pair(tmem_d, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
                        for (int k = 1; k < ksteps; ++k) {
                            adesc += 2;
                            bdesc += 2;
                            tc_mma_f16_pair(tmem_d, adesc, bdesc, idesc, 1u);
                        }
                        tc_commit_pair(smem_u32(&bars->empty_a[stage]));
                        if (last_qb) tc_commit_pair(smem_u32(&bars->empty_b[kb]));
                        if (kb == num_kb - 1) tc_commit_pair(smem_u32(&bars->tmem_full[buf]));
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (last_qb) ++n_local;
                ++unit;
                tc.advance();
            }
        } else {
            // ================= relay (peer CTA): "my query stage is full" -> the leader's full_peer =================
            while (tc.valid) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(smem_u32(&bars->full_a[stage]), phase);
                    if (elect_one()) mbar_arrive_cluster(map_to_cta(smem_u32(&bars->full_peer[stage]), 0));
                    __syncwarp();
                    if (++stage == (uint32_t)n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc.advance();
            }
        }
    } else {
        // ================= A producers: this CTA's 128 query rows per k-block =================
        const int p = threadIdx.x - (TC_EPI_WARPS + 2) * 32;
        const int chunk = p & 7;
        const int rbase = p >> 3;
        constexpr int RSTEP = TC_PRODUCERS / 8;
        constexpr int NR = TC_BM / RSTEP;
        const uint32_t dst_off = (uint32_t)(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));
        const uint32_t a_base = smem_u32(sA);
        uint32_t stage = 0, phase = 0;
        const size_t row_bytes = (size_t)a.dim * sizeof(__half);
        const unsigned char *qh_c = reinterpret_cast<const unsigned char *>(a.qh) + chunk * 16;
        const int row0 = (int)rank * TC_BM + rbase;
        TileCursor tc;
        tc.init(a.items, n_items, first_item, item_step, BM2);
        const unsigned char *src[NR];
        uint32_t nz[NR];
        int qn[NR];
        if (tc.valid) {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int gi = row0 + RSTEP * i;
                const bool real = gi < tc.cur.G;
                const int q = a.gq[tc.cur.g0 + (real ? gi : 0)];
                src[i] = qh_c + (size_t)q * row_bytes;
                nz[i] = real ? a.qmask[(size_t)q * 8 + chunk] : 0u;
            }
        }
        while (tc.valid) {
            TileCursor nx = tc;
            nx.advance();
            const unsigned char *src_n[NR];
            uint32_t nz_n[NR];
            auto produce = [&](int kb) {
                if (kb == 0 && nx.valid) {
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const int gi = nx.qb * BM2 + row0 + RSTEP * i;
                        const bool real = gi < nx.cur.G;
                        const int q = a.gq[nx.cur.g0 + (real ? gi : 0)];
                        qn[i] = real ? q : -1 - q;
                    }
                }
                if (kb == num_kb / 2 && nx.valid) {
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const bool real = qn[i] >= 0;
                        const int q = real ? qn[i] : -1 - qn[i];
                        src_n[i] = qh_c + (size_t)q * row_bytes;
                        nz_n[i] = real ? a.qmask[(size_t)q * 8 + chunk] : 0u;
                    }
                }
                mbar_wait(smem_u32(&bars->empty_a[stage]), phase ^ 1u);
                const uint32_t dst0 = a_base + stage * TC_A_BYTES + dst_off;
#pragma unroll
                for (int i = 0; i < NR; ++i)
                    cp_async16_zfill(dst0 + i * RSTEP * 128, src[i] + (size_t)kb * 128, (nz[i] >> kb) & 1u ? 16u : 0u);
                cp_async_arrive_noinc(smem_u32(&bars->full_a[stage]));
                if (++stage == (uint32_t)n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            };
            if (NUM_KB > 0) {
#pragma unroll
                for (int kb = 0; kb < NUM_KB; ++kb) produce(kb);
            } else {
                for (int kb = 0; kb < num_kb; ++kb) produce(kb);
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                src[i] = src_n[i];
                nz[i] = nz_n[i];
            }
            tc = nx;
        }
        cp_async_wait<0>();
    }

    tc_fence_before();
    cluster_sync_all();  // the peer's shared memory and barriers stay alive until the leader's last MMA retired
    if (warp == TC_EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

// only lists with len_lo < length <= len_hi take part (the scan may split the lists between two kernel variants)
// ======================================================================= swapped operands (scan_ts_kernel)
//
// The list chunk is the A operand and lives in TENSOR MEMORY, the gathered query rows are the B operand.
//   work item = chunk of <= 128 consecutive vectors of a list (one TMEM lane per vector, fp16 pairs packed into
//               dim / 2 columns: 400 of the 512 columns at dim 800), written once per item by the epilogue warps
//               (TMA box -> staging -> ld.shared -> tcgen05.st);
//   tile      = NQ queries of the list's query group: NQ gathered fp16 rows per k-block are the B operand
//               (N = NQ), streamed through a ring that has ALL of shared memory because no list chunk is resident
//               there (14 stages of 12 KB instead of 4 of 16 KB: the gather's L2 latency, which bounds
//               scan_tc_kernel's two-batch ring, is covered);
//   D         = 128 lanes (list vectors) x NQ fp32 columns (queries), one accumulator in front of the A columns.
// Shared-memory traffic per k-block drops from 44 KB (16 KB written + 16 + 12 KB read by the MMAs) to 24 KB.
// Epilogue: thread = list vector, column = query; per column one ballot over the warp's 32 vectors, one atomic
// reservation per (warp, query) with a hit and coalesced 8-byte stores.
// Hand-over unit = TS_BATCH k-blocks: one full / empty barrier pair and one tcgen05.commit per batch.
constexpr int TS_MT = 128;            // list vectors per item (MMA M)
constexpr int TS_BATCH = 2;           // k-blocks per hand-over
constexpr int TS_MAX_SLOTS = 8;       // batch slots in the ring (stages = slots * TS_BATCH)
constexpr int TS_RING = 3;            // staging boxes (TC_BOX list rows x 128 B) per loader warp
constexpr int TS_SBOX = TS_RING * (TS_MT / TC_BOX);
constexpr int TS_BOX_BYTES = TC_BOX * 128;

struct __align__(8) TsBarriers {
    unsigned long long full_q[TS_MAX_SLOTS];
    unsigned long long empty_q[TS_MAX_SLOTS];
    unsigned long long stg_full[TS_SBOX];
    unsigned long long stg_empty[TS_SBOX];
    unsigned long long a_full[TC_MAX_KB];
    unsigned long long tmem_full;
    unsigned long long tmem_empty;
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns <- 8 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint4 &v0, const uint4 &v1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v0.x),
                 "r"(v0.y), "r"(v0.z), "r"(v0.w), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w)
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

// NQ: queries per tile (MMA N; 96 or 112). Warps: 0..7 epilogue + tensor-memory loader (set = warp >> 2 owns the
// columns [set NQ/2, (set + 1) NQ/2) of the accumulator; warps 0..3 also load the list chunk into tensor memory),
// 8 TMA (staging boxes), 9 MMA issuer, 10..13 query gather.
template <int NUM_KB, int NQ, bool DENSE>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_ts_kernel(const __grid_constant__ CUtensorMap tmap_vec, TcScanArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char *tc_smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    constexpr int ST_BYTES = NQ * 128;       // one k-block of a query tile
    constexpr int HALF = NQ / 2;             // accumulator columns per epilogue warp set
    static_assert(NQ % 16 == 0 && HALF % 8 == 0 && HALF >= 32 && HALF <= 64, "unsupported tile width");
    const int num_kb = NUM_KB > 0 ? NUM_KB : (a.dim + TC_BK - 1) / TC_BK;
    const int last_ksteps = (a.dim - (num_kb - 1) * TC_BK) / 16;
    const int n_slots = a.stages / TS_BATCH;
    unsigned char *sQ = tc_smem;
    unsigned char *sStg = tc_smem + (size_t)a.stages * ST_BYTES;
    TsBarriers *bars = reinterpret_cast<TsBarriers *>(sStg + TS_SBOX * TS_BOX_BYTES);
    int32_t *s_qid = reinterpret_cast<int32_t *>(bars + 1);   // [8 warps][64]
    float *s_thr = reinterpret_cast<float *>(s_qid + 8 * 64);  // [8 warps][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = (int)a.item_off[a.nlist];

    if (threadIdx.x == 0) {
        for (int s = 0; s < TS_MAX_SLOTS; ++s) {
            mbar_init(smem_u32(&bars->full_q[s]), TC_PRODUCERS);
            mbar_init(smem_u32(&bars->empty_q[s]), 1);
        }
        for (int s = 0; s < TS_SBOX; ++s) {
            mbar_init(smem_u32(&bars->stg_full[s]), 1);
            mbar_init(smem_u32(&bars->stg_empty[s]), 1);
        }
        for (int kb = 0; kb < TC_MAX_KB; ++kb) mbar_init(smem_u32(&bars->a_full[kb]), 4);
        mbar_init(smem_u32(&bars->tmem_full), 1);
        mbar_init(smem_u32(&bars->tmem_empty), TC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_EPI_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t tmem_a = tmem_base + (uint32_t)a.a_col0;

    if (warp < TC_EPI_WARPS) {
        // ================= epilogue + tensor-memory loader =================
        const int set = warp >> 2, w4 = warp & 3;
        const int row = w4 * 32 + lane;                       // list vector of this thread = TMEM lane
        const uint32_t lane_addr = (uint32_t)(w4 * 32) << 16;
        const float scale = 1.f / a.inv_scale;
        int32_t *my_q = s_qid + warp * 64;
        float *my_thr = s_thr + warp * 64;
        uint32_t tile_n = 0, box_seq = 0;   // box_seq: staging boxes this warp has consumed
        TileCursor tc;
        tc.init(a.items, n_items, blockIdx.x, gridDim.x, NQ);
        // query ids and (pre-scaled) thresholds of my columns live in shared memory (my_q / my_thr); lane l fetches
        // those of columns l and 32 + l of the NEXT tile while the current one is handled (qn / tn)
        const float nan_thr = __int_as_float(0x7fc00000);   // behind the group's end: nothing passes
        int qn[2] = {-1, -1};
        float tn[2] = {nan_thr, nan_thr};
        auto fetch_q = [&](const TileCursor &c) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int j = hh * 32 + lane;
                const int gi = c.qb * NQ + set * HALF + j;
                qn[hh] = (j < HALF && gi < c.cur.G) ? a.gq[c.cur.g0 + gi] : -1;
            }
        };
        auto fetch_thr = [&]() {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) tn[hh] = qn[hh] >= 0 ? a.tau[qn[hh]] * scale : nan_thr;
        };
        auto publish = [&]() {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
                if (hh * 32 + lane < HALF) {
                    my_q[hh * 32 + lane] = qn[hh];
                    my_thr[hh * 32 + lane] = tn[hh];
                }
            __syncwarp();
        };
        if (tc.valid) {
            fetch_q(tc);
            fetch_thr();
            publish();
        }
        while (tc.valid) {
            const TcItem it = tc.cur;
            TileCursor nx = tc;
            nx.advance();
            if (set == 0 && tc.first_qb()) {
                // the list chunk -> tensor memory (warps 0..3, one lane quarter each); every MMA of the previous item
                // has retired (this warp passed the tmem_full wait of its last tile). Each warp has a private ring of
                // TS_RING staging boxes: one producer and one consumer per barrier, always in step (a ring shared by
                // warps that drift apart lets a fast warp's parity wait match the PREVIOUS phase of a slot).
                const int nbox = (it.nv + TC_BOX - 1) / TC_BOX;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (w4 < nbox) {
                        const uint32_t slot = (uint32_t)w4 * TS_RING + box_seq % TS_RING, ph = (box_seq / TS_RING) & 1u;
                        ++box_seq;
                        mbar_wait(smem_u32(&bars->stg_full[slot]), ph);
                        const unsigned char *rowp = sStg + slot * TS_BOX_BYTES + lane * 128;
                        const int ksteps = kb == num_kb - 1 ? last_ksteps : TC_BK / 16;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint4 v0 = *reinterpret_cast<const uint4 *>(rowp + (((2 * ks) ^ (lane & 7)) << 4));
                            const uint4 v1 = *reinterpret_cast<const uint4 *>(rowp + (((2 * ks + 1) ^ (lane & 7)) << 4));
                            tc_st8(tmem_a + lane_addr + (uint32_t)(kb * 32 + ks * 8), v0, v1);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars->stg_empty[slot]));
                    }
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars->a_full[kb]));
                }
            }
            mbar_wait(smem_u32(&bars->tmem_full), tile_n & 1u);
            tc_fence_after();
            const bool rows_here = w4 * 32 < it.nv;
            uint32_t r[HALF];
            if (rows_here) {
                const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(set * HALF);
                tc_ld32(taddr, r);
                if constexpr (HALF >= 48) tc_ld16(taddr + 32, r + 32);
                if constexpr (HALF == 40) tc_ld8(taddr + 32, r + 32);
                if constexpr (HALF == 56) tc_ld8(taddr + 48, r + 48);
                if constexpr (HALF == 64) tc_ld16(taddr + 48, r + 48);
                tc_wait_ld();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty));
            if (nx.valid) fetch_q(nx);   // in flight during pass 1
            const bool valid = row < it.nv;
            const bool work = rows_here && !(a.debug & 2);
            const unsigned long long pos = (unsigned long long)(uint32_t)(it.p0 + row);
            if (DENSE) {
                // round 0 (thresholds at -inf: every score is appended): three passes so that the buffer reservations of
                // a tile are TWO warp-wide atomics (one lane per column): (1) hits per column, (2) reservations, (3) stores
                int hits[2] = {0, 0};        // lane l: hits of columns l and 32 + l
                if (work) {
#pragma unroll
                    for (int j = 0; j < HALF; ++j) {
                        const bool pass = valid && __uint_as_float(r[j]) >= my_thr[j];   // a NaN score never passes
                        const int n = __popc(__ballot_sync(0xffffffffu, pass));
                        if (lane == (j & 31)) hits[j >> 5] = n;
                    }
                }
                int base[2] = {0, 0};
                if (hits[0] > 0) base[0] = atomicAdd(&a.cnt[my_q[lane]], hits[0]);
                if (hits[1] > 0) base[1] = atomicAdd(&a.cnt[my_q[32 + lane]], hits[1]);
                if (nx.valid) fetch_thr();   // in flight during the reservations and pass 3
                if (work && __any_sync(0xffffffffu, (hits[0] | hits[1]) != 0)) {
#pragma unroll
                    for (int j = 0; j < HALF; ++j) {
                        const float v = __uint_as_float(r[j]);
                        const bool pass = valid && v >= my_thr[j];
                        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                        const int slot = __shfl_sync(0xffffffffu, base[j >> 5], j & 31);
                        if (pass) {
                            const int p = slot + __popc(bal & ((1u << lane) - 1u));
                            if (p < a.cap)
                                a.buf[(int64_t)my_q[j] * a.cap + p] =
                                    ((unsigned long long)__float_as_uint(v * a.inv_scale) << 32) | pos;
                        }
                    }
                }
            } else {
                // later rounds: a few per cent of the scores pass; each thread handles its own hits (no warp collectives:
                // a column without a hit costs a compare and a branch)
                if (nx.valid) fetch_thr();
                if (work && valid) {
#pragma unroll
                    for (int j = 0; j < HALF; ++j) {
                        const float v = __uint_as_float(r[j]);
                        if (v >= my_thr[j]) {   // NaN threshold behind the group's end; a NaN score never passes
                            const int q = my_q[j];
                            const int p = atomicAdd(&a.cnt[q], 1);
                            if (p < a.cap)
                                a.buf[(int64_t)q * a.cap + p] = ((unsigned long long)__float_as_uint(v * a.inv_scale) << 32) | pos;
                        }
                    }
                }
            }
            __syncwarp();   // every lane is done with my_q / my_thr of this tile
            if (nx.valid) publish();
            ++tile_n;
            tc = nx;
        }
    } else if (warp == TC_EPI_WARPS) {
        // ================= TMA: the item's list chunk, box by box, into the staging ring =================
        uint32_t sq[TS_MT / TC_BOX] = {0, 0, 0, 0};   // boxes issued so far per loader warp
        TcItem it, nxt;
        int item = blockIdx.x;
        if (item < n_items) it = a.items[item];
        for (; item < n_items; item += gridDim.x) {
            if (item + (int)gridDim.x < n_items) nxt = a.items[item + gridDim.x];
            const int nbox = (it.nv + TC_BOX - 1) / TC_BOX;
            for (int kb = 0; kb < num_kb; ++kb) {
#pragma unroll
                for (int b = 0; b < TS_MT / TC_BOX; ++b) {
                    if (b < nbox) {
                        const uint32_t slot = (uint32_t)b * TS_RING + sq[b] % TS_RING, ph = (sq[b] / TS_RING) & 1u;
                        ++sq[b];
                        mbar_wait(smem_u32(&bars->stg_empty[slot]), ph ^ 1u);
                        if (elect_one()) {
                            const uint32_t fb = smem_u32(&bars->stg_full[slot]);
                            mbar_expect_tx(fb, (uint32_t)TS_BOX_BYTES);
                            tma_load_2d(smem_u32(sStg + slot * TS_BOX_BYTES), &tmap_vec, kb * TC_BK, it.p0 + b * TC_BOX, fb);
                        }
                        __syncwarp();
                    }
                }
            }
            it = nxt;
        }
    } else if (warp == TC_EPI_WARPS + 1) {
        // ================= MMA issuer =================
        const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t q_base = smem_u32(sQ);
        const uint32_t idesc = make_idesc_f16(NQ);
        uint32_t slot = 0, phase = 0, tile_n = 0, item_n = 0;
        TileCursor tc;
        tc.init(a.items, n_items, blockIdx.x, gridDim.x, NQ);
        while (tc.valid) {
            mbar_wait(smem_u32(&bars->tmem_empty), (tile_n & 1u) ^ 1u);
            tc_fence_after();
            const bool first_qb = tc.first_qb();
            for (int kb0 = 0; kb0 < num_kb; kb0 += TS_BATCH) {
                const int n = min(TS_BATCH, num_kb - kb0);
                if (first_qb)
                    for (int j = 0; j < n; ++j) mbar_wait(smem_u32(&bars->a_full[kb0 + j]), item_n & 1u);
                mbar_wait(smem_u32(&bars->full_q[slot]), phase);
                tc_fence_after();
                if (elect_one()) {
                    if (!(a.debug & 16)) {
#pragma unroll
                        for (int j = 0; j < TS_BATCH; ++j) {
                            if (j < n) {
                                const int kb = kb0 + j;
                                uint64_t bdesc = desc_hi | (uint64_t)(((q_base + (slot * TS_BATCH + j) * ST_BYTES) >> 4) & 0x3FFFu);
                                const int ksteps = kb == num_kb - 1 ? last_ksteps : TC_BK / 16;
                                for (int ks = 0; ks < ksteps; ++ks) {
                                    tc_mma_f16_ts(tmem_base, tmem_a + (uint32_t)(kb * 32 + ks * 8), bdesc, idesc, (kb | ks) != 0 ? 1u : 0u);
                                    bdesc += 2;
                                }
                            }
                        }
                    }
                    tc_commit(smem_u32(&bars->empty_q[slot]));
                    if (kb0 + n == num_kb) tc_commit(smem_u32(&bars->tmem_full));
                }
                __syncwarp();
                if (++slot == (uint32_t)n_slots) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
            if (tc.last_qb()) ++item_n;
            ++tile_n;
            tc.advance();
        }
    } else {
        // ================= query gather: NQ rows x 128 B per k-block, TS_BATCH k-blocks per hand-over =================
        const int p = threadIdx.x - (TC_EPI_WARPS + 2) * 32;
        const int chunk = p & 7, rbase = p >> 3;
        constexpr int RSTEP = TC_PRODUCERS / 8;   // 16
        constexpr int NR = NQ / RSTEP;            // rows per thread
        const uint32_t dst_off = (uint32_t)(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));
        const uint32_t q_base = smem_u32(sQ);
        const size_t row_bytes = (size_t)a.dim * sizeof(__half);
        const unsigned char *qh_c = reinterpret_cast<const unsigned char *>(a.qh) + chunk * 16;
        uint32_t slot = 0, phase = 0;
        TileCursor tc;
        tc.init(a.items, n_items, blockIdx.x, gridDim.x, NQ);
        const unsigned char *src[NR];
        uint32_t nz[NR];
        auto fetch_rows = [&](const TileCursor &c, const unsigned char *(&s)[NR], uint32_t (&m)[NR]) {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int gi = c.qb * NQ + rbase + RSTEP * i;
                const bool real = gi < c.cur.G;
                const int q = a.gq[c.cur.g0 + (real ? gi : 0)];
                s[i] = qh_c + (size_t)q * row_bytes;
                m[i] = real ? a.qmask[(size_t)q * 8 + chunk] : 0u;   // padding rows are zero-filled
            }
        };
        if (tc.valid) fetch_rows(tc, src, nz);
        while (tc.valid) {
            TileCursor nx = tc;
            nx.advance();
            const unsigned char *src_n[NR];
            uint32_t nz_n[NR];
            const int n_rowgroups = (min(tc.cur.G - tc.qb * NQ, NQ) + RSTEP - 1) / RSTEP;
            for (int kb0 = 0; kb0 < num_kb; kb0 += TS_BATCH) {
                const int n = min(TS_BATCH, num_kb - kb0);
                if (kb0 == 0 && nx.valid) fetch_rows(nx, src_n, nz_n);   // next tile's rows, one tile ahead
                mbar_wait(smem_u32(&bars->empty_q[slot]), phase ^ 1u);
                if (!(a.debug & 1)) {
#pragma unroll
                    for (int j = 0; j < TS_BATCH; ++j) {
                        if (j < n) {
                            const int kb = kb0 + j;
                            const uint32_t dst0 = q_base + (slot * TS_BATCH + j) * ST_BYTES + dst_off;
#pragma unroll
                            for (int i = 0; i < NR; ++i)
                                if (i < n_rowgroups)
                                    cp_async16_zfill(dst0 + i * RSTEP * 128, src[i] + (size_t)kb * 128, (nz[i] >> kb) & 1u ? 16u : 0u);
                        }
                    }
                }
                cp_async_arrive_noinc(smem_u32(&bars->full_q[slot]));
                if (++slot == (uint32_t)n_slots) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
            if (nx.valid) {
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    src[i] = src_n[i];
                    nz[i] = nz_n[i];
                }
            }
            tc = nx;
        }
        cp_async_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

__global__ void tc_item_count_kernel(const int64_t *__restrict__ goff, const int64_t *__restrict__ list_off, int nlist,
                                     int nb, int64_t len_lo, int64_t len_hi, int32_t *__restrict__ cnt) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    int64_t G = goff[l + 1] - goff[l];
    int64_t len = list_off[l + 1] - list_off[l];
    cnt[l] = (len > len_lo && len <= len_hi && G > 0) ? (int32_t)((len + nb - 1) / nb) : 0;
}

__global__ void tc_item_fill_kernel(const int64_t *__restrict__ goff, const int64_t *__restrict__ list_off,
                                    const int64_t *__restrict__ item_off, int nlist, int nb, TcItem *__restrict__ items) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    const int64_t o = item_off[l];
    const int n = (int)(item_off[l + 1] - o);
    const int64_t p0 = list_off[l];
    const int len = (int)(list_off[l + 1] - p0);
    for (int c = 0; c < n; ++c) {
        TcItem it;
        it.p0 = (int32_t)(p0 + (int64_t)c * nb);
        it.nv = min(nb, len - c * nb);
        it.g0 = (int32_t)goff[l];
        it.G = (int32_t)(goff[l + 1] - goff[l]);
        items[o + c] = it;
    }
}

// Items ordered by their number of query tiles, largest first (counting sort, one CTA): the persistent CTAs take
// items b, b + gridDim.x, ... so every "row" of gridDim.x consecutive sorted items costs each CTA about the same and the
// per-CTA totals end up within a tile or two of each other (list order left up to 23 % between the mean and the slowest CTA).
// cost estimate of an item in quarter tiles: the query tiles plus the chunk load (up to one tile's worth for a full
// chunk: what separates the single-tile items of the first round)
__device__ __forceinline__ int tc_item_cost(const TcItem &it, int bm, int nbin) {
    return min(nbin - 1, 4 * ((it.G + bm - 1) / bm) + (it.nv + 31) / 32);
}

__global__ void __launch_bounds__(1024) tc_item_sort_kernel(const TcItem *__restrict__ in, const int64_t *__restrict__ n_ptr,
                                                           TcItem *__restrict__ out, int bm) {
    constexpr int NBIN = 256;
    __shared__ int s_cnt[NBIN], s_off[NBIN];
    const int n = (int)*n_ptr;
    for (int i = threadIdx.x; i < NBIN; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int t = tc_item_cost(in[i], bm, NBIN);
        atomicAdd(&s_cnt[NBIN - 1 - t], 1);   // bin 0 = the most expensive items
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < NBIN; ++b) {
            s_off[b] = acc;
            acc += s_cnt[b];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const TcItem it = in[i];
        const int t = tc_item_cost(it, bm, NBIN);
        out[atomicAdd(&s_off[NBIN - 1 - t], 1)] = it;
    }
}

// fp32 queries -> scaled fp16 rows + per-(query, chunk) non-zero masks for the zero-fill gather.
// One thread per (query, 16-byte chunk position c): it walks the k-blocks.
__global__ void tc_prepare_queries_kernel(const float *__restrict__ x, int nq, int dim, float scale,
                                          __half *__restrict__ y, uint32_t *__restrict__ qmask) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nq * 8) return;
    const int q = (int)(t >> 3), c = (int)(t & 7);
    const int num_kb = (dim + TC_BK - 1) / TC_BK;
    uint32_t m = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
        const int d0 = kb * TC_BK + c * 8;
        if (d0 >= dim) break;
        const float *src = x + (int64_t)q * dim + d0;
        __align__(16) __half hv[8];
        bool any = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float v = d0 + j < dim ? src[j] * scale : 0.f;
            hv[j] = __float2half_rn(v);
            any |= (__half_as_ushort(hv[j]) & 0x7FFFu) != 0;
        }
        if (d0 + 8 <= dim) {
            *reinterpret_cast<uint4 *>(y + (int64_t)q * dim + d0) = *reinterpret_cast<const uint4 *>(hv);
        } else {
            for (int j = 0; d0 + j < dim; ++j) y[(int64_t)q * dim + d0 + j] = hv[j];
        }
        if (any) m |= 1u << kb;
    }
    qmask[t] = m;
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

static bool tc_encode_rows(void *storage, const void *base, int dim, int64_t rows, int box_rows = TC_BOX) {
    CUtensorMap *m = reinterpret_cast<CUtensorMap *>(storage);
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)dim * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SOLO_REQUIRE(r == CUDA_SUCCESS, SOLO_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return true;
}

// (re)build the TMA descriptor of the list-ordered fp16 vectors; called from ivf_finalize
void tc_make_tensor_map(IvfIndex &ix) {
    ix.tmap_valid = false;
    if (!tc_scan_supported(ix) || ix.nstored == 0) return;
    static_assert(sizeof(CUtensorMap) <= sizeof(ix.tmap_storage), "tensor map storage too small");
    ix.tmap_valid = tc_encode_rows(ix.tmap_storage, ix.vec_h.p, ix.dim, ix.nstored);
    tc_encode_rows(ix.tmap16_storage, ix.vec_h.p, ix.dim, ix.nstored, TC2_BOX);  // CTA-pair kernel: 16-row boxes
}

// the same for the fp16 centroid table; called from ivf_set_centroids
void tc_make_centroid_map(IvfIndex &ix) {
    ix.tmap_cent_valid = false;
    if (!tc_scan_supported(ix) || ix.nlist == 0) return;
    ix.tmap_cent_valid = tc_encode_rows(ix.tmap_cent_storage, ix.cent_h.p, ix.dim, ix.nlist);
}

// SOLO_TC_PROF=1 (debugging): per-role wait-cycle counters of the scan kernel, printed to stderr after
// every launch (synchronises the device around the launch)
static unsigned long long *g_tc_prof = nullptr;
static unsigned long long *tc_prof_buffer() {
    static const bool on = [] {
        const char *e = getenv("SOLO_TC_PROF");
        return e && e[0] == '1';
    }();
    if (!on) return nullptr;
    if (!g_tc_prof) cudaMalloc(&g_tc_prof, kNumSMs * 8 * sizeof(unsigned long long));
    cudaDeviceSynchronize();
    cudaMemset(g_tc_prof, 0, kNumSMs * 8 * sizeof(unsigned long long));
    cudaDeviceSynchronize();
    return g_tc_prof;
}
static void tc_prof_report(solo_handle *h, const char *what, int ctas) {
    (void)h;
    if (!g_tc_prof) return;
    cudaDeviceSynchronize();
    std::vector<unsigned long long> v(kNumSMs * 8);
    cudaMemcpy(v.data(), g_tc_prof, v.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    static const char *names[8] = {"epi:tmem_full", "tma:empty_b", "mma:tmem_empty", "mma:full_b",
                                   "mma:full_a",    "a:empty_a",   "a:cp.async",     "total"};
    fprintf(stderr, "[tc_prof %s]", what);
    for (int c = 0; c < 8; ++c) {
        double sum = 0, mx = 0;
        for (int b = 0; b < ctas; ++b) {
            sum += (double)v[b * 8 + c];
            mx = std::max(mx, (double)v[b * 8 + c]);
        }
        fprintf(stderr, " %s=%.0fk/%.0fk", names[c], sum / ctas / 1e3, mx / 1e3);
    }
    fprintf(stderr, " (mean/max kcycles per CTA)\n");
}

// Shared-memory split: the resident list chunk (NB rows x num_kb k-blocks) and the ring of query
// stages. The ring is what hides the L2 latency of the query gather, so NB is kept as small as the
// list lengths allow: the largest multiple of 32 that still leaves at least four stages, but not
// larger than needed for the longest list (SOLO_TC_NB overrides, for experiments).
static void tc_smem_plan(const IvfIndex &ix, int *nb_out, int *stages_out, int nb_cap = 256, int nb_opt = 0) {
    const int num_kb = (ix.dim + TC_BK - 1) / TC_BK;
    const int avail = TC_SMEM_MAX - 1024 - (int)sizeof(TcBarriers) - 64;
    int nb = (avail - 4 * TC_A_BYTES) / (num_kb * 128) / 32 * 32;
    nb = std::min(nb, nb_cap / 32 * 32);
    static const int env_nb = getenv("SOLO_TC_NB") ? atoi(getenv("SOLO_TC_NB")) : 0;
    if (env_nb >= 32 && env_nb <= nb) nb = env_nb / 32 * 32;
    if (nb_opt >= 32 && nb_opt <= nb) nb = nb_opt / 32 * 32;
    int stages = nb >= 32 ? (avail - num_kb * nb * 128) / TC_A_BYTES : 0;
    *nb_out = nb;
    *stages_out = std::min(stages, TC_MAX_STAGES);
}

void tc_prepare_queries(solo_handle *h, const IvfIndex &ix, const float *q, int nq, int q_scale_log2, __half *qh,
                        uint32_t *qmask) {
    const int64_t n = (int64_t)nq * 8;
    tc_prepare_queries_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(q, nq, ix.dim, ldexpf(1.f, q_scale_log2), qh, qmask);
    SOLO_CUDA(cudaGetLastError());
    h->launches++;
}

// CTA-pair plan: each CTA keeps half of the chunk resident
static void tc2_smem_plan(const IvfIndex &ix, int *nb_out, int *stages_out) {
    const int num_kb = (ix.dim + TC_BK - 1) / TC_BK;
    const int avail = TC_SMEM_MAX - 1024 - (int)sizeof(Tc2Barriers) - 64;
    static const int env_nb = getenv("SOLO_TC2_NB") ? atoi(getenv("SOLO_TC2_NB")) : 0;
    int nb = env_nb >= 32 ? env_nb / 32 * 32 : 128;  // chunk rows per pair (multiple of 32: 16-row boxes per CTA)
    while (nb > 32 && (avail - num_kb * (nb / 2) * 128) / TC_A_BYTES < 4) nb -= 32;
    *nb_out = nb;
    *stages_out = std::min((avail - num_kb * (nb / 2) * 128) / TC_A_BYTES, TC_MAX_STAGES);
}

static void scan_tc_pass(solo_handle *h, IvfIndex &ix, const int64_t *goff, const int32_t *gq, const __half *qh,
                         const uint32_t *qmask, int q_scale_log2, const float *tau, unsigned long long *buf, int32_t *cnt,
                         int cap, DevBuf &item_cnt, DevBuf &item_off, DevBuf &items, bool pairs, bool wide, int64_t len_lo,
                         int64_t len_hi, int wide_nb = 0) {
    const int nlist = ix.nlist;
    int nb, stages;
    if (pairs) tc2_smem_plan(ix, &nb, &stages);
    else if (wide) {
        nb = wide_nb >= 32 ? wide_nb : (h->opt_tc_nb >= 32 ? h->opt_tc_nb : 256);
        stages = std::min((TC_SMEM_MAX - 1024 - (int)sizeof(TcBarriers) - 64) / (TC_A_BYTES + nb * 128), TC_MAX_STAGES);
    } else tc_smem_plan(ix, &nb, &stages, 256, h->opt_tc_nb);
    SOLO_REQUIRE(nb >= 32 && stages >= 2, SOLO_ECAPACITY, "dim %d too large for the tensor-core scan", ix.dim);
    item_cnt.ensure((size_t)nlist * sizeof(int32_t));
    item_off.ensure((size_t)(nlist + 1) * sizeof(int64_t));
    // every list contributes at most ceil(len / nb) <= len / nb + 1 items (second half of the buffer: the sorted copy)
    const size_t max_items = (size_t)(ix.nstored / nb + nlist + 1);
    items.ensure(2 * max_items * sizeof(TcItem));
    tc_item_count_kernel<<<div_up(nlist, 256), 256, 0, h->stream>>>(goff, ix.list_off.as<int64_t>(), nlist, nb, len_lo,
                                                                    len_hi, item_cnt.as<int32_t>());
    scan_counts_i32(h, item_cnt.as<int32_t>(), nlist, item_off.as<int64_t>());
    tc_item_fill_kernel<<<div_up(nlist, 256), 256, 0, h->stream>>>(goff, ix.list_off.as<int64_t>(),
                                                                   item_off.as<int64_t>(), nlist, nb, items.as<TcItem>());
    TcScanArgs a;
    a.items = items.as<TcItem>();
    if (h->opt_sort_items) {
        tc_item_sort_kernel<<<1, 1024, 0, h->stream>>>(items.as<TcItem>(), item_off.as<int64_t>() + nlist,
                                                       items.as<TcItem>() + max_items, pairs ? 2 * TC_BM : TC_BM);
        a.items = items.as<TcItem>() + max_items;
        h->launches++;
    }
    a.item_off = item_off.as<int64_t>();
    a.gq = gq;
    a.qh = qh;
    a.qmask = qmask;
    a.nlist = nlist;
    a.dim = ix.dim;
    a.nb = nb;
    a.stages = stages;
    { static const int env_kbb = getenv("SOLO_TC_KBB") ? atoi(getenv("SOLO_TC_KBB")) : 0; a.kbb = env_kbb > 0 ? env_kbb : h->opt_tc_kbb; }
    a.inv_scale = ldexpf(1.f, -(ix.scale_log2 + q_scale_log2));
    a.tau = tau;
    a.buf = buf;
    a.cnt = cnt;
    a.cap = cap;
    a.dense_out = nullptr;
    a.dense_ld = 0;
    a.prof = tc_prof_buffer();
    static const int v_debug = getenv("SOLO_TC_DEBUG") ? atoi(getenv("SOLO_TC_DEBUG")) : 0;
    a.debug = v_debug | h->opt_tc_debug;
    if (h->opt_tc_stages >= 2 && h->opt_tc_stages < stages && !pairs) a.stages = stages = h->opt_tc_stages;
    const int num_kb = (ix.dim + TC_BK - 1) / TC_BK;
    if (pairs) {
        const size_t smem2 = (size_t)stages * TC_A_BYTES + (size_t)num_kb * (nb / 2) * 128 + sizeof(Tc2Barriers) + 1024;
        SOLO_REQUIRE(smem2 <= (size_t)TC_SMEM_MAX && stages >= 2, SOLO_ECAPACITY, "pair scan kernel needs %zu bytes of shared memory", smem2);
        CUtensorMap map16;
        memcpy(&map16, ix.tmap16_storage, sizeof map16);
        auto kern2 = num_kb == 13 ? scan_tc2_kernel<13> : scan_tc2_kernel<0>;
        SOLO_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        a.prof = nullptr;
        kern2<<<kNumSMs, TC_THREADS, smem2, h->stream>>>(map16, a);  // 74 clusters of two CTAs
        SOLO_CUDA(cudaGetLastError());
        h->launches += 3;
        return;
    }
    const size_t smem = (size_t)stages * TC_A_BYTES + (size_t)(wide ? stages : num_kb) * nb * 128 + sizeof(TcBarriers) + 1024;
    SOLO_REQUIRE(smem <= (size_t)TC_SMEM_MAX && stages >= 2, SOLO_ECAPACITY, "scan kernel needs %zu bytes of shared memory", smem);
    a.wide = wide ? 1 : 0;
    CUtensorMap map;
    memcpy(&map, ix.tmap_storage, sizeof map);
    static const int v_epi2 = getenv("SOLO_TC_EPI2") ? 1 : 0;
    void (*kern)(const CUtensorMap, const CUtensorMap, TcScanArgs) = nullptr;
    if (num_kb == 13) kern = v_epi2 ? scan_tc_kernel<0, 13, 2, 3> : scan_tc_kernel<0, 13, 1, 3>;
    else kern = scan_tc_kernel<0, 0, 1, 3>;
    SOLO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<kNumSMs, TC_THREADS, smem, h->stream>>>(map, map, a);   // (no query map: gathered groups)
    SOLO_CUDA(cudaGetLastError());
    h->launches += 3;
    tc_prof_report(h, "scan", kNumSMs);
}

// The swapped-operand scan (scan_ts_kernel): usable when the list chunk (dim / 2 columns) and an accumulator of at
// least 96 columns fit the 512 columns of tensor memory.
static int ts_tile_queries(const IvfIndex &ix, int want) {
    const int a_cols = (ix.dim + 15) / 16 * 8;
    if (ix.dim % 16 != 0) return 0;
    if (want == 112 && a_cols + 112 <= 512) return 112;
    return a_cols + 96 <= 512 ? 96 : 0;
}

static void scan_ts_pass(solo_handle *h, IvfIndex &ix, const int64_t *goff, const int32_t *gq, const __half *qh,
                         const uint32_t *qmask, int q_scale_log2, const float *tau, unsigned long long *buf, int32_t *cnt,
                         int cap, DevBuf &item_cnt, DevBuf &item_off, DevBuf &items, int nq_tile, bool dense) {
    const int nlist = ix.nlist;
    const int nb = TS_MT;
    const int avail = TC_SMEM_MAX - 1024 - (int)sizeof(TsBarriers) - 8 * 64 * 8 - TS_SBOX * TS_BOX_BYTES;
    int stages = avail / (nq_tile * 128) / TS_BATCH * TS_BATCH;
    stages = std::min(stages, TS_MAX_SLOTS * TS_BATCH);
    static const int env_st = getenv("SOLO_TS_STAGES") ? atoi(getenv("SOLO_TS_STAGES")) : 0;
    if (env_st >= 2 * TS_BATCH && env_st <= stages) stages = env_st / TS_BATCH * TS_BATCH;
    SOLO_REQUIRE(stages >= 2 * TS_BATCH, SOLO_ECAPACITY, "no room for the query ring of the swapped scan");
    item_cnt.ensure((size_t)nlist * sizeof(int32_t));
    item_off.ensure((size_t)(nlist + 1) * sizeof(int64_t));
    items.ensure((size_t)(ix.nstored / nb + nlist + 1) * sizeof(TcItem));
    const int64_t all = (int64_t)1 << 40;
    tc_item_count_kernel<<<div_up(nlist, 256), 256, 0, h->stream>>>(goff, ix.list_off.as<int64_t>(), nlist, nb, 0, all,
                                                                    item_cnt.as<int32_t>());
    scan_counts_i32(h, item_cnt.as<int32_t>(), nlist, item_off.as<int64_t>());
    tc_item_fill_kernel<<<div_up(nlist, 256), 256, 0, h->stream>>>(goff, ix.list_off.as<int64_t>(),
                                                                   item_off.as<int64_t>(), nlist, nb, items.as<TcItem>());
    TcScanArgs a;
    memset(&a, 0, sizeof a);
    a.items = items.as<TcItem>();
    a.item_off = item_off.as<int64_t>();
    a.gq = gq;
    a.qh = qh;
    a.qmask = qmask;
    a.nlist = nlist;
    a.dim = ix.dim;
    a.nb = nb;
    a.stages = stages;
    a.inv_scale = ldexpf(1.f, -(ix.scale_log2 + q_scale_log2));
    a.tau = tau;
    a.buf = buf;
    a.cnt = cnt;
    a.cap = cap;
    a.a_col0 = nq_tile;
    static const int v_debug = getenv("SOLO_TC_DEBUG") ? atoi(getenv("SOLO_TC_DEBUG")) : 0;
    a.debug = v_debug | h->opt_tc_debug;
    const int num_kb = (ix.dim + TC_BK - 1) / TC_BK;
    const size_t smem = (size_t)stages * nq_tile * 128 + (size_t)TS_SBOX * TS_BOX_BYTES + sizeof(TsBarriers) + 8 * 64 * 8 + 1024;
    SOLO_REQUIRE(smem <= (size_t)TC_SMEM_MAX, SOLO_ECAPACITY, "swapped scan kernel needs %zu bytes of shared memory", smem);
    CUtensorMap map;
    memcpy(&map, ix.tmap_storage, sizeof map);
    void (*kern)(const CUtensorMap, TcScanArgs) = nullptr;
    if (nq_tile == 112) {
        if (dense) kern = num_kb == 13 ? scan_ts_kernel<13, 112, true> : scan_ts_kernel<0, 112, true>;
        else kern = num_kb == 13 ? scan_ts_kernel<13, 112, false> : scan_ts_kernel<0, 112, false>;
    } else {
        if (dense) kern = num_kb == 13 ? scan_ts_kernel<13, 96, true> : scan_ts_kernel<0, 96, true>;
        else kern = num_kb == 13 ? scan_ts_kernel<13, 96, false> : scan_ts_kernel<0, 96, false>;
    }
    SOLO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<kNumSMs, TC_THREADS, smem, h->stream>>>(map, a);
    SOLO_CUDA(cudaGetLastError());
    h->launches += 3;
}

// K3 for one round. Lists up to `hybrid` vectors go through the resident-chunk kernel (<= 96-row chunks,
// read once per item), longer lists through the streamed-chunk variant (up to 256 rows per MMA, the
// chunk re-read from L2 per query block): fewer, larger tiles where the issue-bound pipeline pays most.
void launch_scan_tc(solo_handle *h, IvfIndex &ix, const int64_t *goff, const int32_t *gq, const __half *qh,
                    const uint32_t *qmask, int q_scale_log2, const float *tau, unsigned long long *buf, int32_t *cnt,
                    int cap, DevBuf &item_cnt, DevBuf &item_off, DevBuf &items, bool dense_round) {
    SOLO_REQUIRE(ix.tmap_valid, SOLO_ESTATE, "tensor map missing");
    static const bool env_pairs = getenv("SOLO_TC_PAIRS") && atoi(getenv("SOLO_TC_PAIRS")) != 0;
    const bool pairs = h->opt_scan_pairs || env_pairs;
    static const int env_wide = getenv("SOLO_TC_WIDE") ? atoi(getenv("SOLO_TC_WIDE")) : -1;
    const bool wide = !pairs && (env_wide >= 0 ? env_wide != 0 : h->opt_scan_wide);
    static const int env_hybrid = getenv("SOLO_TC_HYBRID") ? atoi(getenv("SOLO_TC_HYBRID")) : -1;
    const int64_t hybrid = env_hybrid >= 0 ? env_hybrid : h->opt_scan_hybrid;
    const int64_t all = (int64_t)1 << 40;
    static const int env_ts = getenv("SOLO_TC_TS") ? atoi(getenv("SOLO_TC_TS")) : -1;
    const int ts = env_ts >= 0 ? env_ts : h->opt_scan_ts;   // 0: off, 96 / 112: queries per tile
    const int nq_tile = (ts > 0 && !pairs && !wide) ? ts_tile_queries(ix, ts) : 0;
    if (nq_tile > 0) {
        scan_ts_pass(h, ix, goff, gq, qh, qmask, q_scale_log2, tau, buf, cnt, cap, item_cnt, item_off, items, nq_tile, dense_round);
        return;
    }
    if (dense_round && h->opt_round0_wide > 0 && !pairs && !wide) {
        // first round: every item is a single tile of a few queries, its list chunk is used once. Streaming the chunk
        // through the ring next to the query stage keeps the loads of the following items in flight (the resident
        // layout looks one item ahead only, and the chunk comes from DRAM)
        scan_tc_pass(h, ix, goff, gq, qh, qmask, q_scale_log2, tau, buf, cnt, cap, item_cnt, item_off, items, false, true, 0,
                     all, h->opt_round0_wide);
        return;
    }
    if (pairs || wide || hybrid <= 0 || ix.max_list_len <= hybrid) {
        scan_tc_pass(h, ix, goff, gq, qh, qmask, q_scale_log2, tau, buf, cnt, cap, item_cnt, item_off, items, pairs, wide, 0,
                     all);
        return;
    }
    scan_tc_pass(h, ix, goff, gq, qh, qmask, q_scale_log2, tau, buf, cnt, cap, item_cnt, item_off, items, false, false, 0,
                 hybrid);
    scan_tc_pass(h, ix, goff, gq, qh, qmask, q_scale_log2, tau, buf, cnt, cap, item_cnt, item_off, items, false, true,
                 hybrid, all);
}

// K2 on the tensor cores: approximate scores of every query against the centroids [0, n_lists) (n_lists < 0: all).
// The centroid table is scanned like one inverted list whose query group is the whole batch, so the A tiles are
// consecutive query rows and come in by TMA. Two forms:
//   dense (tau == null): out[q * ld + c] (ld a multiple of 32);
//   thresholded (tau != null): only scores >= tau[q] leave the SM, appended as (score, centroid) pairs to
//   buf[q * cap ..] with cnt[q] counting them (cnt may exceed cap: the caller checks) — no (Q, nlist) matrix.
void launch_coarse_tc(solo_handle *h, IvfIndex &ix, const __half *qh, const uint32_t *qmask, int nq, int q_scale_log2,
                      float *out, int ld, int n_lists, const float *tau, unsigned long long *buf, int32_t *cnt, int cap) {
    SOLO_REQUIRE(ix.tmap_cent_valid, SOLO_ESTATE, "centroid tensor map missing");
    int nb, stages;
    // 96-centroid chunks and the four-stage ring (the straight-line MMA issue path of scan_tc_kernel); every item is
    // one chunk x one span of 512 queries, so that the (chunk, span) items spread evenly over the SMs (171 chunks
    // of the 16,384 centroids alone would leave 23 SMs with twice the work of the others)
    static const int env_cnb = getenv("SOLO_COARSE_NB") ? atoi(getenv("SOLO_COARSE_NB")) : 96;
    // (the sampled prefix keeps 64-centroid chunks: its dense output rows are exactly n_lists wide)
    int nb_full, st_full, nb_pref, st_pref;
    tc_smem_plan(ix, &nb_full, &st_full, env_cnb);
    tc_smem_plan(ix, &nb_pref, &st_pref, 64);
    nb = n_lists < 0 ? nb_full : nb_pref;
    stages = n_lists < 0 ? st_full : st_pref;
    SOLO_REQUIRE(nb >= 32 && stages >= 2, SOLO_ECAPACITY, "dim %d too large for the tensor-core scan", ix.dim);
    SOLO_REQUIRE(n_lists < 0 || n_lists % nb_pref == 0 || n_lists >= ix.nlist, SOLO_EINVAL, "coarse prefix must be a multiple of %d centroids", nb_pref);
    const int n_chunks = div_up(ix.nlist, nb_full);
    const int n_pref_chunks =
        n_lists < 0 ? std::max(ix.coarse_items_used, 0) : std::min(div_up(ix.nlist, nb_pref), div_up(n_lists, nb_pref));
    const int q_span = 4 * TC_BM, n_spans = div_up(nq, q_span);
    const int n_items = n_chunks * n_spans;
    const int n_used = n_pref_chunks * n_spans;
    constexpr int HDR = 4;  // int64 header: [0, items of the full table | 0, items of the sampled prefix]
    if (ix.coarse_items_nq != nq || (n_lists >= 0 && ix.coarse_items_used != n_pref_chunks)) {  // tiny: built on the host, cached
        std::vector<unsigned char> hbuf(sizeof(int64_t) * HDR + (size_t)(n_items + n_used) * sizeof(TcItem));
        int64_t *hoff = reinterpret_cast<int64_t *>(hbuf.data());
        hoff[0] = 0;
        hoff[1] = n_items;
        hoff[2] = 0;
        hoff[3] = n_used;
        TcItem *hit = reinterpret_cast<TcItem *>(hbuf.data() + HDR * sizeof(int64_t));
        auto fill = [&](TcItem *dst, int chunks, int rows) {
            for (int c = 0; c < chunks; ++c)
                for (int sp = 0; sp < n_spans; ++sp) {
                    TcItem &t = dst[c * n_spans + sp];
                    t.p0 = c * rows;
                    t.nv = std::min(rows, ix.nlist - c * rows);
                    t.g0 = sp * q_span;
                    t.G = std::min(q_span, nq - sp * q_span);
                }
        };
        fill(hit, n_chunks, nb_full);
        fill(hit + n_items, n_pref_chunks, nb_pref);
        SOLO_CUDA(cudaStreamSynchronize(h->stream));  // earlier launches may still read the old descriptors
        ix.coarse_items.ensure(hbuf.size());
        SOLO_CUDA(cudaMemcpy(ix.coarse_items.p, hbuf.data(), hbuf.size(), cudaMemcpyHostToDevice));
        ix.coarse_items_nq = nq;
        ix.coarse_items_used = n_pref_chunks;
    }
    DevBuf &items = ix.coarse_items;
    TcScanArgs a;
    memset(&a, 0, sizeof a);
    a.items = reinterpret_cast<const TcItem *>(items.as<unsigned char>() + HDR * sizeof(int64_t)) + (n_lists < 0 ? 0 : n_items);
    a.item_off = items.as<int64_t>() + (n_lists < 0 ? 0 : 2);
    a.gq = nullptr;
    a.qh = qh;
    a.qmask = qmask;
    a.nlist = 1;
    a.dim = ix.dim;
    a.nb = nb;
    a.stages = stages;
    { static const int env_kbb = getenv("SOLO_TC_KBB") ? atoi(getenv("SOLO_TC_KBB")) : 2; a.kbb = env_kbb; }
    a.inv_scale = ldexpf(1.f, -(ix.cent_scale_log2 + q_scale_log2));
    a.dense_out = out;
    a.dense_ld = ld;
    a.tau = tau;
    a.buf = buf;
    a.cnt = cnt;
    a.cap = cap;
    a.prof = tc_prof_buffer();
    const int num_kb = (ix.dim + TC_BK - 1) / TC_BK;
    const size_t smem = (size_t)stages * TC_A_BYTES + (size_t)num_kb * nb * 128 + sizeof(TcBarriers) + 1024;
    CUtensorMap map, qmap;
    memcpy(&map, ix.tmap_cent_storage, sizeof map);
    static const bool v_gather = getenv("SOLO_COARSE_GATHER") != nullptr;   // cross-check: the 16-byte gather for A
    alignas(64) unsigned char qstore[128];
    if (!v_gather) tc_encode_rows(qstore, qh, ix.dim, nq, TC_BM);
    memcpy(&qmap, v_gather ? (const void *)ix.tmap_cent_storage : (const void *)qstore, sizeof qmap);
    void (*kern)(const CUtensorMap, const CUtensorMap, TcScanArgs);
    if (tau) {
        // (two 32-column groups per reservation: one atomic per query row and 64-centroid chunk)
        if (v_gather) kern = num_kb == 13 ? scan_tc_kernel<0, 13, 2, 3> : scan_tc_kernel<0, 0, 2, 3>;
        else kern = num_kb == 13 ? scan_tc_kernel<0, 13, 2, 7> : scan_tc_kernel<0, 0, 2, 7>;
    } else {
        if (v_gather) kern = num_kb == 13 ? scan_tc_kernel<1, 13, 1, 3> : scan_tc_kernel<1, 0, 1, 3>;
        else kern = num_kb == 13 ? scan_tc_kernel<1, 13, 1, 7> : scan_tc_kernel<1, 0, 1, 7>;
    }
    SOLO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<std::min(kNumSMs, n_lists < 0 ? n_items : n_used), TC_THREADS, smem, h->stream>>>(map, qmap, a);
    (void)n_pref_chunks;
    SOLO_CUDA(cudaGetLastError());
    h->launches += 1;
    tc_prof_report(h, "coarse", std::min(kNumSMs, n_lists < 0 ? n_items : n_used));
}

// Fall-back of the compact probe selection: dense coarse scores for the queries of a device-side list (n_fail of
// them, known only on the device: the items are rewritten there, no host round trip). Rows land at out[q * ld].
__global__ void coarse_fallback_items_kernel(int nlist, int nb, int n_items, const int32_t *__restrict__ n_fail,
                                             TcItem *__restrict__ dst, int64_t *__restrict__ hdr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nf = *n_fail;
    if (i == 0) {
        hdr[0] = 0;
        hdr[1] = nf > 0 ? n_items : 0;
    }
    if (i < n_items) {   // centroid chunk i against the listed queries
        TcItem it;
        it.p0 = i * nb;
        it.nv = min(nb, nlist - i * nb);
        it.g0 = 0;
        it.G = nf;
        dst[i] = it;
    }
}

void launch_coarse_tc_listed(solo_handle *h, IvfIndex &ix, const __half *qh, const uint32_t *qmask, int q_scale_log2,
                             float *out, int ld, const int32_t *q_list, const int32_t *n_listed) {
    SOLO_REQUIRE(ix.tmap_cent_valid && ix.coarse_items_nq >= 0, SOLO_ESTATE, "coarse items missing");
    int nb, stages;
    tc_smem_plan(ix, &nb, &stages, 64);
    const int n_items = div_up(ix.nlist, nb);
    DevBuf &fb = ix.coarse_fb_items;
    fb.ensure(2 * sizeof(int64_t) + (size_t)n_items * sizeof(TcItem));
    TcItem *fb_items = reinterpret_cast<TcItem *>(fb.as<unsigned char>() + 2 * sizeof(int64_t));
    coarse_fallback_items_kernel<<<div_up(n_items, 256), 256, 0, h->stream>>>(ix.nlist, nb, n_items, n_listed, fb_items,
                                                                              fb.as<int64_t>());
    TcScanArgs a;
    memset(&a, 0, sizeof a);
    a.items = fb_items;
    a.item_off = fb.as<int64_t>();
    a.gq = q_list;
    a.qh = qh;
    a.qmask = qmask;
    a.nlist = 1;
    a.dim = ix.dim;
    a.nb = nb;
    a.stages = stages;
    a.kbb = 2;
    a.inv_scale = ldexpf(1.f, -(ix.cent_scale_log2 + q_scale_log2));
    a.dense_out = out;
    a.dense_ld = ld;
    const int num_kb = (ix.dim + TC_BK - 1) / TC_BK;
    const size_t smem = (size_t)stages * TC_A_BYTES + (size_t)num_kb * nb * 128 + sizeof(TcBarriers) + 1024;
    CUtensorMap map;
    memcpy(&map, ix.tmap_cent_storage, sizeof map);
    auto kern = num_kb == 13 ? scan_tc_kernel<1, 13, 1, 3> : scan_tc_kernel<1, 0, 1, 3>;
    SOLO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<std::min(kNumSMs, n_items), TC_THREADS, smem, h->stream>>>(map, map, a);
    SOLO_CUDA(cudaGetLastError());
    h->launches += 2;
}

}  // namespace solo

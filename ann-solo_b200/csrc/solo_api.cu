// C-ABI of libsolo_b200.so (see include/solo_b200.h for the contract and the reference
// interfaces each entry point replaces).
#include <algorithm>
#include <cmath>

#include "ivf.cuh"
#include "k6_ssm_features.cuh"
#include "solo_common.cuh"

using namespace solo;

static thread_local std::string g_create_error;

static const char *kStageNames[ST_COUNT] = {"vectorize", "coarse", "probe_select", "group", "scan",
                                            "topk",      "candidates", "score", "h2d", "d2h"};
namespace solo {
const char *stage_name(int st) { return st >= 0 && st < ST_COUNT ? kStageNames[st] : ""; }
}  // namespace solo

// ---------------------------------------------------------------- host helpers

// MurmurHash3_x86_32 (Austin Appleby, public domain algorithm) of the decimal string of `bin`,
// seed 42, as mmh3.hash(str(bin_idx), 42, signed=False) at reference spectrum.py:163.
static uint32_t murmur3_decimal_host(long long bin, uint32_t seed) {
    char buf[32];
    int len = snprintf(buf, sizeof buf, "%lld", bin);
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t h = seed;
    int nblocks = len / 4;
    for (int i = 0; i < nblocks; ++i) {
        uint32_t k;
        memcpy(&k, buf + 4 * i, 4);
        k *= c1;
        k = (k << 15) | (k >> 17);
        k *= c2;
        h ^= k;
        h = (h << 13) | (h >> 19);
        h = h * 5u + 0xe6546b64u;
    }
    uint32_t k = 0;
    const unsigned char *tail = (const unsigned char *)buf + 4 * nblocks;
    switch (len & 3) {
        case 3: k ^= (uint32_t)tail[2] << 16; /* fallthrough */
        case 2: k ^= (uint32_t)tail[1] << 8;  /* fallthrough */
        case 1:
            k ^= tail[0];
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

static double py_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0 && ((b < 0) != (m < 0))) m += b;
    return m;
}

template <typename F>
static int guarded(solo_handle *h, F &&f) {
    try {
        if (h) SOLO_CUDA(cudaSetDevice(h->device));
        f();
        return SOLO_OK;
    } catch (const Error &e) {
        if (h) h->last_error = e.msg;
        else g_create_error = e.msg;
        return e.code;
    } catch (const std::exception &e) {
        if (h) h->last_error = e.what();
        else g_create_error = e.what();
        return SOLO_ECUDA;
    }
}

template <typename F>
static int guarded_nohandle(char *errbuf, int errbuf_len, F &&f) {
    if (errbuf && errbuf_len > 0) errbuf[0] = 0;
    try {
        f();
        return SOLO_OK;
    } catch (const Error &e) {
        if (errbuf && errbuf_len > 0) snprintf(errbuf, errbuf_len, "%s", e.msg.c_str());
        return e.code;
    } catch (const std::exception &e) {
        if (errbuf && errbuf_len > 0) snprintf(errbuf, errbuf_len, "%s", e.what());
        return SOLO_EINVAL;
    }
}

static void h2d(solo_handle *h, DevBuf &b, const void *src, size_t bytes, cudaStream_t st = nullptr) {
    b.ensure(std::max<size_t>(bytes, 16));
    if (bytes) SOLO_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st ? st : h->stream));
}

// events + copy stream of the asynchronous staging API, created on first use
static solo_handle::SlotSync &slot_sync(solo_handle *h, int slot) {
    if (!h->copy_stream) SOLO_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    solo_handle::SlotSync &s = h->slot_sync[slot];
    if (!s.staged) {
        SOLO_CUDA(cudaEventCreateWithFlags(&s.staged, cudaEventDisableTiming));
        SOLO_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        SOLO_CUDA(cudaEventCreateWithFlags(&s.fetched, cudaEventDisableTiming));
        SOLO_CUDA(cudaMallocHost(&s.n_over, sizeof(int32_t)));
        *s.n_over = 0;
    }
    return s;
}

static void drain_profile(solo_handle *h) {
    for (int s = 0; s < ST_COUNT; ++s) {
        for (auto &pr : h->prof[s].pending) {
            float ms = 0.f;
            if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess)
                h->prof[s].ms += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        h->prof[s].pending.clear();
    }
}

static LibraryStore &get_lib(solo_handle *h, int charge) {
    auto it = h->libs.find(charge);
    SOLO_REQUIRE(it != h->libs.end(), SOLO_ESTATE, "no library loaded for charge %d", charge);
    return it->second;
}

static IvfIndex &get_ivf(solo_handle *h, int charge, bool must_exist) {
    auto it = h->ivf.find(charge);
    if (it == h->ivf.end()) {
        SOLO_REQUIRE(!must_exist, SOLO_ESTATE, "no ANN index for charge %d", charge);
        return h->ivf[charge];
    }
    return it->second;
}

// ---------------------------------------------------------------- window-only candidates (brute force)

namespace solo {

__device__ __forceinline__ bool window_ok(double qm, float lm32, int charge, double tol, int mode) {
    const double lm = (double)lm32;
    // reference spectral_library.py:421-427 (numexpr evaluates in float64)
    if (mode == SOLO_TOL_DA) return __dmul_rn(fabs(__dsub_rn(qm, lm)), (double)charge) <= tol;
    return __dmul_rn(__ddiv_rn(fabs(__dsub_rn(qm, lm)), lm), 1000000.0) <= tol;
}

// one CTA per query; pass 0 counts, pass 1 fills in ascending library row order
__global__ void __launch_bounds__(256)
window_candidates_kernel(const double *__restrict__ q_prec_mz, const float *__restrict__ lib_prec_mz32,
                         const uint8_t *__restrict__ lib_valid, int64_t n_lib, int charge, double tol, int mode,
                         int32_t *__restrict__ counts, const int64_t *__restrict__ off, int32_t *__restrict__ ids) {
    __shared__ int s_wsum[8];
    __shared__ int64_t s_run;
    const int q = blockIdx.x;
    const double qm = q_prec_mz[q];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_lib; base += 256) {
        const int64_t p = base + threadIdx.x;
        bool ok = p < n_lib && lib_valid[p] && window_ok(qm, lib_prec_mz32[p], charge, tol, mode);
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_wsum[warp] = __popc(m);
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) before += s_wsum[w];
            tot += s_wsum[w];
        }
        if (ids && ok) ids[off[q] + s_run + before + __popc(m & ((1u << lane) - 1u))] = (int32_t)p;
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (!ids && threadIdx.x == 0) counts[q] = (int)s_run;
}

// The same candidate sets from the m/z-sorted view of the library (built at load time): binary search
// of a slightly widened m/z interval, then the exact reference predicate on the rows inside it.
// O(log N + rows in the window) per query instead of O(N): what the level-1 'std' search (a few ppm)
// needs. Candidates come out in m/z order; the scorer breaks score ties by library row.
__global__ void __launch_bounds__(128)
window_candidates_sorted_kernel(const double *__restrict__ q_prec_mz, const float *__restrict__ sorted_mz32,
                                const int32_t *__restrict__ sorted_row, const uint8_t *__restrict__ lib_valid,
                                int64_t n_lib, int charge, double tol, int mode, int32_t *__restrict__ counts,
                                const int64_t *__restrict__ off, int32_t *__restrict__ ids) {
    __shared__ int s_n;
    __shared__ int64_t s_lo, s_hi;
    const int q = blockIdx.x;
    const double qm = q_prec_mz[q];
    if (threadIdx.x == 0) {
        s_n = 0;
        double lo_mz, hi_mz;
        if (mode == SOLO_TOL_DA) {
            const double w = charge != 0 ? tol / fabs((double)charge) : INFINITY;
            lo_mz = qm - w;
            hi_mz = qm + w;
        } else {
            const double t = tol * 1e-6;
            lo_mz = qm / (1.0 + t);
            hi_mz = t < 1.0 ? qm / (1.0 - t) : INFINITY;
        }
        // widen: the exact predicate below decides, the interval only has to contain every passing row
        lo_mz = lo_mz - 1e-6 - 1e-9 * fabs(lo_mz);
        hi_mz = hi_mz + 1e-6 + 1e-9 * fabs(hi_mz);
        if (!(lo_mz == lo_mz) || !(hi_mz == hi_mz)) {  // NaN query: scan everything, the predicate rejects
            lo_mz = -INFINITY;
            hi_mz = INFINITY;
        }
        int64_t a = 0, b = n_lib;  // first row with mz >= lo_mz
        while (a < b) {
            const int64_t m = (a + b) >> 1;
            if ((double)sorted_mz32[m] < lo_mz) a = m + 1;
            else b = m;
        }
        s_lo = a;
        b = n_lib;  // first row with mz > hi_mz
        while (a < b) {
            const int64_t m = (a + b) >> 1;
            if ((double)sorted_mz32[m] <= hi_mz) a = m + 1;
            else b = m;
        }
        s_hi = a;
    }
    __syncthreads();
    const int64_t lo = s_lo, hi = s_hi;
    int mine = 0;
    for (int64_t p = lo + threadIdx.x; p < hi; p += blockDim.x) {
        const int row = sorted_row[p];
        if (lib_valid[row] && window_ok(qm, sorted_mz32[p], charge, tol, mode)) {
            if (ids) ids[off[q] + atomicAdd(&s_n, 1)] = row;
            else ++mine;
        }
    }
    if (!ids) {
        if (mine) atomicAdd(&s_n, mine);
        __syncthreads();
        if (threadIdx.x == 0) counts[q] = s_n;
    }
}

}  // namespace solo

// ================================================================ C-ABI

extern "C" {

const char *solo_version(void) { return "solo_b200 0.1 (sm_100a)"; }

int solo_create(int device, solo_handle **out) {
    if (!out) return SOLO_EINVAL;
    *out = nullptr;
    solo_handle *h = nullptr;
    int rc = guarded(nullptr, [&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        SOLO_REQUIRE(e == cudaSuccess && ndev > 0, SOLO_ECUDA,
                     "no CUDA device available (%s); libsolo_b200 has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        SOLO_REQUIRE(device >= 0 && device < ndev, SOLO_EINVAL, "device %d out of range (%d devices)", device, ndev);
        SOLO_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        SOLO_CUDA(cudaGetDeviceProperties(&prop, device));
        SOLO_REQUIRE(prop.major == 10, SOLO_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                     device, prop.major, prop.minor);
        h = new solo_handle();
        h->device = device;
        SOLO_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
    });
    if (rc != SOLO_OK) {
        delete h;
        return rc;
    }
    rc = solo_set_vectorizer(h, 11.0, 2010.0, 0.04, 800);  // reference defaults, config.py:71-76,179-185
    if (rc != SOLO_OK) {
        g_create_error = h->last_error;
        solo_destroy(h);
        return rc;
    }
    *out = h;
    return SOLO_OK;
}

void solo_destroy(solo_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    drain_profile(h);
    auto rel = [](DevBuf &b) { b.release(); };
    rel(h->lut);
    for (auto &kv : h->libs) {
        LibraryStore &L = kv.second;
        rel(L.mz); rel(L.inten); rel(L.chg); rel(L.off); rel(L.prec_mz); rel(L.prec_mz32); rel(L.prec_z); rel(L.valid); rel(L.meta); rel(L.table); rel(L.sorted_mz32); rel(L.sorted_row);
    }
    for (auto &kv : h->ivf) {
        IvfIndex &x = kv.second;
        rel(x.cent); rel(x.cent_h); rel(x.list_off); rel(x.list_ids); rel(x.vec_h); rel(x.sp_off); rel(x.sp_idx);
        rel(x.sp_val); rel(x.row_list); rel(x.row_off); rel(x.row_idx); rel(x.row_val); rel(x.stats);
    }
    rel(h->q_mz); rel(h->q_mz_vec); rel(h->q_int); rel(h->q_off); rel(h->q_prec_mz);
    for (auto &b : h->scratch) rel(b);
    rel(h->r_best_row); rel(h->r_best_score); rel(h->r_n_pairs); rel(h->r_pairs); rel(h->r_n_cand); rel(h->r_ovf);
    for (auto &kv : h->parked)
        for (auto &b : kv.second.b) rel(b);
    for (auto &kv : h->slot_sync) {
        if (kv.second.staged) cudaEventDestroy(kv.second.staged);
        if (kv.second.done) cudaEventDestroy(kv.second.done);
        if (kv.second.fetched) cudaEventDestroy(kv.second.fetched);
        if (kv.second.n_over) cudaFreeHost(kv.second.n_over);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

const char *solo_last_error(const solo_handle *h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int solo_set_stream(solo_handle *h, void *cuda_stream) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    });
}

int solo_set_option(solo_handle *h, const char *key, int64_t value) {
    if (!h || !key) return SOLO_EINVAL;
    return guarded(h, [&] {
        if (strcmp(key, "scan_engine") == 0) {
            SOLO_REQUIRE(value == 0 || value == 1, SOLO_EINVAL, "scan_engine must be 0 (tcgen05) or 1 (exact CUDA cores)");
            h->opt_scan_exact = value == 1;
        } else if (strcmp(key, "round0_scores") == 0) {
            SOLO_REQUIRE(value >= 1024 && value <= 16384, SOLO_EINVAL, "round0_scores must be in [1024, 16384]");
            h->opt_round0_scores = (int)value;
        } else if (strcmp(key, "scan_wide") == 0) {
            h->opt_scan_wide = value != 0;
        } else if (strcmp(key, "scan_hybrid") == 0) {
            SOLO_REQUIRE(value >= 0 && value <= 12288, SOLO_EINVAL, "scan_hybrid must be in [0, 12288]");
            h->opt_scan_hybrid = (int)value;
        } else if (strcmp(key, "scan_pairs") == 0) {
            h->opt_scan_pairs = value != 0;
        } else if (strcmp(key, "round0_wide") == 0) {
            SOLO_REQUIRE(value >= 0 && value <= 256 && value % 32 == 0, SOLO_EINVAL, "round0_wide must be a multiple of 32 in [0, 256]");
            h->opt_round0_wide = (int)value;
        } else if (strcmp(key, "nvtx") == 0) {
            h->opt_nvtx = value != 0;
        } else if (strcmp(key, "sort_items") == 0) {
            h->opt_sort_items = value != 0;
        } else if (strcmp(key, "tc_stages") == 0) {
            h->opt_tc_stages = (int)value;
        } else if (strcmp(key, "tc_debug") == 0) {
            h->opt_tc_debug = (int)value;
        } else if (strcmp(key, "scan_ts") == 0) {
            SOLO_REQUIRE(value == 0 || value == 96 || value == 112, SOLO_EINVAL, "scan_ts must be 0, 96 or 112");
            h->opt_scan_ts = (int)value;
        } else if (strcmp(key, "front_probes") == 0) {
            h->opt_front_probes = value != 0;
        } else if (strcmp(key, "train_balance") == 0) {
            SOLO_REQUIRE(value >= 0 && value <= 64, SOLO_EINVAL, "train_balance must be in [0, 64]");
            h->opt_train_balance = (int)value;
        } else if (strcmp(key, "compact_probes") == 0) {
            h->opt_compact_probes = value != 0;
        } else if (strcmp(key, "tc_nb") == 0) {
            SOLO_REQUIRE(value >= 0 && value <= 256 && value % 32 == 0, SOLO_EINVAL, "tc_nb must be a multiple of 32 in [0, 256]");
            h->opt_tc_nb = (int)value;
        } else if (strcmp(key, "tc_kbb") == 0) {
            SOLO_REQUIRE(value >= 1 && value <= 6, SOLO_EINVAL, "tc_kbb must be in [1, 6]");
            h->opt_tc_kbb = (int)value;
        } else {
            SOLO_REQUIRE(false, SOLO_EINVAL, "unknown option '%s'", key);
        }
    });
}

int solo_synchronize(solo_handle *h) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] { SOLO_CUDA(cudaStreamSynchronize(h->stream)); });
}

// ---------------------------------------------------------------- K1

int solo_set_vectorizer(solo_handle *h, double min_mz, double max_mz, double bin_size, int hash_len) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(bin_size > 0 && max_mz > min_mz, SOLO_EINVAL, "bad m/z range or bin size");
        SOLO_REQUIRE(hash_len >= 1 && hash_len <= 1536, SOLO_EINVAL, "hash_len must be in [1, 1536] (got %d)", hash_len);
        // reference spectrum.py:123-143 get_dim
        double start = min_mz - py_mod(min_mz, bin_size);
        double end = max_mz + bin_size - py_mod(max_mz, bin_size);
        int64_t n_bins = (int64_t)nearbyint((end - start) / bin_size);
        SOLO_REQUIRE(n_bins > 0 && n_bins < (1 << 26), SOLO_EINVAL, "unreasonable number of m/z bins");
        h->min_mz = min_mz;
        h->max_mz = max_mz;
        h->bin_size = bin_size;
        h->hash_len = hash_len;
        h->n_bins = n_bins;
        h->min_bound = start;
        h->h_lut.resize(n_bins + 2);
        for (int64_t b = 0; b < n_bins + 2; ++b)
            h->h_lut[b] = (uint16_t)(murmur3_decimal_host(b, 42u) % (uint32_t)hash_len);
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        h2d(h, h->lut, h->h_lut.data(), h->h_lut.size() * sizeof(uint16_t));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int solo_hash_slot(solo_handle *h, int64_t bin_idx, int32_t *slot) {
    if (!h || !slot) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(bin_idx >= 0 && bin_idx < h->n_bins + 2, SOLO_EINVAL, "bin outside the device LUT");
        uint16_t v = 0;
        SOLO_CUDA(cudaMemcpy(&v, h->lut.as<uint16_t>() + bin_idx, sizeof v, cudaMemcpyDeviceToHost));
        *slot = v;
    });
}

int solo_vectorize(solo_handle *h, const void *mz, int mz_is_f64, const float *intensity, const int64_t *offsets,
                   int64_t n, int norm, float *out) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(n >= 0 && offsets && out, SOLO_EINVAL, "null argument");
        if (n == 0) return;
        const int64_t np = offsets[n];
        SOLO_REQUIRE(offsets[0] == 0 && np >= 0, SOLO_EINVAL, "offsets must start at 0");
        DevBuf &dmz = h->scratch[0], &din = h->scratch[1], &doff = h->scratch[2], &dout = h->scratch[3];
        h2d(h, dmz, mz, np * (mz_is_f64 ? 8 : 4));
        h2d(h, din, intensity, np * 4);
        h2d(h, doff, offsets, (n + 1) * 8);
        dout.ensure((size_t)n * h->hash_len * sizeof(float));
        launch_vectorize(h, dmz.p, mz_is_f64, din.as<float>(), doff.as<int64_t>(), n, np, norm, dout.as<float>(),
                         nullptr, 0);
        SOLO_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)n * h->hash_len * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    });
}

// ---------------------------------------------------------------- library store

int solo_load_library(solo_handle *h, int charge, const float *mz, const float *intensity, const uint8_t *peak_charge,
                      const int64_t *offsets, const double *prec_mz, const float *prec_mz32,
                      const int32_t *prec_charge, const uint8_t *valid, int64_t n) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(n >= 0 && n < 0x7fffffff, SOLO_EINVAL, "bad library size");
        SOLO_REQUIRE(offsets && prec_mz && prec_charge, SOLO_EINVAL, "null argument");
        SOLO_REQUIRE(offsets[0] == 0, SOLO_EINVAL, "offsets must start at 0");
        const int64_t np = offsets[n];
        int maxp = 0;
        for (int64_t i = 0; i < n; ++i) {
            int64_t len = offsets[i + 1] - offsets[i];
            SOLO_REQUIRE(len >= 0, SOLO_EINVAL, "offsets must be non-decreasing (row %lld)", (long long)i);
            maxp = std::max<int64_t>(maxp, len);
            SOLO_REQUIRE(prec_charge[i] >= 0 && prec_charge[i] <= 7, SOLO_ECAPACITY,
                         "library row %lld has precursor charge %d; the scorer supports charges 0..7", (long long)i,
                         prec_charge[i]);
            // the scorer's merge (SpectrumMatch.cpp:39-55) relies on ascending m/z
            for (int64_t p = offsets[i] + 1; p < offsets[i + 1]; ++p)
                SOLO_REQUIRE(mz[p] >= mz[p - 1], SOLO_EINVAL, "library row %lld: m/z not ascending", (long long)i);
        }
        SOLO_REQUIRE(maxp <= 128, SOLO_ECAPACITY, "library spectra may hold at most 128 peaks (got %d)", maxp);
        LibraryStore &L = h->libs[charge];
        L.n = n;
        L.n_peaks = np;
        L.max_peaks = maxp;
        h2d(h, L.mz, mz, np * 4);
        h2d(h, L.inten, intensity, np * 4);
        if (peak_charge) h2d(h, L.chg, peak_charge, np);
        else {
            L.chg.ensure(std::max<int64_t>(np, 16));
            SOLO_CUDA(cudaMemsetAsync(L.chg.p, 0, std::max<int64_t>(np, 16), h->stream));
        }
        h2d(h, L.off, offsets, (n + 1) * 8);
        h2d(h, L.prec_mz, prec_mz, n * 8);
        std::vector<float> tmp32;
        if (!prec_mz32) {  // reader.py:188-189 np.asarray(..., np.float32)
            tmp32.resize(n);
            for (int64_t i = 0; i < n; ++i) tmp32[i] = (float)prec_mz[i];
            prec_mz32 = tmp32.data();
        }
        h2d(h, L.prec_mz32, prec_mz32, n * 4);
        h2d(h, L.prec_z, prec_charge, n * 4);
        if (valid) h2d(h, L.valid, valid, n);
        else {
            L.valid.ensure(std::max<int64_t>(n, 16));
            SOLO_CUDA(cudaMemsetAsync(L.valid.p, 1, std::max<int64_t>(n, 16), h->stream));
        }
        k5_build_aux(h, L);
        {   // m/z-sorted view for the brute-force candidate search (NaN precursors sort last and never match)
            std::vector<int32_t> order(n);
            for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)i;
            std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
                const float fx = prec_mz32[x], fy = prec_mz32[y];
                if (fx != fx) return false;
                if (fy != fy) return true;
                return fx < fy;
            });
            std::vector<float> smz(n);
            for (int64_t i = 0; i < n; ++i) smz[i] = prec_mz32[order[i]];
            h2d(h, L.sorted_mz32, smz.data(), n * 4);
            h2d(h, L.sorted_row, order.data(), n * 4);
            SOLO_CUDA(cudaStreamSynchronize(h->stream));  // the vectors above go out of scope
        }
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    });
}

// ---------------------------------------------------------------- IVF

int solo_ivf_set_centroids(solo_handle *h, int charge, const float *centroids, int nlist, int dim) {
    if (!h || !centroids) return SOLO_EINVAL;
    return guarded(h, [&] { ivf_set_centroids(h, get_ivf(h, charge, false), centroids, nlist, dim); });
}

int solo_ivf_train(solo_handle *h, int charge, const float *x, int64_t n, int dim, int nlist, int iters, uint64_t seed) {
    if (!h || !x) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(nlist > 0 && nlist <= IVF_MAX_NLIST, SOLO_EINVAL, "nlist must be in [1, %d]", IVF_MAX_NLIST);
        ivf_train(h, get_ivf(h, charge, false), x, n, dim, nlist, std::max(iters, 0), seed);
    });
}

int solo_ivf_train_library(solo_handle *h, int charge, int nlist, int iters, uint64_t seed) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        LibraryStore &L = get_lib(h, charge);
        std::vector<int64_t> hoff(L.n + 1);
        SOLO_CUDA(cudaMemcpy(hoff.data(), L.off.p, (L.n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
        ivf_train_rows(h, get_ivf(h, charge, false), L.n, h->hash_len, nlist, std::max(iters, 0), seed,
                       [&](int64_t r0, int64_t m, float *dst) {
                           launch_vectorize(h, L.mz.p, 0, L.inten.as<float>(), L.off.as<int64_t>() + r0, m,
                                            hoff[r0 + m] - hoff[r0], 1, dst, nullptr, 0);
                       });
    });
}

int solo_ivf_get_centroids(solo_handle *h, int charge, float *centroids) {
    if (!h || !centroids) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        SOLO_CUDA(cudaMemcpyAsync(centroids, ix.cent.p, (size_t)ix.nlist * ix.dim * sizeof(float), cudaMemcpyDeviceToHost,
                                  h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int solo_ivf_add(solo_handle *h, int charge, const float *x, int64_t n, int dim) {
    if (!h || (!x && n > 0)) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        SOLO_REQUIRE(dim == ix.dim, SOLO_EINVAL, "dim %d does not match the index (%d)", dim, ix.dim);
        DevBuf &xd = h->scratch[19];
        const int64_t chunk = 1 << 17;
        xd.ensure((size_t)std::min<int64_t>(std::max<int64_t>(n, 1), chunk) * dim * sizeof(float));
        for (int64_t r0 = 0; r0 < n; r0 += chunk) {
            int64_t m = std::min(chunk, n - r0);
            SOLO_CUDA(cudaMemcpyAsync(xd.p, x + r0 * dim, (size_t)m * dim * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            ivf_add_device(h, ix, xd.as<float>(), m);
        }
    });
}

int solo_ivf_add_library(solo_handle *h, int charge) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        LibraryStore &L = get_lib(h, charge);
        SOLO_REQUIRE(ix.dim == h->hash_len, SOLO_EINVAL, "index dim %d != hash_len %d", ix.dim, h->hash_len);
        DevBuf &xd = h->scratch[19], &coff = h->scratch[20];
        const int64_t chunk = 1 << 17;
        xd.ensure((size_t)std::min<int64_t>(std::max<int64_t>(L.n, 1), chunk) * ix.dim * sizeof(float));
        std::vector<int64_t> hoff(L.n + 1);
        SOLO_CUDA(cudaMemcpy(hoff.data(), L.off.p, (L.n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
        for (int64_t r0 = 0; r0 < L.n; r0 += chunk) {
            int64_t m = std::min(chunk, L.n - r0);
            // K1 takes offsets relative to the arrays it is handed: the store's offsets are absolute,
            // so hand it the full arrays and the slice of offsets.
            int64_t npk = hoff[r0 + m] - hoff[r0];
            launch_vectorize(h, L.mz.p, 0, L.inten.as<float>(), L.off.as<int64_t>() + r0, m, npk, 1, xd.as<float>(),
                             nullptr, 0);
            ivf_add_device(h, ix, xd.as<float>(), m);
        }
        (void)coff;
    });
}

int solo_ivf_reset(solo_handle *h, int charge) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] { ivf_reset(get_ivf(h, charge, true)); });
}

int solo_ivf_ntotal(solo_handle *h, int charge, int64_t *ntotal, int32_t *nlist, int32_t *dim) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        if (ntotal) *ntotal = ix.ntotal;
        if (nlist) *nlist = ix.nlist;
        if (dim) *dim = ix.dim;
    });
}

int solo_ivf_get_assignment(solo_handle *h, int charge, int32_t *list_of_row) {
    if (!h || !list_of_row) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        if (ix.ntotal == 0) return;
        SOLO_CUDA(cudaMemcpyAsync(list_of_row, ix.row_list.p, ix.ntotal * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int solo_idxann_inspect(const char *path, solo_idxann_info *info, char *errbuf, int errbuf_len) {
    if (errbuf && errbuf_len > 0) errbuf[0] = 0;
    if (!path || !info) return SOLO_EINVAL;
    try {
        idxann_inspect(path, info);
        return SOLO_OK;
    } catch (const Error &e) {
        if (errbuf && errbuf_len > 0) snprintf(errbuf, errbuf_len, "%s", e.msg.c_str());
        return e.code;
    } catch (const std::exception &e) {
        if (errbuf && errbuf_len > 0) snprintf(errbuf, errbuf_len, "%s", e.what());
        return SOLO_EINVAL;
    }
}

int solo_ivf_read_index(solo_handle *h, int charge, const char *path, int64_t *nprobe) {
    if (!h || !path) return SOLO_EINVAL;
    return guarded(h, [&] { ivf_read_index(h, get_ivf(h, charge, false), path, nprobe); });
}

int solo_ivf_write_index(solo_handle *h, int charge, const char *path, int64_t nprobe) {
    if (!h || !path) return SOLO_EINVAL;
    return guarded(h, [&] { ivf_write_index(h, get_ivf(h, charge, true), path, nprobe); });
}

int solo_ivf_add_assigned(solo_handle *h, int charge, const float *x, int64_t n, int dim, const int32_t *list_of_row) {
    if (!h || ((!x || !list_of_row) && n > 0)) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        SOLO_REQUIRE(dim == ix.dim, SOLO_EINVAL, "dim %d does not match the index (%d)", dim, ix.dim);
        if (n > 0) ivf_add_assigned(h, ix, x, n, list_of_row);
    });
}

int solo_ivf_reconstruct(solo_handle *h, int charge, int64_t row0, int64_t n, float *out) {
    if (!h || (!out && n > 0)) return SOLO_EINVAL;
    return guarded(h, [&] { ivf_reconstruct(h, get_ivf(h, charge, true), row0, n, out); });
}

int solo_ivf_search(solo_handle *h, int charge, const float *queries, int nq, int dim, int k, int nprobe, int64_t *I,
                    float *D) {
    if (!h || !queries || !I) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        SOLO_REQUIRE(dim == ix.dim, SOLO_EINVAL, "dim %d does not match the index (%d)", dim, ix.dim);
        SOLO_REQUIRE(nq >= 0, SOLO_EINVAL, "nq < 0");
        if (nq == 0) return;
        DevBuf &dq = h->scratch[19], &dI = h->scratch[20], &dD = h->scratch[21];
        h2d(h, dq, queries, (size_t)nq * dim * sizeof(float));
        dI.ensure((size_t)nq * k * sizeof(int64_t));
        dD.ensure((size_t)nq * k * sizeof(float));
        IvfSearchArgs a;
        memset(&a, 0, sizeof a);
        a.q = dq.as<float>();
        a.nq = nq;
        a.k = k;
        a.nprobe = nprobe;
        a.I = dI.as<int64_t>();
        a.D = dD.as<float>();
        a.win_tol_mode = -1;
        ivf_search(h, ix, a);
        SOLO_CUDA(cudaMemcpyAsync(I, dI.p, (size_t)nq * k * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        if (D) SOLO_CUDA(cudaMemcpyAsync(D, dD.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int solo_debug_scan_dump(solo_handle *h, int charge, int nq, int32_t *cap, int32_t *counts, uint64_t *entries) {
    if (!h || !cap) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        *cap = 32768;
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        if (counts) SOLO_CUDA(cudaMemcpy(counts, h->scratch[15].p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (entries) {
            std::vector<int32_t> c(nq);
            SOLO_CUDA(cudaMemcpy(c.data(), h->scratch[15].p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost));
            std::vector<int32_t> ids(std::max<int64_t>(ix.nstored, 1));
            SOLO_CUDA(cudaMemcpy(ids.data(), ix.list_ids.p, ix.nstored * sizeof(int32_t), cudaMemcpyDeviceToHost));
            for (int q = 0; q < nq; ++q) {
                int n = std::min(c[q], 32768);
                uint64_t *dst = entries + (size_t)q * 32768;
                SOLO_CUDA(cudaMemcpy(dst, h->scratch[17].as<unsigned long long>() + (size_t)q * 32768,
                                     (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
                for (int i = 0; i < n; ++i)  // list position -> library row
                    dst[i] = (dst[i] & 0xFFFFFFFF00000000ull) | (uint32_t)ids[(uint32_t)(dst[i] & 0xFFFFFFFFull)];
            }
        }
    });
}

int solo_ivf_coarse(solo_handle *h, int charge, const float *queries, int nq, int dim, int nprobe, int32_t *probes) {
    if (!h || !queries || !probes) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        SOLO_REQUIRE(dim == ix.dim, SOLO_EINVAL, "dim %d does not match the index (%d)", dim, ix.dim);
        if (nq <= 0) return;
        const int np = std::min(nprobe, ix.nlist);
        DevBuf &dq = h->scratch[19], &dP = h->scratch[20];
        h2d(h, dq, queries, (size_t)nq * dim * sizeof(float));
        dP.ensure((size_t)nq * np * sizeof(int32_t));
        IvfSearchArgs a;
        memset(&a, 0, sizeof a);
        a.q = dq.as<float>();
        a.nq = nq;
        a.k = 1;
        a.nprobe = nprobe;
        a.probes = dP.as<int32_t>();
        a.sort_probes = 1;
        a.coarse_only = 1;
        a.win_tol_mode = -1;
        ivf_search(h, ix, a);
        SOLO_CUDA(cudaMemcpyAsync(probes, dP.p, (size_t)nq * np * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    });
}

// ---------------------------------------------------------------- K5 standalone

static void stage_queries_impl(solo_handle *h, const float *q_mz, const void *q_mz_vec, const float *q_int,
                               const int64_t *q_off, const double *q_prec_mz, int nq, int mz_is_f64,
                               cudaStream_t st = nullptr) {
    SOLO_REQUIRE(nq >= 0 && q_off && (nq == 0 || (q_mz && q_int && q_prec_mz)), SOLO_EINVAL, "null argument");
    SOLO_REQUIRE(q_off[0] == 0, SOLO_EINVAL, "query offsets must start at 0");
    const int64_t np = q_off[nq];
    int maxp = 0;
    for (int i = 0; i < nq; ++i) {
        int64_t len = q_off[i + 1] - q_off[i];
        SOLO_REQUIRE(len >= 0, SOLO_EINVAL, "query offsets must be non-decreasing");
        maxp = std::max<int64_t>(maxp, len);
        // the scorer's monotone advance (SpectrumMatch.cpp:39-46) assumes ascending query m/z
        for (int64_t p = q_off[i] + 1; p < q_off[i + 1]; ++p)
            SOLO_REQUIRE(q_mz[p] >= q_mz[p - 1], SOLO_EINVAL, "query %d: m/z not ascending", i);
    }
    SOLO_REQUIRE(maxp <= 128, SOLO_ECAPACITY, "query spectra may hold at most 128 peaks (got %d)", maxp);
    const double h2d_bytes = (double)(np * (8 + (q_mz_vec ? (mz_is_f64 ? 8 : 4) : 0)) + (nq + 1) * 8 + nq * 8);
    if (st) h->prof[ST_H2D].units += h2d_bytes;   // copy stream: bytes are counted, the time hides under the kernels
    StageTimer t(h, ST_H2D, 0, h2d_bytes, /*enabled=*/st == nullptr);
    h->nq = nq;
    h->q_peaks = np;
    h->q_max_peaks = maxp;
    h->q_mz_is_f64 = mz_is_f64;
    h2d(h, h->q_mz, q_mz, np * 4, st);
    h2d(h, h->q_int, q_int, np * 4, st);
    h2d(h, h->q_off, q_off, (size_t)(nq + 1) * 8, st);
    h2d(h, h->q_prec_mz, q_prec_mz, (size_t)nq * 8, st);
    if (q_mz_vec && (mz_is_f64 || q_mz_vec != (const void *)q_mz)) h2d(h, h->q_mz_vec, q_mz_vec, np * (mz_is_f64 ? 8 : 4), st);
    else h->q_mz_is_f64 = -1;  // binning reads the float32 scorer array
}

static void ensure_results(solo_handle *h, int nq, int max_pairs) {
    h->r_best_row.ensure(std::max(nq, 1) * sizeof(int32_t));
    h->r_best_score.ensure(std::max(nq, 1) * sizeof(double));
    h->r_n_pairs.ensure(std::max(nq, 1) * sizeof(int32_t));
    h->r_n_cand.ensure(std::max(nq, 1) * sizeof(int32_t));
    h->r_pairs.ensure((size_t)std::max(nq, 1) * max_pairs * 2 * sizeof(uint32_t));
    h->r_nq = nq;
    h->r_max_pairs = max_pairs;
}

int solo_best_match_batch(solo_handle *h, int charge, const float *q_mz, const float *q_intensity, const int64_t *q_off,
                          const double *q_prec_mz, int nq, const int32_t *cand_ids, const int64_t *cand_off,
                          double fragment_mz_tolerance, int allow_shift, int max_pairs, int32_t *best_pos,
                          double *best_score, int32_t *n_pairs, uint32_t *pairs) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        LibraryStore &L = get_lib(h, charge);
        SOLO_REQUIRE(cand_off && best_pos && best_score && n_pairs && pairs && max_pairs > 0, SOLO_EINVAL, "null argument");
        if (nq == 0) return;
        stage_queries_impl(h, q_mz, nullptr, q_intensity, q_off, q_prec_mz, nq, 0);
        const int64_t nc = cand_off[nq];
        for (int64_t i = 0; i < nc; ++i)
            SOLO_REQUIRE(cand_ids[i] >= 0 && cand_ids[i] < L.n, SOLO_EINVAL, "candidate id %d out of range", cand_ids[i]);
        DevBuf &dci = h->scratch[0], &dco = h->scratch[1], &dpos = h->scratch[2], &ovf = h->scratch[3];
        h2d(h, dci, cand_ids, nc * sizeof(int32_t));
        h2d(h, dco, cand_off, (size_t)(nq + 1) * sizeof(int64_t));
        dpos.ensure((size_t)nq * sizeof(int32_t));
        ovf.ensure(16);
        SOLO_CUDA(cudaMemsetAsync(ovf.p, 0, 16, h->stream));
        ensure_results(h, nq, max_pairs);
        ScoreArgs a;
        a.q_mz = h->q_mz.as<float>();
        a.q_int = h->q_int.as<float>();
        a.q_off = h->q_off.as<int64_t>();
        a.q_prec_mz = h->q_prec_mz.as<double>();
        a.nq = nq;
        a.q_max_peaks = h->q_max_peaks;
        a.lib = &L;
        a.cand_ids = dci.as<int32_t>();
        a.cand_off = dco.as<int64_t>();
        a.tol = fragment_mz_tolerance;
        a.allow_shift = allow_shift;
        a.max_pairs = max_pairs;
        a.best_pos = dpos.as<int32_t>();
        a.best_row = nullptr;
        a.best_score = h->r_best_score.as<double>();
        a.n_pairs = h->r_n_pairs.as<int32_t>();
        a.pairs = h->r_pairs.as<uint32_t>();
        a.overflow = ovf.as<int32_t>();
        launch_best_match(h, a);
        int32_t n_over = 0;
        SOLO_CUDA(cudaMemcpyAsync(best_pos, dpos.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaMemcpyAsync(best_score, h->r_best_score.p, (size_t)nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaMemcpyAsync(n_pairs, h->r_n_pairs.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaMemcpyAsync(pairs, h->r_pairs.p, (size_t)nq * max_pairs * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaMemcpyAsync(&n_over, ovf.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        h->k5_overflow_pairs += n_over;  // pairs that took the exact re-enumeration path (statistics, not an error)
    });
}

// ---------------------------------------------------------------- fused open search

static void slot_fields(solo_handle *h, DevBuf **f) {
    DevBuf *all[11] = {&h->q_mz, &h->q_mz_vec, &h->q_int, &h->q_off, &h->q_prec_mz, &h->r_best_row,
                       &h->r_best_score, &h->r_n_pairs, &h->r_pairs, &h->r_n_cand, &h->r_ovf};
    for (int i = 0; i < 11; ++i) f[i] = all[i];
}

int solo_select_slot(solo_handle *h, int slot) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        if (slot == h->active_slot) return;
        DevBuf *f[11];
        slot_fields(h, f);
        solo_handle::Slot &out = h->parked[h->active_slot];
        out.nq = h->nq; out.q_peaks = h->q_peaks; out.q_max_peaks = h->q_max_peaks; out.q_mz_is_f64 = h->q_mz_is_f64;
        out.r_nq = h->r_nq; out.r_max_pairs = h->r_max_pairs;
        for (int i = 0; i < 11; ++i) out.b[i] = *f[i];
        solo_handle::Slot in = h->parked[slot];  // empty when new
        h->parked.erase(slot);
        h->nq = in.nq; h->q_peaks = in.q_peaks; h->q_max_peaks = in.q_max_peaks; h->q_mz_is_f64 = in.q_mz_is_f64;
        h->r_nq = in.r_nq; h->r_max_pairs = in.r_max_pairs;
        for (int i = 0; i < 11; ++i) *f[i] = in.b[i];
        h->active_slot = slot;
    });
}

int solo_stage_queries(solo_handle *h, const float *q_mz, const void *q_mz_vec, const float *q_intensity,
                       const int64_t *q_off, const double *q_prec_mz, int nq, int mz_is_f64) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] { stage_queries_impl(h, q_mz, q_mz_vec ? q_mz_vec : q_mz, q_intensity, q_off, q_prec_mz, nq, mz_is_f64); });
}

int solo_search_staged(solo_handle *h, int charge, const solo_search_params *p) {
    if (!h || !p) return SOLO_EINVAL;
    return guarded(h, [&] {
        LibraryStore &L = get_lib(h, charge);
        const int nq = h->nq;
        SOLO_REQUIRE(p->tol_mode == SOLO_TOL_DA || p->tol_mode == SOLO_TOL_PPM, SOLO_EINVAL,
                     "Unknown precursor tolerance mode");  // spectral_library.py:429
        SOLO_REQUIRE(p->max_pairs > 0, SOLO_EINVAL, "max_pairs must be positive");
        ensure_results(h, nq, p->max_pairs);
        if (nq == 0) return;
        // an asynchronously staged batch (solo_stage_queries_async) lands on the copy stream, and the slot's previous
        // results may still be on their way to the host (solo_fetch_results_async): the kernels wait for both; the
        // `done` event recorded behind the last kernel is what later copies of this slot wait for
        auto sync_it = h->slot_sync.find(h->active_slot);
        if (sync_it != h->slot_sync.end()) {
            if (sync_it->second.wait_staged) SOLO_CUDA(cudaStreamWaitEvent(h->stream, sync_it->second.staged, 0));
            if (sync_it->second.has_fetched) SOLO_CUDA(cudaStreamWaitEvent(h->stream, sync_it->second.fetched, 0));
            sync_it->second.wait_staged = false;
        }
        struct DoneMark {
            solo_handle *h;
            ~DoneMark() {
                auto it = h->slot_sync.find(h->active_slot);
                if (it != h->slot_sync.end() && cudaEventRecord(it->second.done, h->stream) == cudaSuccess) it->second.has_done = true;
            }
        } done_mark{h};
        DevBuf &ovf = h->r_ovf;
        ovf.ensure(16);
        SOLO_CUDA(cudaMemsetAsync(ovf.p, 0, 16, h->stream));
        ScoreArgs a;
        a.q_mz = h->q_mz.as<float>();
        a.q_int = h->q_int.as<float>();
        a.q_off = h->q_off.as<int64_t>();
        a.q_prec_mz = h->q_prec_mz.as<double>();
        a.nq = nq;
        a.q_max_peaks = h->q_max_peaks;
        a.lib = &L;
        a.tol = p->fragment_mz_tolerance;
        a.allow_shift = p->allow_shift;
        a.max_pairs = p->max_pairs;
        DevBuf &dpos = h->scratch[23];
        dpos.ensure((size_t)nq * sizeof(int32_t));
        a.best_pos = dpos.as<int32_t>();
        a.best_row = h->r_best_row.as<int32_t>();
        a.best_score = h->r_best_score.as<double>();
        a.n_pairs = h->r_n_pairs.as<int32_t>();
        a.pairs = h->r_pairs.as<uint32_t>();
        a.overflow = ovf.as<int32_t>();
        a.tie_by_row = 1;
        const bool ann = p->use_ann && h->ivf.count(charge) && h->ivf[charge].nlist > 0;
        if (ann) {
            IvfIndex &ix = h->ivf[charge];
            SOLO_REQUIRE(ix.dim == h->hash_len, SOLO_EINVAL, "index dim %d != hash_len %d", ix.dim, h->hash_len);
            SOLO_REQUIRE(p->k >= 1 && p->k <= IVF_MAX_K, SOLO_EINVAL, "num_candidates must be in [1, %d]", IVF_MAX_K);
            // index row i must be library row i (spectral_library.py:443-451): a stale or foreign index would
            // otherwise address the peak store out of bounds
            SOLO_REQUIRE(ix.ntotal == L.n, SOLO_ESTATE, "ANN index of charge %d holds %lld rows but the library has %lld",
                         charge, (long long)ix.ntotal, (long long)L.n);
            // K1: query vectors
            DevBuf &qv = h->scratch[19], &sel = h->scratch[20];
            qv.ensure((size_t)nq * h->hash_len * sizeof(float));
            const void *mzv = h->q_mz_is_f64 < 0 ? h->q_mz.p : h->q_mz_vec.p;
            launch_vectorize(h, mzv, h->q_mz_is_f64 > 0 ? 1 : 0, h->q_int.as<float>(), h->q_off.as<int64_t>(), nq,
                             h->q_peaks, 1, qv.as<float>(), nullptr, 0);
            sel.ensure((size_t)nq * p->k * sizeof(int32_t));
            IvfSearchArgs s;
            memset(&s, 0, sizeof s);
            s.q = qv.as<float>();
            s.nq = nq;
            s.k = p->k;
            s.nprobe = p->nprobe;
            s.sel_ids = sel.as<int32_t>();
            s.sel_cnt = h->r_n_cand.as<int32_t>();
            s.win_q_prec_mz = h->q_prec_mz.as<double>();
            s.win_lib_prec_mz32 = L.prec_mz32.as<float>();
            s.win_lib_valid = L.valid.as<uint8_t>();
            s.win_charge = charge;
            s.win_tol = p->tol_value;
            s.win_tol_mode = p->tol_mode;
            ivf_search(h, ix, s);
            a.cand_ids = sel.as<int32_t>();
            a.cand_off = nullptr;
            a.cand_cnt = h->r_n_cand.as<int32_t>();
            a.cand_stride = p->k;
        } else {
            // brute force: every valid library row inside the precursor window
            DevBuf &coff = h->scratch[20], &cids = h->scratch[21];
            coff.ensure((size_t)(nq + 1) * sizeof(int64_t));
            int64_t total = 0;
            {
                StageTimer t(h, ST_CANDIDATES, 3);
                static const bool v_scan_all = getenv("SOLO_BF_SCAN_ALL") != nullptr;  // cross-check: the O(N) kernel
                if (v_scan_all)
                    window_candidates_kernel<<<nq, 256, 0, h->stream>>>(h->q_prec_mz.as<double>(), L.prec_mz32.as<float>(),
                                                                        L.valid.as<uint8_t>(), L.n, charge, p->tol_value,
                                                                        p->tol_mode, h->r_n_cand.as<int32_t>(), nullptr, nullptr);
                else
                    window_candidates_sorted_kernel<<<nq, 128, 0, h->stream>>>(
                        h->q_prec_mz.as<double>(), L.sorted_mz32.as<float>(), L.sorted_row.as<int32_t>(),
                        L.valid.as<uint8_t>(), L.n, charge, p->tol_value, p->tol_mode, h->r_n_cand.as<int32_t>(), nullptr,
                        nullptr);
                SOLO_CUDA(cudaGetLastError());
                // offsets = exclusive scan of the counts (single-CTA scan kernel lives in ivf.cu)
                scan_counts_i32(h, h->r_n_cand.as<int32_t>(), nq, coff.as<int64_t>());
                SOLO_CUDA(cudaMemcpyAsync(&total, coff.as<int64_t>() + nq, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
                SOLO_CUDA(cudaStreamSynchronize(h->stream));
                cids.ensure(std::max<int64_t>(total, 1) * sizeof(int32_t));
                if (v_scan_all)
                    window_candidates_kernel<<<nq, 256, 0, h->stream>>>(h->q_prec_mz.as<double>(), L.prec_mz32.as<float>(),
                                                                        L.valid.as<uint8_t>(), L.n, charge, p->tol_value,
                                                                        p->tol_mode, nullptr, coff.as<int64_t>(), cids.as<int32_t>());
                else
                    window_candidates_sorted_kernel<<<nq, 128, 0, h->stream>>>(
                        h->q_prec_mz.as<double>(), L.sorted_mz32.as<float>(), L.sorted_row.as<int32_t>(),
                        L.valid.as<uint8_t>(), L.n, charge, p->tol_value, p->tol_mode, nullptr, coff.as<int64_t>(),
                        cids.as<int32_t>());
                SOLO_CUDA(cudaGetLastError());
            }
            a.cand_ids = cids.as<int32_t>();
            a.cand_off = coff.as<int64_t>();
        }
        launch_best_match(h, a);
    });
}

int solo_fetch_results(solo_handle *h, int32_t *best_row, double *best_score, int32_t *n_pairs, uint32_t *pairs,
                       int32_t *n_cand) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        const int nq = h->r_nq;
        if (nq == 0) return;
        int32_t n_over = 0;
        {
            StageTimer t(h, ST_D2H, 0, (double)nq * (4 + 8 + 4 + 4 + (double)h->r_max_pairs * 8));
            auto cp = [&](void *dst, const DevBuf &src, size_t bytes) {
                if (dst) SOLO_CUDA(cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, h->stream));
            };
            cp(best_row, h->r_best_row, (size_t)nq * 4);
            cp(best_score, h->r_best_score, (size_t)nq * 8);
            cp(n_pairs, h->r_n_pairs, (size_t)nq * 4);
            cp(pairs, h->r_pairs, (size_t)nq * h->r_max_pairs * 8);
            cp(n_cand, h->r_n_cand, (size_t)nq * 4);
            cp(&n_over, h->r_ovf, 4);
        }
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        h->k5_overflow_pairs += n_over;  // pairs that took the exact re-enumeration path (statistics, not an error)
    });
}

int solo_fetch_results_range(solo_handle *h, int q_begin, int n, int32_t *best_row, double *best_score, int32_t *n_pairs,
                             uint32_t *pairs, int32_t *n_cand) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(q_begin >= 0 && n >= 0 && q_begin + n <= h->r_nq, SOLO_EINVAL, "result rows [%d, %d) out of range (%d staged)",
                     q_begin, q_begin + n, h->r_nq);
        if (n == 0) return;
        int32_t n_over = 0;
        const size_t mp = (size_t)h->r_max_pairs;
        {
            StageTimer t(h, ST_D2H, 0, (double)n * (4 + 8 + 4 + 4 + (double)mp * 8));
            auto cp = [&](void *dst, const DevBuf &src, size_t elem) {
                if (dst) SOLO_CUDA(cudaMemcpyAsync(dst, src.as<unsigned char>() + (size_t)q_begin * elem, (size_t)n * elem,
                                                   cudaMemcpyDeviceToHost, h->stream));
            };
            cp(best_row, h->r_best_row, 4);
            cp(best_score, h->r_best_score, 8);
            cp(n_pairs, h->r_n_pairs, 4);
            cp(pairs, h->r_pairs, mp * 8);
            cp(n_cand, h->r_n_cand, 4);
            SOLO_CUDA(cudaMemcpyAsync(&n_over, h->r_ovf.p, 4, cudaMemcpyDeviceToHost, h->stream));
        }
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        h->k5_overflow_pairs += n_over;
    });
}

int solo_search_batch(solo_handle *h, int charge, const solo_search_params *p, const float *q_mz, const void *q_mz_vec,
                      const float *q_intensity, const int64_t *q_off, const double *q_prec_mz, int nq,
                      int32_t *best_row, double *best_score, int32_t *n_pairs, uint32_t *pairs, int32_t *n_cand) {
    int rc = solo_stage_queries(h, q_mz, q_mz_vec, q_intensity, q_off, q_prec_mz, nq, p ? p->mz_is_f64 : 0);
    if (rc) return rc;
    rc = solo_search_staged(h, charge, p);
    if (rc) return rc;
    return solo_fetch_results(h, best_row, best_score, n_pairs, pairs, n_cand);
}

// ---- streaming: copies of batch i+1 / i-1 under the kernels of batch i ------------------------------------------
int solo_reserve_slot(solo_handle *h, int nq, int64_t n_peaks, int max_pairs, int mz_is_f64) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(nq >= 0 && n_peaks >= 0 && max_pairs >= 0, SOLO_EINVAL, "negative size");
        // growing a buffer frees and allocates (a device-wide synchronisation): done here, ahead of the stream
        h->q_mz.ensure(std::max<size_t>(n_peaks * 4, 16));
        h->q_int.ensure(std::max<size_t>(n_peaks * 4, 16));
        h->q_off.ensure((size_t)(nq + 1) * 8);
        h->q_prec_mz.ensure(std::max<size_t>((size_t)nq * 8, 16));
        if (mz_is_f64) h->q_mz_vec.ensure(std::max<size_t>(n_peaks * 8, 16));
        h->r_best_row.ensure(std::max(nq, 1) * sizeof(int32_t));
        h->r_best_score.ensure(std::max(nq, 1) * sizeof(double));
        h->r_n_pairs.ensure(std::max(nq, 1) * sizeof(int32_t));
        h->r_n_cand.ensure(std::max(nq, 1) * sizeof(int32_t));
        h->r_pairs.ensure((size_t)std::max(nq, 1) * std::max(max_pairs, 1) * 2 * sizeof(uint32_t));
        h->r_ovf.ensure(16);
    });
}

int solo_stage_queries_async(solo_handle *h, const float *q_mz, const void *q_mz_vec, const float *q_intensity,
                             const int64_t *q_off, const double *q_prec_mz, int nq, int mz_is_f64) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        solo_handle::SlotSync &s = slot_sync(h, h->active_slot);
        // the slot's previous batch must have been read by its kernels before its buffers are overwritten
        if (s.has_done) SOLO_CUDA(cudaStreamWaitEvent(h->copy_stream, s.done, 0));
        stage_queries_impl(h, q_mz, q_mz_vec ? q_mz_vec : q_mz, q_intensity, q_off, q_prec_mz, nq, mz_is_f64, h->copy_stream);
        SOLO_CUDA(cudaEventRecord(s.staged, h->copy_stream));
        s.wait_staged = true;
    });
}

int solo_fetch_results_async(solo_handle *h, int32_t *best_row, double *best_score, int32_t *n_pairs, uint32_t *pairs,
                             int32_t *n_cand) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        solo_handle::SlotSync &s = slot_sync(h, h->active_slot);
        const int nq = h->r_nq;
        if (!s.has_done) {   // nothing was searched through the event-marked path: order behind the compute stream
            SOLO_CUDA(cudaEventRecord(s.done, h->stream));
            s.has_done = true;
        }
        SOLO_CUDA(cudaStreamWaitEvent(h->copy_stream, s.done, 0));
        auto cp = [&](void *dst, const DevBuf &src, size_t bytes) {
            if (dst && bytes) SOLO_CUDA(cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, h->copy_stream));
        };
        cp(best_row, h->r_best_row, (size_t)nq * 4);
        cp(best_score, h->r_best_score, (size_t)nq * 8);
        cp(n_pairs, h->r_n_pairs, (size_t)nq * 4);
        cp(pairs, h->r_pairs, (size_t)nq * h->r_max_pairs * 8);
        cp(n_cand, h->r_n_cand, (size_t)nq * 4);
        if (nq) cp(s.n_over, h->r_ovf, 4);
        h->prof[ST_D2H].units += (double)nq * (4 + 8 + 4 + 4 + (double)h->r_max_pairs * 8);
        SOLO_CUDA(cudaEventRecord(s.fetched, h->copy_stream));
        s.has_fetched = true;
    });
}

int solo_wait_results(solo_handle *h, int slot) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        auto it = h->slot_sync.find(slot);
        SOLO_REQUIRE(it != h->slot_sync.end() && it->second.has_fetched, SOLO_ESTATE, "slot %d has no fetch in flight", slot);
        SOLO_CUDA(cudaEventSynchronize(it->second.fetched));
        h->k5_overflow_pairs += *it->second.n_over;
        *it->second.n_over = 0;
    });
}

// ---- K0 / .splib: ingestion -----------------------------------------------------------------
int solo_splib_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_peptide_bytes, char *errbuf,
                     int errbuf_len) {
    if (!path || !n_spectra || !n_peaks || !n_peptide_bytes) return SOLO_EINVAL;
    return guarded_nohandle(errbuf, errbuf_len, [&] { splib_count(path, n_spectra, n_peaks, n_peptide_bytes); });
}

int solo_splib_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t n_peptide_bytes, uint32_t *identifier,
                    double *prec_mz, int32_t *prec_charge, uint8_t *is_decoy, int64_t *file_offset,
                    int64_t *peak_offsets, float *mz, float *intensity, uint8_t *peak_charge, int64_t *peptide_offsets,
                    char *peptides, char *errbuf, int errbuf_len) {
    if (!path || !identifier || !prec_mz || !prec_charge || !is_decoy || !file_offset || !peak_offsets || !mz ||
        !intensity || !peak_charge || !peptide_offsets || !peptides)
        return SOLO_EINVAL;
    return guarded_nohandle(errbuf, errbuf_len, [&] {
        splib_read(path, n_spectra, n_peaks, n_peptide_bytes, identifier, prec_mz, prec_charge, is_decoy, file_offset,
                   peak_offsets, mz, intensity, peak_charge, peptide_offsets, peptides);
    });
}

int solo_mgf_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_identifier_bytes,
                   int64_t *n_seq_bytes, char *errbuf, int errbuf_len) {
    if (!path || !n_spectra || !n_peaks || !n_identifier_bytes || !n_seq_bytes) return SOLO_EINVAL;
    return guarded_nohandle(errbuf, errbuf_len,
                            [&] { mgf_count(path, n_spectra, n_peaks, n_identifier_bytes, n_seq_bytes); });
}

int solo_mgf_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t n_identifier_bytes, int64_t n_seq_bytes,
                  double *prec_mz, int32_t *prec_charge, double *rt_seconds, uint8_t *is_decoy, int64_t *peak_offsets,
                  double *mz, float *intensity, int64_t *identifier_offsets, char *identifiers, int64_t *seq_offsets,
                  char *seqs, char *errbuf, int errbuf_len) {
    if (!path || !prec_mz || !prec_charge || !rt_seconds || !is_decoy || !peak_offsets || !mz || !intensity ||
        !identifier_offsets || !identifiers || !seq_offsets || !seqs)
        return SOLO_EINVAL;
    return guarded_nohandle(errbuf, errbuf_len, [&] {
        mgf_read(path, n_spectra, n_peaks, n_identifier_bytes, n_seq_bytes, prec_mz, prec_charge, rt_seconds, is_decoy,
                 peak_offsets, mz, intensity, identifier_offsets, identifiers, seq_offsets, seqs);
    });
}

int solo_mzml_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_skipped, char *errbuf,
                    int errbuf_len) {
    if (!path || !n_spectra || !n_peaks || !n_skipped) return SOLO_EINVAL;
    return guarded_nohandle(errbuf, errbuf_len, [&] { mzml_count(path, n_spectra, n_peaks, n_skipped); });
}

int solo_mzml_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t *scan_nr, int32_t *index, double *prec_mz,
                   int32_t *prec_charge, double *rt, int64_t *peak_offsets, double *mz, float *intensity, char *errbuf,
                   int errbuf_len) {
    if (!path || !scan_nr || !index || !prec_mz || !prec_charge || !rt || !peak_offsets || !mz || !intensity)
        return SOLO_EINVAL;
    return guarded_nohandle(errbuf, errbuf_len, [&] {
        mzml_read(path, n_spectra, n_peaks, scan_nr, index, prec_mz, prec_charge, rt, peak_offsets, mz, intensity);
    });
}

int solo_mzxml_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_skipped, char *errbuf,
                     int errbuf_len) {
    if (!path || !n_spectra || !n_peaks || !n_skipped) return SOLO_EINVAL;
    return guarded_nohandle(errbuf, errbuf_len, [&] { mzxml_count(path, n_spectra, n_peaks, n_skipped); });
}

int solo_mzxml_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t *scan_nr, int32_t *index, double *prec_mz,
                    int32_t *prec_charge, double *rt, int64_t *peak_offsets, double *mz, float *intensity, char *errbuf,
                    int errbuf_len) {
    if (!path || !scan_nr || !index || !prec_mz || !prec_charge || !rt || !peak_offsets || !mz || !intensity)
        return SOLO_EINVAL;
    return guarded_nohandle(errbuf, errbuf_len, [&] {
        mzxml_read(path, n_spectra, n_peaks, scan_nr, index, prec_mz, prec_charge, rt, peak_offsets, mz, intensity);
    });
}

int solo_process_spectra(solo_handle *h, const void *mz, int mz_is_f64, const float *intensity, const int64_t *offsets,
                         const double *prec_mz, const int32_t *prec_charge, int64_t n, const solo_process_params *p,
                         void *out_mz, float *out_intensity, int32_t *out_index, int32_t *out_count, uint8_t *out_valid) {
    if (!h || !p || n < 0 || (n > 0 && (!mz || !intensity || !offsets || !out_mz || !out_intensity || !out_index ||
                                        !out_count || !out_valid)))
        return SOLO_EINVAL;
    return guarded(h, [&] {
        if (n == 0) return;
        SOLO_REQUIRE(p->max_peaks >= 1 && p->max_peaks <= 128, SOLO_ECAPACITY, "max_peaks must be in [1, 128] (got %d)",
                     p->max_peaks);
        SOLO_REQUIRE(p->scaling >= SOLO_SCALING_NONE && p->scaling <= SOLO_SCALING_RANK, SOLO_EINVAL,
                     "Unknown intensity scaling");
        SOLO_REQUIRE(!p->remove_precursor || (prec_mz && prec_charge), SOLO_EINVAL,
                     "remove_precursor needs precursor m/z and charge");
        SOLO_REQUIRE(p->resolution >= -1 && p->resolution <= 12, SOLO_EINVAL, "resolution must be -1 (none) or 0..12 decimals");
        SOLO_REQUIRE(n < (int64_t)0x7fffffff && offsets[0] == 0, SOLO_EINVAL, "bad offsets");
        const int64_t npk = offsets[n];
        const size_t esz = mz_is_f64 ? 8 : 4;
        for (int64_t i = 0; i < n; ++i) {
            SOLO_REQUIRE(offsets[i + 1] >= offsets[i], SOLO_EINVAL, "offsets must be non-decreasing");
            SOLO_REQUIRE(offsets[i + 1] - offsets[i] <= 8192, SOLO_ECAPACITY,
                         "spectrum %lld holds %lld raw peaks; the kernel stages at most 8192", (long long)i,
                         (long long)(offsets[i + 1] - offsets[i]));
        }
        DevBuf &dmz = h->scratch[21], &din = h->scratch[22], &doff = h->scratch[23], &dpm = h->scratch[24],
               &dz = h->scratch[25], &omz = h->scratch[26], &oin = h->scratch[27], &oidx = h->scratch[28],
               &ocnt = h->scratch[29], &oval = h->scratch[30], &err = h->scratch[31];
        h2d(h, dmz, mz, (size_t)npk * esz);
        h2d(h, din, intensity, (size_t)npk * 4);
        h2d(h, doff, offsets, (size_t)(n + 1) * 8);
        if (prec_mz) h2d(h, dpm, prec_mz, (size_t)n * 8);
        if (prec_charge) h2d(h, dz, prec_charge, (size_t)n * 4);
        const size_t rows = (size_t)n * p->max_peaks;
        omz.ensure(rows * esz);
        oin.ensure(rows * 4);
        oidx.ensure(rows * 4);
        ocnt.ensure((size_t)n * 4);
        oval.ensure((size_t)n);
        err.ensure(4);
        SOLO_CUDA(cudaMemsetAsync(err.p, 0, 4, h->stream));
        SOLO_CUDA(cudaMemsetAsync(omz.p, 0, rows * esz, h->stream));
        SOLO_CUDA(cudaMemsetAsync(oin.p, 0, rows * 4, h->stream));
        SOLO_CUDA(cudaMemsetAsync(oidx.p, 0xff, rows * 4, h->stream));
        ProcessArgs a;
        a.mz = dmz.p;
        a.inten = din.as<float>();
        a.off = doff.as<int64_t>();
        a.prec_mz = prec_mz ? dpm.as<double>() : nullptr;
        a.prec_charge = prec_charge ? dz.as<int32_t>() : nullptr;
        a.n = (int)n;
        a.p = *p;
        a.out_mz = omz.p;
        a.out_int = oin.as<float>();
        a.out_idx = oidx.as<int32_t>();
        a.out_cnt = ocnt.as<int32_t>();
        a.out_valid = oval.as<uint8_t>();
        a.err = err.as<int32_t>();
        launch_process(h, a, mz_is_f64);
        int32_t n_err = 0;
        auto d2h = [&](void *dst, const DevBuf &src, size_t bytes) {
            SOLO_CUDA(cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, h->stream));
        };
        d2h(out_mz, omz, rows * esz);
        d2h(out_intensity, oin, rows * 4);
        d2h(out_index, oidx, rows * 4);
        d2h(out_count, ocnt, (size_t)n * 4);
        d2h(out_valid, oval, (size_t)n);
        d2h(&n_err, err, 4);
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        SOLO_REQUIRE(n_err == 0, SOLO_ECAPACITY, "%d spectra exceeded the raw-peak capacity", n_err);
    });
}

// ---- K6: SSM feature table ----------------------------------------------------------------
static const char *kFeatureNames[SOLO_N_SSM_FEATURES] = {
    "sequence_len", "precursor_charge_2", "precursor_charge_3", "precursor_charge_4", "precursor_charge_5",
    "query_prec_mz", "lib_prec_mz", "mz_diff_ppm", "abs_mz_diff_ppm", "mz_diff_da", "abs_mz_diff_da", "cosine",
    "cosine_top5", "n_matched_peaks", "frac_n_peaks_query", "frac_n_peaks_lib", "frac_n_peaks_lib_top5",
    "frac_int_query", "frac_int_lib", "frac_int_lib_top5", "mse_mz", "mse_mz_top5", "mse_int", "mse_int_top5",
    "contrast_angle", "contrast_angle_top5", "hypergeometric_score", "kendalltau", "ms_for_id_v1", "ms_for_id_v2",
    "entropy_unweighted", "entropy_weighted", "scribe_fragment_acc", "scribe_fragment_acc_top5", "manhattan",
    "euclidean", "chebyshev", "pearsonr", "pearsonr_top5", "spearmanr", "spearmanr_top5", "braycurtis", "canberra",
    "ruzicka"};

const char *solo_ssm_feature_name(int column) {
    return column >= 0 && column < SOLO_N_SSM_FEATURES ? kFeatureNames[column] : "";
}

// device-side arguments are complete except out/bad; runs the kernel and copies the table to the host
static void run_features(solo_handle *h, FeatureArgs &a, const int32_t *h_q_charge, const int32_t *h_seq_len,
                         double *h_out) {
    static_assert(SOLO_N_SSM_FEATURES == k6::N_FEATURES, "header and kernel disagree on the column count");
    const int n = a.n;
    if (n <= 0) return;
    DevBuf &out = h->scratch[21], &zc = h->scratch[22], &sl = h->scratch[23], &bad = h->scratch[24];
    out.ensure((size_t)n * k6::N_FEATURES * sizeof(double));
    bad.ensure(sizeof(int32_t));
    SOLO_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int32_t), h->stream));
    a.q_charge = nullptr;
    a.sequence_len = nullptr;
    if (h_q_charge) {
        h2d(h, zc, h_q_charge, (size_t)n * sizeof(int32_t));
        a.q_charge = zc.as<int32_t>();
    }
    if (h_seq_len) {
        h2d(h, sl, h_seq_len, (size_t)n * sizeof(int32_t));
        a.sequence_len = sl.as<int32_t>();
    }
    a.n_peak_bins = h->n_bins;   // get_dim(config.min_mz, config.max_mz, config.bin_size) (utils.py:398-404)
    {   // log-gamma tables for the binomials of hypergeometric_score
        std::vector<double> tab(k6::LFACT_LEN + k6::LBIG_LEN);
        k6::fill_log_tables(h->n_bins, tab.data(), tab.data() + k6::LFACT_LEN);
        DevBuf &dt = h->scratch[18];
        h2d(h, dt, tab.data(), tab.size() * sizeof(double));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));  // `tab` is pageable and dies with this scope
        a.lfact = dt.as<double>();
        a.lbig = dt.as<double>() + k6::LFACT_LEN;
    }
    a.out = out.as<double>();
    a.bad = bad.as<int32_t>();
    launch_ssm_features(h, a);
    int32_t n_bad = 0;
    SOLO_CUDA(cudaMemcpyAsync(&n_bad, bad.p, sizeof n_bad, cudaMemcpyDeviceToHost, h->stream));
    SOLO_CUDA(cudaMemcpyAsync(h_out, out.p, (size_t)n * k6::N_FEATURES * sizeof(double), cudaMemcpyDeviceToHost,
                              h->stream));
    SOLO_CUDA(cudaStreamSynchronize(h->stream));
    SOLO_REQUIRE(n_bad == 0, SOLO_ECAPACITY,
                 "%d SSMs hold more than %d peaks, more pairs than max_pairs, or a pair outside its spectra", n_bad,
                 k6::MAX_PEAKS);
}

static void fill_library(FeatureArgs &a, const LibraryStore &L) {
    a.l_mz = L.mz.as<float>();
    a.l_int = L.inten.as<float>();
    a.l_off = L.off.as<int64_t>();
    a.l_prec_mz = L.prec_mz.as<double>();
}

int solo_ssm_features(solo_handle *h, int charge, const void *q_mz, int q_mz_is_f64, const float *q_intensity,
                      const int64_t *q_off, const double *q_prec_mz, const int32_t *q_prec_charge, int n_ssm,
                      const int32_t *lib_row, const uint32_t *pairs, const int32_t *n_pairs, int max_pairs,
                      const int32_t *sequence_len, double *out) {
    if (!h || n_ssm < 0 || (n_ssm > 0 && (!q_mz || !q_intensity || !q_off || !q_prec_mz || !lib_row || !pairs ||
                                          !n_pairs || !out)))
        return SOLO_EINVAL;
    return guarded(h, [&] {
        if (n_ssm == 0) return;
        LibraryStore &L = get_lib(h, charge);
        SOLO_REQUIRE(max_pairs >= 1, SOLO_EINVAL, "max_pairs must be >= 1");
        SOLO_REQUIRE(q_off[0] == 0, SOLO_EINVAL, "query offsets must start at 0");
        const int64_t npk = q_off[n_ssm];
        for (int i = 0; i < n_ssm; ++i) {
            SOLO_REQUIRE(lib_row[i] < L.n, SOLO_EINVAL, "SSM %d: library row %d outside the store of charge %d (%lld rows)",
                         i, lib_row[i], charge, (long long)L.n);
            const int64_t len = q_off[i + 1] - q_off[i];
            SOLO_REQUIRE(len >= 0, SOLO_EINVAL, "query offsets must be non-decreasing");
            if (lib_row[i] < 0) continue;
            SOLO_REQUIRE(n_pairs[i] <= max_pairs, SOLO_EINVAL, "SSM %d: %d pairs > max_pairs %d", i, n_pairs[i], max_pairs);
            for (int k = 0; k < n_pairs[i]; ++k)
                SOLO_REQUIRE((int64_t)pairs[((int64_t)i * max_pairs + k) * 2] < len, SOLO_EINVAL,
                             "SSM %d: pair %d names query peak %u of %lld", i, k, pairs[((int64_t)i * max_pairs + k) * 2],
                             (long long)len);
        }
        DevBuf &dmz = h->scratch[25], &din = h->scratch[26], &doff = h->scratch[27], &dpm = h->scratch[28],
               &drow = h->scratch[29], &dpairs = h->scratch[30], &dnp = h->scratch[31];
        h2d(h, dmz, q_mz, (size_t)npk * (q_mz_is_f64 ? 8 : 4));
        h2d(h, din, q_intensity, (size_t)npk * 4);
        h2d(h, doff, q_off, (size_t)(n_ssm + 1) * 8);
        h2d(h, dpm, q_prec_mz, (size_t)n_ssm * 8);
        h2d(h, drow, lib_row, (size_t)n_ssm * 4);
        h2d(h, dpairs, pairs, (size_t)n_ssm * max_pairs * 8);
        h2d(h, dnp, n_pairs, (size_t)n_ssm * 4);
        FeatureArgs a;
        memset(&a, 0, sizeof a);
        a.q_mz32 = q_mz_is_f64 ? nullptr : dmz.as<float>();
        a.q_mz64 = q_mz_is_f64 ? dmz.as<double>() : nullptr;
        a.q_int = din.as<float>();
        a.q_off = doff.as<int64_t>();
        a.q_prec_mz = dpm.as<double>();
        a.q_charge_all = charge;
        fill_library(a, L);
        a.lib_row = drow.as<int32_t>();
        a.pairs = dpairs.as<uint32_t>();
        a.n_pairs = dnp.as<int32_t>();
        a.max_pairs = max_pairs;
        a.n = n_ssm;
        run_features(h, a, q_prec_charge, sequence_len, out);
    });
}

int solo_ssm_features_staged(solo_handle *h, int charge, const int32_t *q_prec_charge, const int32_t *sequence_len,
                             double *out) {
    if (!h || !out) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(h->r_nq == h->nq && h->nq > 0, SOLO_ESTATE, "no search results staged in the active slot");
        LibraryStore &L = get_lib(h, charge);
        FeatureArgs a;
        memset(&a, 0, sizeof a);
        a.q_mz32 = h->q_mz.as<float>();
        a.q_mz64 = h->q_mz_is_f64 > 0 ? h->q_mz_vec.as<double>() : nullptr;
        a.q_int = h->q_int.as<float>();
        a.q_off = h->q_off.as<int64_t>();
        a.q_prec_mz = h->q_prec_mz.as<double>();
        a.q_charge_all = charge;
        fill_library(a, L);
        a.lib_row = h->r_best_row.as<int32_t>();
        a.pairs = h->r_pairs.as<uint32_t>();
        a.n_pairs = h->r_n_pairs.as<int32_t>();
        a.max_pairs = h->r_max_pairs;
        a.n = h->nq;
        run_features(h, a, q_prec_charge, sequence_len, out);
    });
}

}  // extern "C"

// ---------------------------------------------------------------- mode B: inverted lists sharded over GPUs

namespace solo {

// merge `parts` sorted (score desc, id asc) top-k rows per query into one: rank of an element = its
// position in its own part + number of greater keys in every other part (binary search); ids are
// unique across parts (every library row is stored on exactly one GPU), -1 ids are padding.
__global__ void __launch_bounds__(256)
merge_topk_kernel(const float *__restrict__ D, const int64_t *__restrict__ I, int parts, int nq, int k, int q_begin,
                  float *__restrict__ Do, int64_t *__restrict__ Io) {
    const int ql = blockIdx.x, q = q_begin + ql;
    auto key_of = [&](int p, int i) -> unsigned long long {
        const int64_t id = I[((int64_t)p * nq + q) * k + i];
        if (id < 0) return 0ull;
        return ((unsigned long long)ivf_f2o(D[((int64_t)p * nq + q) * k + i]) << 32) |
               (unsigned long long)(0xFFFFFFFFu - (uint32_t)id);
    };
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        Io[(int64_t)ql * k + i] = -1;
        Do[(int64_t)ql * k + i] = -INFINITY;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < parts * k; e += blockDim.x) {
        const int p = e / k, i = e % k;
        const unsigned long long key = key_of(p, i);
        if (key == 0ull) continue;
        int rank = i;
        for (int p2 = 0; p2 < parts && rank < k; ++p2) {
            if (p2 == p) continue;
            int lo = 0, hi = k;  // number of keys in part p2 greater than key (parts are sorted descending)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (key_of(p2, mid) > key) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            Io[(int64_t)ql * k + rank] = I[((int64_t)p * nq + q) * k + i];
            Do[(int64_t)ql * k + rank] = D[((int64_t)p * nq + q) * k + i];
        }
    }
}

// candidate lists from given top-k ids: precursor window AND valid (spectral_library.py:441-454)
__global__ void __launch_bounds__(256)
filter_ids_kernel(const int64_t *__restrict__ I, int k, const double *__restrict__ q_prec_mz,
                  const float *__restrict__ lib_prec_mz32, const uint8_t *__restrict__ lib_valid, int64_t n_lib, int charge,
                  double tol, int mode, int32_t *__restrict__ sel_ids, int32_t *__restrict__ sel_cnt) {
    __shared__ int s_n;
    const int q = blockIdx.x;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const double qm = q_prec_mz[q];
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const int64_t id = I[(int64_t)q * k + i];
        if (id >= 0 && id < n_lib && lib_valid[id] && window_ok(qm, lib_prec_mz32[id], charge, tol, mode))
            sel_ids[(int64_t)q * k + atomicAdd(&s_n, 1)] = (int32_t)id;
    }
    __syncthreads();
    if (threadIdx.x == 0) sel_cnt[q] = s_n;
}

}  // namespace solo

extern "C" {

int solo_ivf_set_owned_lists(solo_handle *h, int charge, const uint8_t *owned, int nlist) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        if (!owned) {
            ix.owned.clear();
        } else {
            SOLO_REQUIRE(nlist == ix.nlist, SOLO_EINVAL, "owned mask has %d entries, the index has %d lists", nlist, ix.nlist);
            ix.owned.assign(owned, owned + nlist);
        }
        ix.dirty = true;
    });
}

int solo_ivf_search_staged(solo_handle *h, int charge, int k, int nprobe, int64_t *d_I, float *d_D) {
    if (!h || !d_I) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        SOLO_REQUIRE(ix.dim == h->hash_len, SOLO_EINVAL, "index dim %d != hash_len %d", ix.dim, h->hash_len);
        const int nq = h->nq;
        if (nq == 0) return;
        DevBuf &qv = h->scratch[19];
        qv.ensure((size_t)nq * h->hash_len * sizeof(float));
        const void *mzv = h->q_mz_is_f64 < 0 ? h->q_mz.p : h->q_mz_vec.p;
        launch_vectorize(h, mzv, h->q_mz_is_f64 > 0 ? 1 : 0, h->q_int.as<float>(), h->q_off.as<int64_t>(), nq, h->q_peaks,
                         1, qv.as<float>(), nullptr, 0);
        IvfSearchArgs s;
        memset(&s, 0, sizeof s);
        s.q = qv.as<float>();
        s.nq = nq;
        s.k = k;
        s.nprobe = nprobe;
        s.I = d_I;
        s.D = d_D;
        s.win_tol_mode = -1;
        ivf_search(h, ix, s);
    });
}

int solo_ivf_probe_staged(solo_handle *h, int charge, int nprobe, int q_begin, int nq_slice, int32_t *d_probes) {
    if (!h || !d_probes) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        SOLO_REQUIRE(ix.dim == h->hash_len, SOLO_EINVAL, "index dim %d != hash_len %d", ix.dim, h->hash_len);
        SOLO_REQUIRE(q_begin >= 0 && nq_slice >= 0 && q_begin + nq_slice <= h->nq, SOLO_EINVAL, "query slice out of range");
        if (nq_slice == 0) return;
        DevBuf &qv = h->scratch[19];
        qv.ensure((size_t)h->nq * h->hash_len * sizeof(float));
        float *qs = qv.as<float>() + (size_t)q_begin * h->hash_len;
        const void *mzv = h->q_mz_is_f64 < 0 ? h->q_mz.p : h->q_mz_vec.p;
        // CSR offsets are absolute: a shifted view of the offsets is a valid batch over the same peak arrays
        launch_vectorize(h, mzv, h->q_mz_is_f64 > 0 ? 1 : 0, h->q_int.as<float>(), h->q_off.as<int64_t>() + q_begin, nq_slice,
                         h->q_peaks, 1, qs, nullptr, 0);
        IvfSearchArgs s;
        memset(&s, 0, sizeof s);
        s.q = qs;
        s.nq = nq_slice;
        s.k = 1;
        s.nprobe = nprobe;
        s.probes = d_probes;
        s.coarse_only = 1;
        s.win_tol_mode = -1;
        ivf_search(h, ix, s);
    });
}

int solo_ivf_scan_staged(solo_handle *h, int charge, int k, int nprobe, const int32_t *d_probes, int64_t *d_I, float *d_D,
                         uint64_t *d_packed) {
    if (!h || (!d_I && !d_packed) || !d_probes) return SOLO_EINVAL;
    return guarded(h, [&] {
        IvfIndex &ix = get_ivf(h, charge, true);
        SOLO_REQUIRE(ix.dim == h->hash_len, SOLO_EINVAL, "index dim %d != hash_len %d", ix.dim, h->hash_len);
        SOLO_REQUIRE(nprobe >= 1 && nprobe <= ix.nlist, SOLO_EINVAL, "the given probe rows hold %d lists, the index has %d",
                     nprobe, ix.nlist);
        const int nq = h->nq;
        if (nq == 0) return;
        DevBuf &qv = h->scratch[19];
        qv.ensure((size_t)nq * h->hash_len * sizeof(float));
        const void *mzv = h->q_mz_is_f64 < 0 ? h->q_mz.p : h->q_mz_vec.p;
        launch_vectorize(h, mzv, h->q_mz_is_f64 > 0 ? 1 : 0, h->q_int.as<float>(), h->q_off.as<int64_t>(), nq, h->q_peaks,
                         1, qv.as<float>(), nullptr, 0);
        IvfSearchArgs s;
        memset(&s, 0, sizeof s);
        s.q = qv.as<float>();
        s.nq = nq;
        s.k = k;
        s.nprobe = nprobe;
        s.given_probes = d_probes;
        s.I = d_I;
        s.D = d_D;
        s.packed = reinterpret_cast<unsigned long long *>(d_packed);
        s.win_tol_mode = -1;
        ivf_search(h, ix, s);
    });
}

int solo_merge_topk_device(solo_handle *h, const float *d_D_parts, const int64_t *d_I_parts, int parts, int nq, int k,
                           int q_begin, int nq_out, float *d_D, int64_t *d_I) {
    if (!h || !d_D_parts || !d_I_parts || !d_D || !d_I) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_REQUIRE(parts >= 1 && k >= 1 && q_begin >= 0 && nq_out >= 0 && q_begin + nq_out <= nq, SOLO_EINVAL,
                     "bad merge shape");
        if (nq_out == 0) return;
        StageTimer t(h, ST_TOPK, 1);
        merge_topk_kernel<<<nq_out, 256, 0, h->stream>>>(d_D_parts, d_I_parts, parts, nq, k, q_begin, d_D, d_I);
        SOLO_CUDA(cudaGetLastError());
    });
}

// K5 for the staged queries [q_begin, q_begin + nq_slice) over candidate rows sel (nq_slice, p->k) / r_n_cand
static void score_slice(solo_handle *h, LibraryStore &L, const solo_search_params *p, const int32_t *sel, int q_begin,
                        int nq_slice) {
    DevBuf &dpos = h->scratch[23];
    dpos.ensure((size_t)h->nq * sizeof(int32_t));
    ScoreArgs a;
    a.q_mz = h->q_mz.as<float>();
    a.q_int = h->q_int.as<float>();
    a.q_off = h->q_off.as<int64_t>() + q_begin;  // CSR offsets are absolute: a shifted view is a valid batch
    a.q_prec_mz = h->q_prec_mz.as<double>() + q_begin;
    a.nq = nq_slice;
    a.q_max_peaks = h->q_max_peaks;
    a.lib = &L;
    a.tol = p->fragment_mz_tolerance;
    a.allow_shift = p->allow_shift;
    a.max_pairs = p->max_pairs;
    a.best_pos = dpos.as<int32_t>() + q_begin;
    a.best_row = h->r_best_row.as<int32_t>() + q_begin;
    a.best_score = h->r_best_score.as<double>() + q_begin;
    a.n_pairs = h->r_n_pairs.as<int32_t>() + q_begin;
    a.pairs = h->r_pairs.as<uint32_t>() + (size_t)q_begin * p->max_pairs * 2;
    a.overflow = h->r_ovf.as<int32_t>();
    a.tie_by_row = 1;
    a.cand_ids = sel;
    a.cand_off = nullptr;
    a.cand_cnt = h->r_n_cand.as<int32_t>() + q_begin;
    a.cand_stride = p->k;
    launch_best_match(h, a);
}

static void check_slice_call(solo_handle *h, int charge, LibraryStore &L, const solo_search_params *p, int q_begin,
                             int nq_slice) {
    const int nq = h->nq;
    SOLO_REQUIRE(p->tol_mode == SOLO_TOL_DA || p->tol_mode == SOLO_TOL_PPM, SOLO_EINVAL, "Unknown precursor tolerance mode");
    SOLO_REQUIRE(p->max_pairs > 0 && p->k >= 1, SOLO_EINVAL, "bad parameters");
    SOLO_REQUIRE(q_begin >= 0 && nq_slice >= 0 && q_begin + nq_slice <= nq, SOLO_EINVAL, "query slice out of range");
    if (h->ivf.count(charge) && h->ivf[charge].nlist > 0)
        SOLO_REQUIRE(h->ivf[charge].ntotal == L.n, SOLO_ESTATE, "ANN index of charge %d holds %lld rows but the library has %lld",
                     charge, (long long)h->ivf[charge].ntotal, (long long)L.n);
    if (h->r_nq != nq || h->r_max_pairs != p->max_pairs) ensure_results(h, nq, p->max_pairs);
    h->r_ovf.ensure(16);
    SOLO_CUDA(cudaMemsetAsync(h->r_ovf.p, 0, 16, h->stream));
}

int solo_score_staged_ids(solo_handle *h, int charge, const solo_search_params *p, const int64_t *d_I, int q_begin,
                          int nq_slice) {
    if (!h || !p || !d_I) return SOLO_EINVAL;
    return guarded(h, [&] {
        LibraryStore &L = get_lib(h, charge);
        check_slice_call(h, charge, L, p, q_begin, nq_slice);
        if (nq_slice == 0) return;
        DevBuf &sel = h->scratch[20];
        sel.ensure((size_t)nq_slice * p->k * sizeof(int32_t));
        {
            StageTimer t(h, ST_CANDIDATES, 1);
            filter_ids_kernel<<<nq_slice, 256, 0, h->stream>>>(d_I, p->k, h->q_prec_mz.as<double>() + q_begin,
                                                               L.prec_mz32.as<float>(), L.valid.as<uint8_t>(), L.n, charge,
                                                               p->tol_value, p->tol_mode, sel.as<int32_t>(),
                                                               h->r_n_cand.as<int32_t>() + q_begin);
            SOLO_CUDA(cudaGetLastError());
        }
        score_slice(h, L, p, sel.as<int32_t>(), q_begin, nq_slice);
    });
}

int solo_merge_score_staged(solo_handle *h, int charge, const solo_search_params *p, const uint64_t *d_parts, int parts,
                            int slice_len, int q_begin, int nq_slice) {
    if (!h || !p || !d_parts) return SOLO_EINVAL;
    return guarded(h, [&] {
        LibraryStore &L = get_lib(h, charge);
        IvfIndex &ix = get_ivf(h, charge, true);
        check_slice_call(h, charge, L, p, q_begin, nq_slice);
        SOLO_REQUIRE(parts >= 1 && nq_slice <= slice_len, SOLO_EINVAL, "bad merge shape");
        if (nq_slice == 0) return;
        DevBuf &sel = h->scratch[20];
        sel.ensure((size_t)nq_slice * p->k * sizeof(int32_t));
        IvfSearchArgs win;
        memset(&win, 0, sizeof win);
        win.win_q_prec_mz = h->q_prec_mz.as<double>() + q_begin;
        win.win_lib_prec_mz32 = L.prec_mz32.as<float>();
        win.win_lib_valid = L.valid.as<uint8_t>();
        win.win_charge = charge;
        win.win_tol = p->tol_value;
        win.win_tol_mode = p->tol_mode;
        // the fp32 query vectors and their norms were left by solo_ivf_scan_staged (all staged queries)
        ivf_merge_select(h, ix, reinterpret_cast<const unsigned long long *>(d_parts), parts, slice_len, p->k,
                         h->scratch[19].as<float>() + (size_t)q_begin * h->hash_len, h->scratch[22].as<float>() + q_begin,
                         nq_slice, win, sel.as<int32_t>(), h->r_n_cand.as<int32_t>() + q_begin);
        score_slice(h, L, p, sel.as<int32_t>(), q_begin, nq_slice);
    });
}

}  // extern "C"

extern "C" {

// ---------------------------------------------------------------- instrumentation

int solo_profile_enable(solo_handle *h, int on) {
    if (!h) return SOLO_EINVAL;
    h->profile = on != 0;
    return SOLO_OK;
}

int solo_profile_reset(solo_handle *h) {
    if (!h) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        drain_profile(h);
        for (int s = 0; s < ST_COUNT; ++s) {
            h->prof[s].ms = 0;
            h->prof[s].launches = 0;
            h->prof[s].units = 0;
        }
        h->launches = 0;
    });
}

int solo_profile_num_stages(void) { return ST_COUNT; }
const char *solo_stage_name(int stage) { return stage >= 0 && stage < ST_COUNT ? kStageNames[stage] : ""; }

int solo_profile_get(solo_handle *h, int stage, double *ms_total, int64_t *launches, double *units) {
    if (!h || stage < 0 || stage >= ST_COUNT) return SOLO_EINVAL;
    return guarded(h, [&] {
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        drain_profile(h);
        if (ms_total) *ms_total = h->prof[stage].ms;
        if (launches) *launches = h->prof[stage].launches;
        if (units) *units = h->prof[stage].units;
    });
}

int64_t solo_kernel_launches(const solo_handle *h) { return h ? h->launches : 0; }

}  // extern "C"

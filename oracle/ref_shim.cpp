// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.
//
// extern "C" shim around the reference's *unmodified* C++ scorer
// (/root/reference/src/ann_solo/SpectrumMatch.cpp:8-133, SpectrumMatch.h:10-61).
// The reference sources are compiled from where they lie (see oracle/Makefile);
// nothing from them is copied into this repository. The shim replaces the
// reference's Cython marshalling (spectrum_match.pyx:28-108), whose memoryview
// temporaries dangle, with stable caller-owned buffers.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference leg may load the resulting oracle/_ref/libsolo_ref.so.
#include <cstdint>
#include <cstring>
#include <vector>

#include "SpectrumMatch.h"  // found via -I/root/reference/src/ann_solo

using ann_solo::Spectrum;
using ann_solo::SpectrumMatcher;
using ann_solo::SpectrumSpectrumMatch;

namespace {

// One query against n_cand candidates taken from a library peak store.
// Returns the number of matched peak pairs (or -1 if there are no candidates).
int best_match_one(const float *q_mz, const float *q_int, int q_n, double q_prec_mz, int q_charge,
                   const float *lib_mz, const float *lib_int, const uint8_t *lib_chg,
                   const int64_t *lib_off, const double *lib_prec_mz, const int32_t *lib_prec_z,
                   const int32_t *cand, int n_cand, double tol, int allow_shift, int32_t *best_pos,
                   double *best_score, uint32_t *pairs, int max_pairs) {
    if (n_cand <= 0) {
        *best_pos = -1;
        *best_score = 0.0;
        return -1;
    }
    std::vector<uint8_t> q_chg(q_n > 0 ? q_n : 1, 0);  // pyx:88 — query charges are all zero
    Spectrum query(q_prec_mz, (unsigned)q_charge, (unsigned)q_n, const_cast<float *>(q_mz),
                   const_cast<float *>(q_int), q_chg.data());
    std::vector<Spectrum *> cands;
    cands.reserve(n_cand);
    for (int i = 0; i < n_cand; ++i) {
        int64_t id = cand[i];
        int64_t b = lib_off[id], e = lib_off[id + 1];
        cands.push_back(new Spectrum(lib_prec_mz[id], (unsigned)lib_prec_z[id], (unsigned)(e - b),
                                     const_cast<float *>(lib_mz + b), const_cast<float *>(lib_int + b),
                                     const_cast<uint8_t *>(lib_chg + b)));
    }
    SpectrumMatcher matcher;
    SpectrumSpectrumMatch *res = matcher.dot(&query, cands, tol, allow_shift != 0);
    *best_pos = (int32_t)res->getCandidateIndex();
    *best_score = res->getScore();
    auto *pm = res->getPeakMatches();
    int n = (int)pm->size();
    for (int i = 0; i < n && i < max_pairs; ++i) {
        pairs[2 * i] = (*pm)[i].first;
        pairs[2 * i + 1] = (*pm)[i].second;
    }
    delete res;
    for (auto *c : cands) delete c;
    return n;
}

}  // namespace

extern "C" {

// Batched: queries in CSR, candidate ids (rows of the library store) in CSR.
// best_pos[q] = position inside the query's candidate list (-1 when the list is empty),
// n_pairs[q] = number of peak pairs, pairs[q*2*max_pairs ...] = (query_peak, library_peak).
// n_threads <= 1 runs serially like the reference; >1 uses OpenMP over queries (each call
// into the reference is independent).
int ref_best_match_batch(const float *q_mz, const float *q_int, const int64_t *q_off,
                         const double *q_prec_mz, const int32_t *q_charge, int nq,
                         const float *lib_mz, const float *lib_int, const uint8_t *lib_chg,
                         const int64_t *lib_off, const double *lib_prec_mz, const int32_t *lib_prec_z,
                         const int32_t *cand_ids, const int64_t *cand_off, double tol, int allow_shift,
                         int max_pairs, int n_threads, int32_t *best_pos, double *best_score,
                         int32_t *n_pairs, uint32_t *pairs) {
#pragma omp parallel for schedule(dynamic, 8) num_threads(n_threads > 0 ? n_threads : 1)
    for (int q = 0; q < nq; ++q) {
        int64_t b = q_off[q], e = q_off[q + 1];
        int64_t cb = cand_off[q], ce = cand_off[q + 1];
        n_pairs[q] = best_match_one(q_mz + b, q_int + b, (int)(e - b), q_prec_mz[q], q_charge[q], lib_mz,
                                    lib_int, lib_chg, lib_off, lib_prec_mz, lib_prec_z, cand_ids + cb,
                                    (int)(ce - cb), tol, allow_shift, best_pos + q, best_score + q,
                                    pairs + (size_t)q * 2 * max_pairs, max_pairs);
    }
    return 0;
}

}  // extern "C"

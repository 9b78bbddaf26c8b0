/* solo_b200.h — C-ABI of the B200-native ANN-SoLo hot path (libsolo_b200.so).
 *
 * Plain pointers and sizes only; every function returns 0 on success and a negative
 * SOLO_E* code on failure, with a message available from solo_last_error(). All host
 * buffers are owned by the caller (contiguous, little-endian); all device memory is owned
 * by the handle. Calls are synchronous on return unless stated, one handle per GPU, not
 * re-entrant. There is NO CPU fallback: without a CUDA device solo_create() fails.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/src/ann_solo/).
 */
#ifndef SOLO_B200_H
#define SOLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct solo_handle solo_handle;

enum {
    SOLO_OK = 0,
    SOLO_EINVAL = -1,    /* bad argument (maps to ValueError in the Python shim) */
    SOLO_ECUDA = -2,     /* CUDA runtime / driver error */
    SOLO_ENOMEM = -3,    /* device allocation failed */
    SOLO_ESTATE = -4,    /* call order (e.g. search before a library/index is loaded) */
    SOLO_ECAPACITY = -5  /* a fixed capacity was exceeded (reported, never truncated) */
};

enum { SOLO_TOL_DA = 0, SOLO_TOL_PPM = 1 };

/* ---- lifetime ------------------------------------------------------------------------ */
/* Replaces faiss.StandardGpuResources() + the per-call new/delete of C++ objects in
 * spectrum_match.pyx:83-108 (spectral_library.py:73-75). */
int solo_create(int device, solo_handle **out);
void solo_destroy(solo_handle *h);
const char *solo_last_error(const solo_handle *h); /* h may be NULL: last create() error */
const char *solo_version(void);
/* Run all work on an externally owned cudaStream_t (e.g. torch's current stream) so that
 * the caller's CUDA events bracket the kernels; NULL restores the handle's own stream. */
int solo_set_stream(solo_handle *h, void *cuda_stream);
/* Tuning/diagnostic switches (results never depend on them). "scan_engine": 0 = tcgen05 tensor-core list scan with
 * exact band re-rank (default), 1 = exact CUDA-core list scan (cross-check). "scan_ts": 0 (default) | 96 | 112 =
 * swapped-operand scan with the list chunk in tensor memory, queries per tile. "round0_scores" (4096): size of the
 * unconditional first scan round. "sort_items" (1): scan items ordered by cost. "compact_probes" (1): thresholded
 * coarse pass. "train_balance" (0): balancing rounds of the k-means. "nvtx" (0): an NVTX range per stage.
 * "scan_pairs", "scan_wide", "scan_hybrid", "round0_wide", "front_probes", "tc_nb", "tc_kbb", "tc_stages", "tc_debug":
 * experiments documented in DESIGN.md section 4. */
int solo_set_option(solo_handle *h, const char *key, int64_t value);
/* Block until all work queued by this handle is complete. */
int solo_synchronize(solo_handle *h);

/* ---- K1: feature-hashed vectorisation -------------------------------------------------
 * Replaces spectrum.py:166-214 spectrum_to_vector (+ :147-163 hash_idx, :123-143 get_dim),
 * called per spectrum at spectral_library.py:157-161 and :435-440. */
int solo_set_vectorizer(solo_handle *h, double min_mz, double max_mz, double bin_size, int hash_len);
/* CSR batch -> (n, hash_len) float32 rows. mz is float32 or float64 (mz_is_f64); the bin
 * arithmetic is carried out in that precision, as NumPy does for the array the caller holds. */
int solo_vectorize(solo_handle *h, const void *mz, int mz_is_f64, const float *intensity,
                   const int64_t *offsets, int64_t n, int norm, float *out);
/* The hashed slot of one m/z bin (for tests of the device LUT). */
int solo_hash_slot(solo_handle *h, int64_t bin_idx, int32_t *slot);

/* ---- library peak store (per precursor charge) ----------------------------------------
 * Replaces the per-candidate marshalling of spectrum_match.pyx:72-85 and the lru_cache'd
 * read_spectrum(idx, True) of spectral_library.py:449-454 / reader.py:218-246: processed library
 * spectra are flattened once into a device-resident store. Row r of the store is position r of
 * spec_info['charge'][charge]['id'] (reader.py:180-189). prec_mz32 is the float32 copy used by
 * the window mask (reader.py:188-189), prec_mz the float64 value handed to the scorer. */
int solo_load_library(solo_handle *h, int charge, const float *mz, const float *intensity,
                      const uint8_t *peak_charge, const int64_t *offsets, const double *prec_mz,
                      const float *prec_mz32, const int32_t *prec_charge, const uint8_t *valid,
                      int64_t n);

/* ---- IVF-Flat inner-product index (per precursor charge) ------------------------------
 * Replaces faiss.IndexFlatIP + faiss.IndexIVFFlat(METRIC_INNER_PRODUCT):
 *   train / add / write_index      spectral_library.py:167-181
 *   read_index / nprobe / search   spectral_library.py:443-444, :490-497
 *   reset                          spectral_library.py:191, :485 */
int solo_ivf_set_centroids(solo_handle *h, int charge, const float *centroids, int nlist, int dim);
/* spherical k-means on the device (Faiss `train`); x is (n, dim) float32 on the host. */
int solo_ivf_train(solo_handle *h, int charge, const float *x, int64_t n, int dim, int nlist, int iters,
                   uint64_t seed);
/* the same, with the loaded library store of this charge vectorised on the device as the
 * training set (what _create_ann_indexes feeds to train(), spectral_library.py:152-178). */
int solo_ivf_train_library(solo_handle *h, int charge, int nlist, int iters, uint64_t seed);
int solo_ivf_get_centroids(solo_handle *h, int charge, float *centroids /* nlist*dim */);
/* Append n dense rows; ids are insertion order (Faiss `add`). Rows holding NaN are skipped,
 * as Faiss does for rows its quantizer cannot assign. */
int solo_ivf_add(solo_handle *h, int charge, const float *x, int64_t n, int dim);
/* Vectorise the loaded library store of this charge on the device and add every row. */
int solo_ivf_add_library(solo_handle *h, int charge);
int solo_ivf_reset(solo_handle *h, int charge);
int solo_ivf_ntotal(solo_handle *h, int charge, int64_t *ntotal, int32_t *nlist, int32_t *dim);
/* list_of_row[i] = inverted list holding row i (-1 = not stored). */
int solo_ivf_get_assignment(solo_handle *h, int charge, int32_t *list_of_row);
/* index.search(x, k) with index.nprobe = nprobe: I (nq,k) int64 padded with -1, D (nq,k)
 * float32 descending (exact fp32 scores) padded with -inf; D may be NULL. */
int solo_ivf_search(solo_handle *h, int charge, const float *queries, int nq, int dim, int k, int nprobe,
                    int64_t *I, float *D);
/* Diagnostic: the candidate buffers the list scan of the LAST solo_ivf_search left behind —
 * per query `counts[q]` entries of (float32 score bits << 32 | library row); *cap = row stride
 * of `entries` in elements. With k >= everything scanned this exposes every raw scan score
 * (approximate fp16-input scores for the tensor-core engine). */
int solo_debug_scan_dump(solo_handle *h, int charge, int nq, int32_t *cap, int32_t *counts, uint64_t *entries);
/* Coarse quantizer alone (IndexFlatIP.search on the centroids): probes (nq, nprobe) int32. */
int solo_ivf_coarse(solo_handle *h, int charge, const float *queries, int nq, int dim, int nprobe,
                    int32_t *probes);

/* ---- Faiss index files (".idxann") -----------------------------------------------------
 * Replaces faiss.write_index(ann_index, filename) (spectral_library.py:181) and
 * faiss.read_index(filename) (spectral_library.py:490) for the index type the reference builds:
 * IndexIVFFlat over an IndexFlatIP quantizer, float32 codes, sequential ids. The byte layout is
 * Faiss' own ("IwFl" / "IxFI" / "ilar"), so a file written here loads in Faiss and a file the
 * reference wrote loads here with its trained centroids AND its list assignment untouched. */
typedef struct solo_idxann_info {
    int32_t d;             /* vector dimension (hash_len) */
    int32_t metric;        /* 0 = inner product, 1 = L2 */
    int32_t is_trained;
    int32_t reserved;
    int64_t ntotal;        /* ids run 0..ntotal-1 */
    int64_t nlist;
    int64_t nprobe;        /* value stored in the file (Faiss default 1; the reference sets it after loading) */
    int64_t code_size;     /* bytes per stored row = 4 * d */
    int64_t nstored;       /* sum of the list sizes */
    int64_t max_list_len;
    int64_t bytes_parsed;  /* == file size for a well-formed file */
    char fourcc[8];            /* "IwFl" */
    char quantizer_fourcc[8];  /* "IxFI" */
} solo_idxann_info;
/* Parse and validate the headers and list table of an index file on the host. Needs no handle and
 * no GPU (it moves no vectors). */
int solo_idxann_inspect(const char *path, solo_idxann_info *info, char *errbuf, int errbuf_len);
/* faiss.read_index: replace the index of `charge` by the file's. *nprobe (may be NULL) receives the
 * stored nprobe. SOLO_EINVAL for other index types, truncated files, non-sequential ids. */
int solo_ivf_read_index(solo_handle *h, int charge, const char *path, int64_t *nprobe);
/* faiss.write_index: `nprobe` is stored in the file's nprobe field (the reference writes the Faiss
 * default, 1). Rows skipped by add() (NaN) keep their id but appear in no list. */
int solo_ivf_write_index(solo_handle *h, int charge, const char *path, int64_t nprobe);
/* add() with the inverted list of every row chosen by the caller (list_of_row[i] in [0, nlist), or
 * -1 to reserve id i without storing the row) instead of the arg-max-centroid rule: what a Faiss
 * file carries, and what mode B (lists sharded over GPUs) uses to replay one global assignment. */
int solo_ivf_add_assigned(solo_handle *h, int charge, const float *x, int64_t n, int dim,
                          const int32_t *list_of_row);
/* index.reconstruct_n(row0, n): dense float32 copies of stored rows (skipped rows come back zero). */
int solo_ivf_reconstruct(solo_handle *h, int charge, int64_t row0, int64_t n, float *out);

/* ---- K5: (shifted) dot product, greedy peak assignment, best candidate ----------------
 * Replaces spectrum_match.pyx:28 get_best_match -> SpectrumMatch.cpp:8 SpectrumMatcher::dot,
 * batched over queries. Candidates are rows of the loaded library store of `charge`, given as
 * CSR (cand_off[nq+1], cand_ids). Outputs per query: best_pos = position inside its candidate
 * list (first maximum wins, SpectrumMatch.cpp:118; -1 when the list is empty, where the
 * reference caller skips the query, spectral_library.py:359), score, number of matched peak
 * pairs and the pairs (query_peak, library_peak) in greedy order, max_pairs slots per query. */
int solo_best_match_batch(solo_handle *h, int charge, const float *q_mz, const float *q_intensity,
                          const int64_t *q_off, const double *q_prec_mz, int nq, const int32_t *cand_ids,
                          const int64_t *cand_off, double fragment_mz_tolerance, int allow_shift,
                          int max_pairs, int32_t *best_pos, double *best_score, int32_t *n_pairs,
                          uint32_t *pairs);

/* ---- fused open search for one batch --------------------------------------------------
 * Replaces SpectralLibrary._search_batch (spectral_library.py:328-370) with
 * _get_library_candidates (:372-455) for mode == 'open', config.mode == 'ann':
 * vectorise -> IVF top-k -> precursor window AND valid (applied AFTER the top-k, :441-446)
 * -> best match. use_ann = 0 gives the brute-force candidate set (every library row in the
 * window; config.mode == 'bf', level-1 'std' search, or a charge without an index). */
typedef struct {
    int32_t use_ann;         /* 1: ANN top-k AND window; 0: window only */
    int32_t k;               /* config.num_candidates */
    int32_t nprobe;          /* config.num_probe */
    int32_t tol_mode;        /* SOLO_TOL_DA / SOLO_TOL_PPM (precursor window) */
    double tol_value;        /* precursor tolerance */
    double fragment_mz_tolerance;
    int32_t allow_shift;     /* config.allow_peak_shifts */
    int32_t mz_is_f64;       /* dtype of the q_mz_vec array used for binning */
    int32_t max_pairs;       /* slots per query in `pairs` */
    int32_t reserved;
} solo_search_params;

/* The staged batch and its results live in a "slot"; slot 0 is active after solo_create().
 * Selecting another slot parks the current one on the device, so several batches (e.g. one per
 * precursor charge) can stay resident in HBM and be searched back to back. */
int solo_select_slot(solo_handle *h, int slot);
/* Stage the queries of one batch on the device (host -> device copies, asynchronous). q_mz is
 * the float32 m/z used by the scorer (pyx:86 astype(float32)); q_mz_vec (may equal q_mz) is the
 * array used for binning in the precision the caller holds. */
int solo_stage_queries(solo_handle *h, const float *q_mz, const void *q_mz_vec, const float *q_intensity,
                       const int64_t *q_off, const double *q_prec_mz, int nq, int mz_is_f64);
/* Run the staged batch; results stay on the device. Asynchronous (stream-ordered). */
int solo_search_staged(solo_handle *h, int charge, const solo_search_params *p);
/* Copy the results of the last solo_search_staged() to the host (synchronous).
 * best_row: library row of the winner (-1 = no candidate); n_cand: candidates scored. */
int solo_fetch_results(solo_handle *h, int32_t *best_row, double *best_score, int32_t *n_pairs,
                       uint32_t *pairs, int32_t *n_cand);
/* solo_fetch_results for the result rows [q_begin, q_begin + n) only (mode B: a GPU finishes one slice of the batch). */
int solo_fetch_results_range(solo_handle *h, int q_begin, int n, int32_t *best_row, double *best_score, int32_t *n_pairs,
                             uint32_t *pairs, int32_t *n_cand);
/* stage + search + fetch. */
int solo_search_batch(solo_handle *h, int charge, const solo_search_params *p, const float *q_mz,
                      const void *q_mz_vec, const float *q_intensity, const int64_t *q_off,
                      const double *q_prec_mz, int nq, int32_t *best_row, double *best_score,
                      int32_t *n_pairs, uint32_t *pairs, int32_t *n_cand);

/* ---- library ingestion and preprocessing -----------------------------------
 * K0 replaces spectrum.process_spectrum (spectrum.py:57-119) per spectrum by one launch over a CSR
 * batch of raw spectra (m/z ascending): set_mz_range, validity (:14-36, re-checked after every step),
 * remove_precursor_peak(tol, 'Da', 2), filter_intensity(min_intensity, max_peaks),
 * scale_intensity('root' | 'rank', max_rank = max_peaks), L2 norm. The config keys are the reference's
 * (config.py:71-117); `resolution` = decimals of spectrum.round(d, 'sum') (spectrum.py:84-89) or -1. Outputs are
 * fixed-stride rows of max_peaks entries: out_mz (same type as mz), out_intensity, out_index (position
 * of the kept peak in its raw spectrum — carries annotations/peak charges along), out_count (0 for an
 * invalid spectrum), out_valid (is_valid). */
enum { SOLO_SCALING_NONE = 0, SOLO_SCALING_ROOT = 1, SOLO_SCALING_RANK = 2 };
typedef struct solo_process_params {
    double min_mz, max_mz;                 /* config.min_mz / max_mz */
    double min_mz_range;                   /* config.min_mz_range */
    double remove_precursor_tolerance;     /* config.remove_precursor_tolerance (Da) */
    double min_intensity;                  /* config.min_intensity */
    int32_t min_peaks;                     /* config.min_peaks */
    int32_t max_peaks;                     /* config.max_peaks_used(_library), <= 128 */
    int32_t remove_precursor;              /* config.remove_precursor */
    int32_t scaling;                       /* config.scaling: SOLO_SCALING_* ('sqrt' == root) */
    int32_t resolution;                    /* config.resolution: decimals of round(d, 'sum') (0..12), -1 = None */
    int32_t reserved;
} solo_process_params;
int solo_process_spectra(solo_handle *h, const void *mz, int mz_is_f64, const float *intensity,
                         const int64_t *offsets, const double *prec_mz, const int32_t *prec_charge, int64_t n,
                         const solo_process_params *p, void *out_mz, float *out_intensity, int32_t *out_index,
                         int32_t *out_count, uint8_t *out_valid);
/* SpectraST binary libraries: replaces parsers.SplibParser (parsers.pyx:41-186) for bulk ingestion.
 * Host code, no handle: count, let the caller allocate, fill. Spectra come in file order; peptide
 * strings are concatenated (peptide_offsets[n+1]); peak_charge is what the reference's annotation
 * parser yields for a/b/y ions and 0 elsewhere; file_offset is the byte offset the reference stores
 * in spec_info['offset'] (reader.py:180-187). errbuf receives the message on failure. */
int solo_splib_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_peptide_bytes,
                     char *errbuf, int errbuf_len);
int solo_splib_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t n_peptide_bytes,
                    uint32_t *identifier, double *prec_mz, int32_t *prec_charge, uint8_t *is_decoy,
                    int64_t *file_offset, int64_t *peak_offsets, float *mz, float *intensity,
                    uint8_t *peak_charge, int64_t *peptide_offsets, char *peptides, char *errbuf,
                    int errbuf_len);

/* MGF query files: replaces reader.read_mgf (reader.py:868-911, pyteomics) for bulk ingestion of raw
 * query spectra. Host code, no handle; same count / allocate / fill protocol. precursor_charge 0 = no
 * CHARGE line (None in the reference); rt_seconds NaN = no RTINSECONDS; identifiers = TITLE, else
 * SCAN(S), else the 1-based index; peaks come back m/z-ascending, m/z float64, intensity float32. */
int solo_mgf_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_identifier_bytes,
                   int64_t *n_seq_bytes, char *errbuf, int errbuf_len);
int solo_mgf_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t n_identifier_bytes,
                  int64_t n_seq_bytes, double *prec_mz, int32_t *prec_charge, double *rt_seconds,
                  uint8_t *is_decoy, int64_t *peak_offsets, double *mz, float *intensity,
                  int64_t *identifier_offsets, char *identifiers, int64_t *seq_offsets, char *seqs,
                  char *errbuf, int errbuf_len);

/* mzML query files: replaces reader.read_mzml / _parse_spectrum_mzml (reader.py:659-741, pyteomics + lxml).
 * Only MS level 2 spectra are returned; scan_nr = the integer after "scan=" (else "index=") in the spectrum
 * id (the reference's identifier, as str), index = position among ALL spectra of the file; spectra the
 * reference would skip with a warning (no scan/index number, no selected ion) are counted in n_skipped.
 * 32/64-bit float arrays, uncompressed or zlib. Same count / allocate / fill protocol, host code. */
int solo_mzml_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_skipped, char *errbuf,
                    int errbuf_len);
int solo_mzml_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t *scan_nr, int32_t *index,
                   double *prec_mz, int32_t *prec_charge, double *rt, int64_t *peak_offsets, double *mz,
                   float *intensity, char *errbuf, int errbuf_len);

/* mzXML query files: replaces reader.read_mzxml / _parse_spectrum_mzxml (reader.py:743-811). Same
 * outputs and protocol as the mzML pair: scan_nr = int(scan num), index = position among all scans
 * (nested ones in document order), rt in minutes from the xsd:duration, network-order 32/64-bit peaks,
 * uncompressed or zlib. */
int solo_mzxml_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_skipped, char *errbuf,
                     int errbuf_len);
int solo_mzxml_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t *scan_nr, int32_t *index,
                    double *prec_mz, int32_t *prec_charge, double *rt, int64_t *peak_offsets, double *mz,
                    float *intensity, char *errbuf, int errbuf_len);

/* ---- K6: SSM feature table for rescoring --------------------------------
 * Replaces utils._compute_ssm_features (utils.py:276-457): for every spectrum-spectrum match the 44
 * numeric columns that function derives from spectrum_similarity.SpectrumSimilarityCalculator(ssm) and
 * SpectrumSimilarityCalculator(ssm, top=5) (spectrum_similarity.py:13-730), one row per SSM, float64,
 * columns in the order of solo_ssm_feature_name(0..SOLO_N_SSM_FEATURES-1) (= the keys of the reference's
 * `features` dict without index / sequence / is_target). The library side of each SSM is a row of the
 * loaded peak store of `charge`; hypergeometric_score uses the vectoriser's (min_mz, max_mz, bin_size)
 * like the reference (utils.py:398-404). SSMs with lib_row < 0 or without peak matches get NaN rows
 * (the reference skips them, utils.py:332-333). q_prec_charge NULL: every query has charge `charge`;
 * sequence_len NULL: column 0 is 0. pairs/n_pairs/max_pairs have the layout of solo_fetch_results. */
#define SOLO_N_SSM_FEATURES 44
const char *solo_ssm_feature_name(int column);
int solo_ssm_features(solo_handle *h, int charge, const void *q_mz, int q_mz_is_f64, const float *q_intensity,
                      const int64_t *q_off, const double *q_prec_mz, const int32_t *q_prec_charge, int n_ssm,
                      const int32_t *lib_row, const uint32_t *pairs, const int32_t *n_pairs, int max_pairs,
                      const int32_t *sequence_len, double *out /* n_ssm x SOLO_N_SSM_FEATURES */);
/* The same for the batch and the results resident in the active slot after solo_search_staged /
 * solo_search_batch: nothing but the two optional per-query int arrays goes up, only the table comes
 * back. out is (staged queries) x SOLO_N_SSM_FEATURES. */
int solo_ssm_features_staged(solo_handle *h, int charge, const int32_t *q_prec_charge,
                             const int32_t *sequence_len, double *out);

/* ---- streaming (reference: the batch loop of _search_cascade, spectral_library.py:301-306) -----------------
 * The same three steps as solo_search_batch, split so that the copies of neighbouring batches hide under the
 * kernels: staging and fetching run on a copy stream of the handle's own, ordered against the compute stream by
 * events per query slot (solo_select_slot). Host buffers must be page-locked for the copies to overlap and must
 * stay untouched until solo_wait_results(slot) returns. Typical loop with two slots:
 *   select(s[(i+1)%2]); stage_async(batch i+1);  select(s[i%2]); search_staged(batch i); fetch_async(out i);
 *   wait_results(s[(i-1)%2]);  -> results of batch i-1 are on the host */
/* Grows the active slot's device buffers for batches of up to nq queries / n_peaks peaks (growing later would
 * free and allocate inside the stream, a device-wide synchronisation). */
int solo_reserve_slot(solo_handle *h, int nq, int64_t n_peaks, int max_pairs, int mz_is_f64);
int solo_stage_queries_async(solo_handle *h, const float *q_mz, const void *q_mz_vec, const float *q_intensity,
                             const int64_t *q_off, const double *q_prec_mz, int nq, int mz_is_f64);
int solo_fetch_results_async(solo_handle *h, int32_t *best_row, double *best_score, int32_t *n_pairs, uint32_t *pairs,
                             int32_t *n_cand);
int solo_wait_results(solo_handle *h, int slot);

/* ---- mode B: inverted lists sharded over GPUs ---------------------
 * The reference has no multi-GPU path; these entry points are what a one-process-per-GPU driver
 * needs around its collective (NCCL all-gather of the per-GPU top-k rows). All pointers named d_*
 * are DEVICE pointers (e.g. torch tensors); the calls are asynchronous on the handle's stream. */
/* owned[l] != 0: list l is stored on this GPU; the other lists stay empty here (every GPU adds
 * the whole library, row ids stay global). NULL restores "all lists". */
int solo_ivf_set_owned_lists(solo_handle *h, int charge, const uint8_t *owned, int nlist);
/* Vectorise the staged query batch and search this GPU's lists: d_I (nq,k) int64 / d_D (nq,k)
 * float32, sorted (score desc, id asc), padded with -1 / -inf. */
int solo_ivf_search_staged(solo_handle *h, int charge, int k, int nprobe, int64_t *d_I, float *d_D);
/* The same search split at the probe set, so that coarse scoring is sharded by QUERIES and the list scan by LISTS:
 * solo_ivf_probe_staged vectorises the staged queries [q_begin, q_begin + nq_slice) and writes their nprobe selected
 * lists (closest first, like the single-GPU path) to d_probes (nq_slice, nprobe) int32; after the ranks exchanged
 * their rows (all-gather), solo_ivf_scan_staged scans this GPU's lists for ALL staged queries with the given
 * d_probes (nq, nprobe) and leaves the local top-k like solo_ivf_search_staged. */
int solo_ivf_probe_staged(solo_handle *h, int charge, int nprobe, int q_begin, int nq_slice, int32_t *d_probes);
int solo_ivf_scan_staged(solo_handle *h, int charge, int k, int nprobe, const int32_t *d_probes, int64_t *d_I,
                         float *d_D, uint64_t *d_packed);
/* d_packed (nq, k) uint64 instead of d_I / d_D: the exchange format — (float32 score bits << 32 | library row),
 * unsorted, padded with 0xFF800000FFFFFFFF; certain members of the local top-k carry their approximate score, band
 * members their exact one. solo_merge_score_staged takes `parts` such tensors laid out (parts, slice_len, k) — what an
 * all-to-all leaves on the owner of the query slice [q_begin, q_begin + nq_slice) —, selects the exact global top-k
 * (same band rule as the single-GPU path, re-scoring from the sparse rows every GPU keeps), applies the precursor
 * window AND valid, and runs the best match; results land in the rows [q_begin, ...) of solo_fetch_results. */
int solo_merge_score_staged(solo_handle *h, int charge, const solo_search_params *p, const uint64_t *d_parts, int parts,
                            int slice_len, int q_begin, int nq_slice);
/* Merge `parts` such results, laid out (parts, nq, k) as all_gather_into_tensor leaves them, for the
 * queries [q_begin, q_begin + nq_out): d_D/d_I (nq_out, k), same order and padding. */
int solo_merge_topk_device(solo_handle *h, const float *d_D_parts, const int64_t *d_I_parts, int parts, int nq, int k,
                           int q_begin, int nq_out, float *d_D, int64_t *d_I);
/* Finish the staged batch's queries [q_begin, q_begin + nq_slice) from given top-k ids d_I
 * (nq_slice, p->k): precursor window AND valid, then the best match; results land in the rows
 * [q_begin, ...) that solo_fetch_results returns. */
int solo_score_staged_ids(solo_handle *h, int charge, const solo_search_params *p, const int64_t *d_I, int q_begin,
                          int nq_slice);

/* ---- instrumentation ------------------------------------------------------------------
 * Per-stage device times (CUDA events on the launching stream) accumulated since the last
 * reset, and the number of kernels this library launched. Stage names: solo_stage_name(i). */
int solo_profile_enable(solo_handle *h, int on);
int solo_profile_reset(solo_handle *h);
int solo_profile_num_stages(void);
const char *solo_stage_name(int stage);
int solo_profile_get(solo_handle *h, int stage, double *ms_total, int64_t *launches, double *units);
int64_t solo_kernel_launches(const solo_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* SOLO_B200_H */

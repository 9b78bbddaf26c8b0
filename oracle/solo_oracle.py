"""TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.

ctypes front-end for the CPU oracle (oracle/solo_oracle.cpp, the restatement) and for
oracle/_ref/libsolo_ref.so (the reference's own SpectrumMatch.cpp behind ref_shim.cpp), plus a
NumPy restatement of ``process_spectrum`` (reference src/ann_solo/spectrum.py:57-119; the five
peak operations live in spectrum_utils, which is absent -> "parity unpinned" for that stage).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module. The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "_build", "libsolo_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libsolo_ref.so")

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> None:
    """Compile the restatement and, when /root/reference is present, oracle/_ref."""
    if force or not os.path.isfile(_PORT) or (
            os.path.getmtime(_PORT) < os.path.getmtime(os.path.join(_HERE, "solo_oracle.cpp"))):
        subprocess.check_call(["make", "-C", _HERE, "_build/libsolo_oracle.so"],
                              stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE, "_build/libk6_host_check.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src/ann_solo") and (force or not os.path.isfile(_REF)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        build()
        _port = C.CDLL(_PORT)
        _port.oracle_hash_idx.restype = C.c_uint32
        _port.oracle_hash_idx.argtypes = [C.c_int64, C.c_int]
        _port.oracle_mz_to_bin_f32.restype = C.c_int64
        _port.oracle_mz_to_bin_f32.argtypes = [C.c_float, C.c_double, C.c_double]
        _port.oracle_mz_to_bin_f64.restype = C.c_int64
        _port.oracle_mz_to_bin_f64.argtypes = [C.c_double, C.c_double, C.c_double]
        _port.oracle_ip.restype = C.c_float
        _port.oracle_candidates.restype = C.c_int64
        _port.oracle_num_threads.restype = C.c_int
    return _port


def have_ref() -> bool:
    if not os.path.isfile(_REF):
        try:
            build()
        except Exception:
            return False
    return os.path.isfile(_REF)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libsolo_ref.so is not built (reference sources absent)")
        _ref = C.CDLL(_REF)
    return _ref


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def num_threads() -> int:
    """Host threads the CPU baseline uses: every core this process may run on (torchrun pins
    OMP_NUM_THREADS=1 for its workers, which would misreport the box)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return int(port().oracle_num_threads())


# ---------------------------------------------------------------- vectorisation (A2)
def hash_idx(bin_idx: int, hash_len: int) -> int:
    return int(port().oracle_hash_idx(int(bin_idx), int(hash_len)))


def get_dim(min_mz: float, max_mz: float, bin_size: float):
    n = C.c_int64()
    s = C.c_double()
    e = C.c_double()
    port().oracle_get_dim(C.c_double(min_mz), C.c_double(max_mz), C.c_double(bin_size), C.byref(n),
                          C.byref(s), C.byref(e))
    return n.value, s.value, e.value


def mz_to_bin(mz, min_bound: float, bin_size: float) -> int:
    if isinstance(mz, np.float32):
        return int(port().oracle_mz_to_bin_f32(C.c_float(float(mz)), min_bound, bin_size))
    return int(port().oracle_mz_to_bin_f64(C.c_double(float(mz)), min_bound, bin_size))


def vectorize(mz: np.ndarray, inten: np.ndarray, off: np.ndarray, min_mz=11.0, max_mz=2010.0,
              bin_size=0.04, hash_len=800, norm=True) -> np.ndarray:
    """Batched spectrum_to_vector (spectrum.py:166-214); arithmetic precision follows mz.dtype."""
    off = np.ascontiguousarray(off, np.int64)
    inten = np.ascontiguousarray(inten, np.float32)
    n = len(off) - 1
    out = np.empty((n, hash_len), np.float32)
    if mz.dtype == np.float32:
        mz = np.ascontiguousarray(mz)
        port().oracle_vectorize_f32(_p(mz, _f32p), _p(inten, _f32p), _p(off, _i64p), C.c_int64(n),
                                    C.c_double(min_mz), C.c_double(max_mz), C.c_double(bin_size),
                                    C.c_int(hash_len), C.c_int(int(norm)), _p(out, _f32p))
    else:
        mz = np.ascontiguousarray(mz, np.float64)
        port().oracle_vectorize_f64(_p(mz, _f64p), _p(inten, _f32p), _p(off, _i64p), C.c_int64(n),
                                    C.c_double(min_mz), C.c_double(max_mz), C.c_double(bin_size),
                                    C.c_int(hash_len), C.c_int(int(norm)), _p(out, _f32p))
    return out


# ---------------------------------------------------------------- scorer (A6/A7)
def _scorer_call(fn, q, lib, cand_ids, cand_off, tol, allow_shift, max_pairs, n_threads, extra):
    nq = len(q["off"]) - 1
    best_pos = np.empty(nq, np.int32)
    best_score = np.empty(nq, np.float64)
    n_pairs = np.empty(nq, np.int32)
    pairs = np.zeros((nq, max_pairs, 2), np.uint32)
    cand_ids = np.ascontiguousarray(cand_ids, np.int32)
    cand_off = np.ascontiguousarray(cand_off, np.int64)
    args = [_p(q["mz"], _f32p), _p(q["inten"], _f32p), _p(q["off"], _i64p), _p(q["prec_mz"], _f64p)]
    args += extra(q)
    args += [C.c_int(nq), _p(lib["mz"], _f32p), _p(lib["inten"], _f32p), _p(lib["chg"], _u8p),
             _p(lib["off"], _i64p), _p(lib["prec_mz"], _f64p), _p(lib["prec_z"], _i32p),
             _p(cand_ids, _i32p), _p(cand_off, _i64p), C.c_double(tol), C.c_int(int(allow_shift))]
    return args, best_pos, best_score, n_pairs, pairs


def best_match_batch(q: dict, lib: dict, cand_ids, cand_off, tol: float, allow_shift: bool,
                     sort_mode: int = 1, max_pairs: int = 64, n_threads: int = 0):
    """Restated SpectrumMatcher::dot over a batch. q/lib are dicts of contiguous arrays:
    q: mz f32, inten f32, off i64, prec_mz f64; lib: + chg u8, prec_z i32.
    Returns (best_pos i32 [-1 = no candidates], score f64, n_pairs i32, pairs u32 (nq,max_pairs,2))."""
    args, bp, bs, npairs, pairs = _scorer_call(None, q, lib, cand_ids, cand_off, tol, allow_shift,
                                               max_pairs, n_threads, lambda q: [])
    args += [C.c_int(sort_mode), C.c_int(max_pairs), C.c_int(n_threads or num_threads()),
             _p(bp, _i32p), _p(bs, _f64p), _p(npairs, _i32p), _p(pairs, _u32p)]
    port().oracle_best_match_batch(*args)
    return bp, bs, npairs, pairs


def ref_best_match_batch(q: dict, lib: dict, cand_ids, cand_off, tol: float, allow_shift: bool,
                         max_pairs: int = 64, n_threads: int = 1):
    """The reference's own compiled SpectrumMatcher::dot (oracle/_ref)."""
    qz = np.ascontiguousarray(q.get("charge", np.zeros(len(q["off"]) - 1)), np.int32)
    args, bp, bs, npairs, pairs = _scorer_call(None, q, lib, cand_ids, cand_off, tol, allow_shift,
                                               max_pairs, n_threads, lambda q: [_p(qz, _i32p)])
    args += [C.c_int(max_pairs), C.c_int(n_threads), _p(bp, _i32p), _p(bs, _f64p), _p(npairs, _i32p),
             _p(pairs, _u32p)]
    ref().ref_best_match_batch(*args)
    return bp, bs, npairs, pairs


# ---------------------------------------------------------------- window mask (A3)
def candidates(q_prec_mz, lib_prec_mz32, lib_valid, charge: int, tol: float, tol_mode: str,
               ann_ids=None):
    """Candidate CSR per query (spectral_library.py:416-454): window AND (optional) ANN ids AND
    is_valid, ascending library position."""
    q_prec_mz = np.ascontiguousarray(q_prec_mz, np.float64)
    lib_prec_mz32 = np.ascontiguousarray(lib_prec_mz32, np.float32)
    lib_valid = np.ascontiguousarray(lib_valid, np.uint8)
    nq = len(q_prec_mz)
    if tol_mode not in ("Da", "ppm"):
        raise ValueError("Unknown precursor tolerance mode")
    k = 0
    if ann_ids is not None:
        ann_ids = np.ascontiguousarray(ann_ids, np.int64)
        k = ann_ids.shape[1]
    off = np.empty(nq + 1, np.int64)
    common = [_p(q_prec_mz, _f64p), C.c_int(nq), _p(lib_prec_mz32, _f32p), _p(lib_valid, _u8p),
              C.c_int64(len(lib_prec_mz32)), C.c_int(charge), C.c_double(tol),
              C.c_int(int(tol_mode == "ppm")), _p(ann_ids, _i64p), C.c_int(k), _p(off, _i64p)]
    total = port().oracle_candidates(*common, None)
    ids = np.empty(max(total, 1), np.int32)
    port().oracle_candidates(*common, _p(ids, _i32p))
    return ids[:total], off


# ---------------------------------------------------------------- IVF (A4/A5)
def ip(a, b) -> float:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(port().oracle_ip(_p(a, _f32p), _p(b, _f32p), C.c_int(len(a))))


def ivf_assign(x, centroids, fast=True, n_threads=0):
    """List of every row: best centroid under (score desc, id asc); -1 for NaN rows. fast=True
    uses the transposed sparse-aware loop (bit-identical, see solo_oracle.cpp)."""
    x = np.ascontiguousarray(x, np.float32)
    centroids = np.ascontiguousarray(centroids, np.float32)
    out = np.empty(len(x), np.int32)
    if fast:
        port().oracle_ivf_assign_fast(_p(x, _f32p), C.c_int64(len(x)), C.c_int(x.shape[1]), _p(centroids, _f32p),
                                      C.c_int(len(centroids)), C.c_int(n_threads or num_threads()), _p(out, _i32p))
    else:
        port().oracle_ivf_assign(_p(x, _f32p), C.c_int64(len(x)), C.c_int(x.shape[1]), _p(centroids, _f32p),
                                 C.c_int(len(centroids)), _p(out, _i32p))
    return out


def build_lists(x, assign, nlist):
    """Inverted lists in insertion order (Faiss `add`): list_off, list_ids, list_vecs."""
    keep = np.flatnonzero(assign >= 0)
    order = keep[np.argsort(assign[keep], kind="stable")]
    counts = np.bincount(assign[keep], minlength=nlist)
    off = np.zeros(nlist + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    return off, order.astype(np.int64), np.ascontiguousarray(x[order], np.float32)


def ivf_coarse(q, centroids, nprobe):
    q = np.ascontiguousarray(q, np.float32)
    centroids = np.ascontiguousarray(centroids, np.float32)
    nprobe = min(nprobe, len(centroids))
    probes = np.empty((len(q), nprobe), np.int32)
    scores = np.empty((len(q), nprobe), np.float32)
    port().oracle_ivf_coarse(_p(q, _f32p), C.c_int(len(q)), C.c_int(q.shape[1]), _p(centroids, _f32p),
                             C.c_int(len(centroids)), C.c_int(nprobe), _p(probes, _i32p), _p(scores, _f32p))
    return probes, scores


def ivf_search(q, centroids, list_off, list_ids, list_vecs, nprobe, k, simd=False, n_threads=0):
    q = np.ascontiguousarray(q, np.float32)
    centroids = np.ascontiguousarray(centroids, np.float32)
    list_off = np.ascontiguousarray(list_off, np.int64)
    list_ids = np.ascontiguousarray(list_ids, np.int64)
    list_vecs = np.ascontiguousarray(list_vecs, np.float32)
    I = np.empty((len(q), k), np.int64)
    D = np.empty((len(q), k), np.float32)
    args = [_p(q, _f32p), C.c_int(len(q)), C.c_int(q.shape[1]), _p(centroids, _f32p),
            C.c_int(len(centroids)), _p(list_off, _i64p), _p(list_ids, _i64p), _p(list_vecs, _f32p),
            C.c_int(nprobe), C.c_int(k)]
    if simd:
        port().oracle_ivf_search_simd(*args, C.c_int(n_threads or num_threads()), _p(I, _i64p), _p(D, _f32p))
    else:
        port().oracle_ivf_search(*args, _p(I, _i64p), _p(D, _f32p))
    return D, I


def kmeans(x, nlist, seed=4, iters=4, n_threads=0):
    x = np.ascontiguousarray(x, np.float32)
    rng = np.random.default_rng(seed)
    finite = np.flatnonzero(np.isfinite(x).all(axis=1))
    init = np.sort(rng.choice(finite, size=nlist, replace=False)).astype(np.int64)
    cent = np.empty((nlist, x.shape[1]), np.float32)
    port().oracle_kmeans_fast(_p(x, _f32p), C.c_int64(len(x)), C.c_int(x.shape[1]), C.c_int(nlist),
                              _p(init, _i64p), C.c_int(iters), C.c_int(n_threads or num_threads()), _p(cent, _f32p))
    return cent


# ---------------------------------------------------------------- preprocessing (A1)
def process_spectrum_np(mz, intensity, precursor_mz, precursor_charge, *, min_mz=11.0, max_mz=2010.0,
                        min_peaks=10, min_mz_range=250.0, resolution=None, remove_precursor=False,
                        remove_precursor_tolerance=0.0, min_intensity=0.01, max_peaks=50,
                        scaling="rank"):
    """NumPy restatement of process_spectrum (spectrum.py:57-119) with spectrum_utils' peak
    operations as described in SURVEY.md §8c (spectrum_utils is absent: PARITY UNPINNED).
    Returns (mz, intensity f32, is_valid, kept_index) — kept_index maps to the input peaks."""
    mz = np.asarray(mz)
    intensity = np.asarray(intensity, np.float32)
    idx = np.arange(len(mz))

    def valid(m):
        return len(m) >= min_peaks and len(m) > 0 and (m[-1] - m[0]) >= m.dtype.type(min_mz_range)

    # scalars enter comparisons in the precision of the m/z array (what NumPy >= 2 does for Python
    # floats; spelled out so that the result does not depend on the NumPy version or scalar types)
    T = mz.dtype.type
    keep = (mz >= T(min_mz)) & (mz <= T(max_mz))  # set_mz_range
    mz, intensity, idx = mz[keep], intensity[keep], idx[keep]
    if not valid(mz):
        return mz, intensity, False, idx
    if resolution is not None:  # round(decimals, 'sum'): NumPy's round in the array's precision, equal values merged
        r = np.round(mz, resolution)
        uniq, inv = np.unique(r, return_inverse=True)
        summed = np.zeros(len(uniq), np.float32)
        np.add.at(summed, inv, intensity)          # float32, in peak order
        rep = np.full(len(uniq), -1)
        top = np.full(len(uniq), -np.inf)
        for i in range(len(r)):                     # the group's most intense peak (first on ties) carries the annotation
            if intensity[i] > top[inv[i]]:
                top[inv[i]], rep[inv[i]] = intensity[i], idx[i]
        mz, intensity, idx = uniq.astype(mz.dtype), summed, rep
        if not valid(mz):
            return mz, intensity, False, idx
    if remove_precursor:  # remove_precursor_peak(tol, 'Da', isotope=2)
        neutral = (float(precursor_mz) - 1.0072766) * int(precursor_charge)
        rm = np.zeros(len(mz), bool)
        for c in range(int(precursor_charge), 0, -1):
            for iso in range(3):
                rm |= np.abs(mz - T((neutral + iso) / c + 1.0072766)) <= T(remove_precursor_tolerance)
        mz, intensity, idx = mz[~rm], intensity[~rm], idx[~rm]
        if not valid(mz):
            return mz, intensity, False, idx
    # filter_intensity(min_intensity, max_num_peaks)
    order = np.argsort(intensity, kind="stable")
    thr = np.float32(min_intensity) * (intensity[order[-1]] if len(order) else np.float32(0.0))
    start = 0
    while start < len(order) and intensity[order[start]] <= thr:
        start += 1
    sel = order[max(start, len(order) - max_peaks):]
    mask = np.zeros(len(mz), bool)
    mask[sel] = True
    mz, intensity, idx = mz[mask], intensity[mask], idx[mask]
    if not valid(mz):
        return mz, intensity, False, idx
    if scaling in ("sqrt", "root"):
        intensity = np.sqrt(intensity).astype(np.float32)
    elif scaling == "rank":
        intensity = (max_peaks - np.argsort(np.argsort(intensity, kind="stable")[::-1],
                                            kind="stable")).astype(np.float32)
    nrm = np.float32(np.sqrt(np.sum(intensity.astype(np.float64) ** 2)))
    intensity = (intensity / nrm).astype(np.float32)
    return mz, intensity, True, idx

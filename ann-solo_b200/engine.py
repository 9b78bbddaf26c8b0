"""NumPy-facing wrapper of one libsolo_b200 handle (one per GPU).

All compute happens in the hand-written sm_100a kernels behind the C-ABI; this module only
checks dtypes/contiguity and hands out pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import SearchParams, SoloError, TOL_DA, TOL_PPM


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class SoloEngine:
    """Device-side state of the hot path: vectoriser LUT, per-charge library peak stores and
    per-charge IVF indexes (replaces the faiss resources and Cython matcher objects of
    reference spectral_library.py:73-87 and spectrum_match.pyx)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.solo_create(int(device), C.byref(h))
        if rc != 0:
            raise SoloError(rc, self._lib.solo_last_error(None).decode())
        self._h = h
        self.device = device
        self.hash_len = 800
        self._lib_max_peaks = {}

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.solo_last_error(self._h).decode()
            if rc == _lib.SOLO_EINVAL:
                raise ValueError(msg)
            raise SoloError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.solo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: Optional[int]):
        """Run this engine's kernels and copies on the given CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``);
        None: the engine's own non-blocking stream. Handle 0 is CUDA's legacy default stream — torch's default current
        stream — and is passed as cudaStreamLegacy so that it is not mistaken for "none"."""
        if cuda_stream is None:
            handle = None
        else:
            handle = C.c_void_p(1 if int(cuda_stream) == 0 else int(cuda_stream))   # 0x1 = cudaStreamLegacy
        self._check(self._lib.solo_set_stream(self._h, handle))

    def set_option(self, key: str, value: int):
        self._check(self._lib.solo_set_option(self._h, key.encode(), int(value)))

    def synchronize(self):
        self._check(self._lib.solo_synchronize(self._h))

    # ------------------------------------------------------------------ K1
    def set_vectorizer(self, min_mz: float, max_mz: float, bin_size: float, hash_len: int):
        self._check(self._lib.solo_set_vectorizer(self._h, min_mz, max_mz, bin_size, int(hash_len)))
        self.hash_len = int(hash_len)

    def hash_slot(self, bin_idx: int) -> int:
        s = C.c_int32()
        self._check(self._lib.solo_hash_slot(self._h, int(bin_idx), C.byref(s)))
        return s.value

    def vectorize(self, mz: np.ndarray, intensity: np.ndarray, offsets: np.ndarray, norm: bool = True) -> np.ndarray:
        """Batched spectrum_to_vector. The bin arithmetic follows mz.dtype (float32 or float64)."""
        is64 = mz.dtype == np.float64
        mz = _c(mz, np.float64 if is64 else np.float32)
        intensity = _c(intensity, np.float32)
        offsets = _c(offsets, np.int64)
        n = len(offsets) - 1
        out = np.empty((n, self.hash_len), np.float32)
        self._check(self._lib.solo_vectorize(self._h, _ptr(mz), int(is64), _ptr(intensity), _ptr(offsets), n,
                                             int(norm), _ptr(out)))
        return out

    # ------------------------------------------------------------------ library store
    def load_library(self, charge: int, store: dict):
        """store: peak-store dict (mz f32, inten f32, chg u8, off i64, prec_mz f64, prec_z i32,
        optional valid u8, optional prec_mz32 f32)."""
        mz = _c(store["mz"], np.float32)
        inten = _c(store["inten"], np.float32)
        chg = _c(store["chg"], np.uint8) if store.get("chg") is not None else None
        off = _c(store["off"], np.int64)
        pm = _c(store["prec_mz"], np.float64)
        pm32 = _c(store["prec_mz32"], np.float32) if store.get("prec_mz32") is not None else None
        pz = _c(store["prec_z"], np.int32)
        valid = _c(store["valid"], np.uint8) if store.get("valid") is not None else None
        n = len(off) - 1
        self._check(self._lib.solo_load_library(self._h, int(charge), _ptr(mz), _ptr(inten), _ptr(chg), _ptr(off),
                                                _ptr(pm), _ptr(pm32), _ptr(pz), _ptr(valid), n))
        self._lib_max_peaks[int(charge)] = int(np.diff(off).max()) if n else 0

    # ------------------------------------------------------------------ IVF
    def ivf_set_centroids(self, charge: int, centroids: np.ndarray):
        c = _c(centroids, np.float32)
        self._check(self._lib.solo_ivf_set_centroids(self._h, int(charge), _ptr(c), c.shape[0], c.shape[1]))

    def ivf_train(self, charge: int, x: np.ndarray, nlist: int, iters: int = 10, seed: int = 1234):
        x = _c(x, np.float32)
        self._check(self._lib.solo_ivf_train(self._h, int(charge), _ptr(x), x.shape[0], x.shape[1], int(nlist),
                                             int(iters), int(seed)))

    def ivf_train_library(self, charge: int, nlist: int, iters: int = 10, seed: int = 1234):
        """k-means on the loaded library store of `charge`, vectorised on the device."""
        self._check(self._lib.solo_ivf_train_library(self._h, int(charge), int(nlist), int(iters), int(seed)))

    def ivf_info(self, charge: int):
        n, nl, d = C.c_int64(), C.c_int32(), C.c_int32()
        self._check(self._lib.solo_ivf_ntotal(self._h, int(charge), C.byref(n), C.byref(nl), C.byref(d)))
        return n.value, nl.value, d.value

    def ivf_get_centroids(self, charge: int) -> np.ndarray:
        _, nl, d = self.ivf_info(charge)
        out = np.empty((nl, d), np.float32)
        self._check(self._lib.solo_ivf_get_centroids(self._h, int(charge), _ptr(out)))
        return out

    def ivf_add(self, charge: int, x: np.ndarray):
        x = _c(x, np.float32)
        self._check(self._lib.solo_ivf_add(self._h, int(charge), _ptr(x), x.shape[0], x.shape[1]))

    def ivf_add_library(self, charge: int):
        self._check(self._lib.solo_ivf_add_library(self._h, int(charge)))

    def ivf_reset(self, charge: int):
        self._check(self._lib.solo_ivf_reset(self._h, int(charge)))

    def ivf_assignment(self, charge: int) -> np.ndarray:
        n, _, _ = self.ivf_info(charge)
        out = np.empty(n, np.int32)
        self._check(self._lib.solo_ivf_get_assignment(self._h, int(charge), _ptr(out)))
        return out

    def ivf_add_assigned(self, charge: int, x: np.ndarray, list_of_row: np.ndarray):
        """add() with the inverted list of every row given by the caller (-1 = reserve the id only)."""
        x = _c(x, np.float32)
        lst = _c(list_of_row, np.int32)
        if lst.shape != (x.shape[0],):
            raise ValueError("list_of_row must have one entry per row")
        self._check(self._lib.solo_ivf_add_assigned(self._h, int(charge), _ptr(x), x.shape[0], x.shape[1], _ptr(lst)))

    def ivf_reconstruct(self, charge: int, row0: int = 0, n: Optional[int] = None) -> np.ndarray:
        ntotal, _, d = self.ivf_info(charge)
        n = ntotal - row0 if n is None else n
        out = np.empty((n, d), np.float32)
        self._check(self._lib.solo_ivf_reconstruct(self._h, int(charge), int(row0), int(n), _ptr(out)))
        return out

    def ivf_write_index(self, charge: int, path: str, nprobe: int = 1):
        """faiss.write_index (reference spectral_library.py:181)."""
        self._check(self._lib.solo_ivf_write_index(self._h, int(charge), str(path).encode(), int(nprobe)))

    def ivf_read_index(self, charge: int, path: str) -> int:
        """faiss.read_index (reference spectral_library.py:490); returns the nprobe stored in the file."""
        nprobe = C.c_int64()
        self._check(self._lib.solo_ivf_read_index(self._h, int(charge), str(path).encode(), C.byref(nprobe)))
        return nprobe.value

    def ivf_search(self, charge: int, queries: np.ndarray, k: int, nprobe: int, want_d: bool = True):
        q = _c(queries, np.float32)
        nq, d = q.shape
        I = np.empty((nq, k), np.int64)
        D = np.empty((nq, k), np.float32) if want_d else None
        self._check(self._lib.solo_ivf_search(self._h, int(charge), _ptr(q), nq, d, int(k), int(nprobe), _ptr(I),
                                              _ptr(D)))
        return D, I

    def debug_scan_dump(self, charge: int, nq: int):
        """Raw scan output of the last ivf_search: list of (scores f32, rows i64) per query."""
        cap = C.c_int32()
        counts = np.empty(nq, np.int32)
        self._check(self._lib.solo_debug_scan_dump(self._h, int(charge), nq, C.byref(cap), _ptr(counts), None))
        entries = np.zeros((nq, cap.value), np.uint64)
        self._check(self._lib.solo_debug_scan_dump(self._h, int(charge), nq, C.byref(cap), _ptr(counts), _ptr(entries)))
        out = []
        for q in range(nq):
            e = entries[q, :min(counts[q], cap.value)]
            out.append(((e >> np.uint64(32)).astype(np.uint32).view(np.float32), (e & np.uint64(0xFFFFFFFF)).astype(np.int64)))
        return counts, out

    def ivf_coarse(self, charge: int, queries: np.ndarray, nprobe: int) -> np.ndarray:
        q = _c(queries, np.float32)
        nq, d = q.shape
        _, nl, _ = self.ivf_info(charge)
        out = np.empty((nq, min(nprobe, nl)), np.int32)
        self._check(self._lib.solo_ivf_coarse(self._h, int(charge), _ptr(q), nq, d, int(nprobe), _ptr(out)))
        return out

    # ------------------------------------------------------------------ K5
    def best_match_batch(self, charge: int, q: dict, cand_ids: np.ndarray, cand_off: np.ndarray, tol: float,
                         allow_shift: bool, max_pairs: Optional[int] = None):
        qmz = _c(q["mz"], np.float32)
        qin = _c(q["inten"], np.float32)
        qoff = _c(q["off"], np.int64)
        qpm = _c(q["prec_mz"], np.float64)
        cand_ids = _c(cand_ids, np.int32)
        cand_off = _c(cand_off, np.int64)
        nq = len(qoff) - 1
        if max_pairs is None:
            max_pairs = max(1, int(np.diff(qoff).max()) if nq else 1)
        bp = np.empty(nq, np.int32)
        bs = np.empty(nq, np.float64)
        npairs = np.empty(nq, np.int32)
        pairs = np.zeros((nq, max_pairs, 2), np.uint32)
        self._check(self._lib.solo_best_match_batch(self._h, int(charge), _ptr(qmz), _ptr(qin), _ptr(qoff), _ptr(qpm),
                                                    nq, _ptr(cand_ids), _ptr(cand_off), float(tol), int(allow_shift),
                                                    int(max_pairs), _ptr(bp), _ptr(bs), _ptr(npairs), _ptr(pairs)))
        return bp, bs, npairs, pairs

    # ------------------------------------------------------------------ fused search
    @staticmethod
    def make_params(use_ann: bool, k: int, nprobe: int, tol_value: float, tol_mode: str, fragment_mz_tolerance: float,
                    allow_shift: bool, max_pairs: int = 64, mz_is_f64: bool = False) -> SearchParams:
        if tol_mode not in ("Da", "ppm"):
            raise ValueError("Unknown precursor tolerance mode")  # reference spectral_library.py:429
        return SearchParams(int(use_ann), int(k), int(nprobe), TOL_DA if tol_mode == "Da" else TOL_PPM,
                            float(tol_value), float(fragment_mz_tolerance), int(allow_shift), int(mz_is_f64),
                            int(max_pairs), 0)

    def select_slot(self, slot: int):
        """Make `slot` the active query slot (staged batch + results); others stay parked in HBM."""
        self._check(self._lib.solo_select_slot(self._h, int(slot)))
        self._slot_state = getattr(self, "_slot_state", {})
        self._slot_state[getattr(self, "_slot", 0)] = (getattr(self, "_staged", None), getattr(self, "_staged_nq", 0),
                                                       getattr(self, "_staged_max_pairs", 0))
        self._slot = int(slot)
        self._staged, self._staged_nq, self._staged_max_pairs = self._slot_state.get(self._slot, (None, 0, 0))

    def stage_queries(self, q: dict, mz_vec: Optional[np.ndarray] = None):
        """Queue the host->device copies of one query batch. The arrays must stay alive until the
        next synchronising call (they are kept referenced here)."""
        qmz = _c(q["mz"], np.float32)
        qin = _c(q["inten"], np.float32)
        qoff = _c(q["off"], np.int64)
        qpm = _c(q["prec_mz"], np.float64)
        is64 = mz_vec is not None and mz_vec.dtype == np.float64
        if mz_vec is not None:
            mz_vec = _c(mz_vec, np.float64 if is64 else np.float32)
        self._staged = (qmz, qin, qoff, qpm, mz_vec)
        self._staged_nq = len(qoff) - 1
        self._check(self._lib.solo_stage_queries(self._h, _ptr(qmz), _ptr(mz_vec), _ptr(qin), _ptr(qoff), _ptr(qpm),
                                                 self._staged_nq, int(is64)))

    def search_staged(self, charge: int, params: SearchParams):
        self._check(self._lib.solo_search_staged(self._h, int(charge), C.byref(params)))
        self._staged_max_pairs = params.max_pairs

    def fetch_results(self, out: Optional[dict] = None) -> dict:
        nq, mp = self._staged_nq, self._staged_max_pairs
        if out is None:
            out = dict(best_row=np.empty(nq, np.int32), score=np.empty(nq, np.float64),
                       n_pairs=np.empty(nq, np.int32), pairs=np.empty((nq, mp, 2), np.uint32),
                       n_cand=np.empty(nq, np.int32))
        self._check(self._lib.solo_fetch_results(self._h, _ptr(out["best_row"]), _ptr(out["score"]),
                                                 _ptr(out["n_pairs"]), _ptr(out["pairs"]), _ptr(out["n_cand"])))
        return out

    def fetch_results_range(self, q_begin: int, n: int) -> dict:
        """Result rows [q_begin, q_begin + n) of the staged batch."""
        mp = self._staged_max_pairs
        out = dict(best_row=np.empty(n, np.int32), score=np.empty(n, np.float64), n_pairs=np.empty(n, np.int32),
                   pairs=np.empty((n, mp, 2), np.uint32), n_cand=np.empty(n, np.int32))
        self._check(self._lib.solo_fetch_results_range(self._h, int(q_begin), int(n), _ptr(out["best_row"]), _ptr(out["score"]),
                                                       _ptr(out["n_pairs"]), _ptr(out["pairs"]), _ptr(out["n_cand"])))
        return out

    def search_batch(self, charge: int, params: SearchParams, q: dict, mz_vec: Optional[np.ndarray] = None,
                     out: Optional[dict] = None) -> dict:
        """One batch through the whole hot path with host buffers (vectorise -> IVF top-k ->
        window mask -> best match)."""
        self.stage_queries(q, mz_vec)
        self.search_staged(charge, params)
        return self.fetch_results(out)

    # ------------------------------------------------------------------ streaming
    def stage_queries_async(self, q: dict, mz_vec: Optional[np.ndarray] = None):
        """stage_queries on the engine's copy stream (returns at once). The arrays should be page-locked and must not
        change until the batch has been searched; they are kept referenced here."""
        qmz = _c(q["mz"], np.float32)
        qin = _c(q["inten"], np.float32)
        qoff = _c(q["off"], np.int64)
        qpm = _c(q["prec_mz"], np.float64)
        is64 = mz_vec is not None and mz_vec.dtype == np.float64
        if mz_vec is not None:
            mz_vec = _c(mz_vec, np.float64 if is64 else np.float32)
        self._staged = (qmz, qin, qoff, qpm, mz_vec)
        self._staged_nq = len(qoff) - 1
        self._check(self._lib.solo_stage_queries_async(self._h, _ptr(qmz), _ptr(mz_vec), _ptr(qin), _ptr(qoff), _ptr(qpm),
                                                       self._staged_nq, int(is64)))

    def fetch_results_async(self, out: Optional[dict] = None) -> dict:
        """Queue the device->host copies of the active slot's results behind its kernels, on the copy stream; the
        arrays are valid after ``wait_results(slot)``."""
        nq, mp = self._staged_nq, self._staged_max_pairs
        if out is None:
            out = dict(best_row=np.empty(nq, np.int32), score=np.empty(nq, np.float64),
                       n_pairs=np.empty(nq, np.int32), pairs=np.empty((nq, mp, 2), np.uint32),
                       n_cand=np.empty(nq, np.int32))
        self._check(self._lib.solo_fetch_results_async(self._h, _ptr(out["best_row"]), _ptr(out["score"]),
                                                       _ptr(out["n_pairs"]), _ptr(out["pairs"]), _ptr(out["n_cand"])))
        return out

    def wait_results(self, slot: int):
        self._check(self._lib.solo_wait_results(self._h, int(slot)))

    def search_stream(self, params: SearchParams, batches, slots=(9001, 9002)):
        """The reference's batch loop (spectral_library.py:301-306) as a software pipeline: ``batches`` yields
        ``(charge, q, out)`` — a host peak store, optionally the float64 m/z under ``q['mz_vec']``, and an optional
        dict of result arrays to fill (page-locked buffers make the copies asynchronous). The host->device copy of
        batch i+1 and the device->host copy of batch i-1 run on the copy stream under the kernels of batch i.
        Yields ``(charge, results)`` in input order; the result arrays of a batch are complete when it is yielded."""
        import os
        import time
        trace = [] if os.environ.get("SOLO_STREAM_TRACE") == "1" else None
        it = iter(batches)
        prev_slot = self._slot if hasattr(self, "_slot") else 0
        nxt = next(it, None)
        # both slots are kept as large as the largest batch seen so far (a buffer that has to grow in mid-stream
        # costs a device-wide synchronisation)
        cap = getattr(self, "_stream_cap", [0, 0, 0, 0])

        def reserve(q):
            want = [len(q["off"]) - 1, int(q["off"][-1]), params.max_pairs, int(q.get("mz_vec") is not None)]
            if any(w > c for w, c in zip(want, cap)):
                cap[:] = [max(w, c) for w, c in zip(want, cap)]
                keep = self._slot if hasattr(self, "_slot") else 0
                for s_ in slots:
                    self.select_slot(s_)
                    self._check(self._lib.solo_reserve_slot(self._h, cap[0], cap[1], cap[2], cap[3]))
                self.select_slot(keep)
            self._stream_cap = cap

        if nxt is not None:
            reserve(nxt[1])
            self.select_slot(slots[0])
            self.stage_queries_async(nxt[1], nxt[1].get("mz_vec"))
        i = 0
        pending = None   # (slot, charge, out) of the batch whose fetch is in flight
        while nxt is not None:
            cur, nxt = nxt, next(it, None)
            t0 = time.perf_counter()
            if nxt is not None:
                reserve(nxt[1])
                self.select_slot(slots[(i + 1) % 2])
                self.stage_queries_async(nxt[1], nxt[1].get("mz_vec"))
            t1 = time.perf_counter()
            self.select_slot(slots[i % 2])
            self.search_staged(cur[0], params)
            t2 = time.perf_counter()
            out = self.fetch_results_async(cur[2] if len(cur) > 2 else None)
            t3 = time.perf_counter()
            if pending is not None:
                self.wait_results(pending[0])
                if trace is not None:
                    trace.append((cur[0], t1 - t0, t2 - t1, t3 - t2, time.perf_counter() - t3))
                yield pending[1], pending[2]
            pending = (slots[i % 2], cur[0], out)
            i += 1
        if trace:
            import sys
            print("[search_stream] host ms per batch (charge: stage_async / search_staged / fetch_async / wait_results): " +
                  "  ".join(f"{z}: {a * 1e3:.2f}/{b * 1e3:.2f}/{c * 1e3:.2f}/{d * 1e3:.2f}" for z, a, b, c, d in trace[:12]),
                  file=sys.stderr)
        if pending is not None:
            self.wait_results(pending[0])
            yield pending[1], pending[2]
        self.select_slot(prev_slot)

    # ------------------------------------------------------------------ K0: process_spectrum, batched
    def process_spectra(self, store: dict, min_mz=11.0, max_mz=2010.0, min_peaks=10, min_mz_range=250.0,
                        remove_precursor=False, remove_precursor_tolerance=0.0, min_intensity=0.01, max_peaks=50,
                        scaling="rank", resolution=None, mz_vec: Optional[np.ndarray] = None) -> dict:
        """Reference spectrum.process_spectrum (spectrum.py:57-119) for every spectrum of a raw peak
        store in one launch. Returns a processed peak store (``valid`` = is_valid; invalid spectra
        keep no peaks) plus ``src`` = index of every kept peak inside its raw spectrum; peak charges
        (``chg``) and float64 m/z follow the kept peaks."""
        if resolution is not None and not 0 <= int(resolution) <= 12:
            raise ValueError("resolution must be a number of decimals in [0, 12]")
        if scaling not in _lib.SCALING:
            raise ValueError("Unknown intensity scaling")
        is64 = mz_vec is not None and mz_vec.dtype == np.float64
        mz = _c(mz_vec if is64 else store["mz"], np.float64 if is64 else np.float32)
        inten = _c(store["inten"], np.float32)
        off = _c(store["off"], np.int64)
        n = len(off) - 1
        pm = _c(store["prec_mz"], np.float64) if store.get("prec_mz") is not None else None
        pz = _c(store["prec_z"], np.int32) if store.get("prec_z") is not None else None
        prm = _lib.ProcessParams(float(min_mz), float(max_mz), float(min_mz_range), float(remove_precursor_tolerance),
                                 float(min_intensity), int(min_peaks), int(max_peaks), int(bool(remove_precursor)),
                                 _lib.SCALING[scaling], -1 if resolution is None else int(resolution), 0)
        o_mz = np.empty((n, max_peaks), mz.dtype)
        o_in = np.empty((n, max_peaks), np.float32)
        o_ix = np.empty((n, max_peaks), np.int32)
        o_ct = np.empty(n, np.int32)
        o_va = np.empty(n, np.uint8)
        self._check(self._lib.solo_process_spectra(self._h, _ptr(mz), int(is64), _ptr(inten), _ptr(off), _ptr(pm),
                                                   _ptr(pz), n, C.byref(prm), _ptr(o_mz), _ptr(o_in), _ptr(o_ix),
                                                   _ptr(o_ct), _ptr(o_va)))
        keep = np.arange(max_peaks)[None, :] < o_ct[:, None]
        new_off = np.zeros(n + 1, np.int64)
        np.cumsum(o_ct, out=new_off[1:])
        src = o_ix[keep]
        gsrc = np.repeat(off[:-1], o_ct) + src
        # (with `resolution` a merged peak stands at its group's most intense member: src / chg follow that one)
        out = dict(mz=o_mz[keep].astype(np.float32), inten=o_in[keep], off=new_off, valid=o_va, src=src,
                   prec_mz=store.get("prec_mz"), prec_z=store.get("prec_z"))
        if is64:
            out["mz64"] = o_mz[keep]
        if store.get("chg") is not None:
            out["chg"] = np.asarray(store["chg"], np.uint8)[gsrc]
        for key in ("is_decoy", "id", "peptide", "file_offset"):
            if key in store:
                out[key] = store[key]
        return out

    # ------------------------------------------------------------------ K6: SSM features
    def feature_names(self):
        return [self._lib.solo_ssm_feature_name(i).decode() for i in range(_lib.N_SSM_FEATURES)]

    def ssm_features(self, charge: int, q: dict, lib_row: np.ndarray, pairs: np.ndarray, n_pairs: np.ndarray,
                     q_charge: Optional[np.ndarray] = None, sequence_len: Optional[np.ndarray] = None,
                     mz_vec: Optional[np.ndarray] = None) -> np.ndarray:
        """Feature table (n_ssm, 44) float64 for SSMs between the queries of peak store `q` and rows
        `lib_row` of the loaded library store of `charge` (reference utils._compute_ssm_features).
        `pairs` is (n_ssm, max_pairs, 2) as returned by the search; `mz_vec` optionally carries the
        query m/z in float64 (the precision the caller's spectra hold)."""
        is64 = mz_vec is not None and mz_vec.dtype == np.float64
        qmz = _c(mz_vec if is64 else q["mz"], np.float64 if is64 else np.float32)
        qin = _c(q["inten"], np.float32)
        qoff = _c(q["off"], np.int64)
        qpm = _c(q["prec_mz"], np.float64)
        n = len(qoff) - 1
        rows = _c(lib_row, np.int32)
        pairs = _c(pairs, np.uint32)
        npairs = _c(n_pairs, np.int32)
        if pairs.ndim != 3 or pairs.shape[0] != n or pairs.shape[2] != 2 or rows.shape != (n,) or npairs.shape != (n,):
            raise ValueError("pairs must be (n_ssm, max_pairs, 2); lib_row and n_pairs (n_ssm,)")
        zc = None if q_charge is None else _c(q_charge, np.int32)
        sl = None if sequence_len is None else _c(sequence_len, np.int32)
        out = np.empty((n, _lib.N_SSM_FEATURES), np.float64)
        self._check(self._lib.solo_ssm_features(self._h, int(charge), _ptr(qmz), int(is64), _ptr(qin), _ptr(qoff),
                                                _ptr(qpm), _ptr(zc), n, _ptr(rows), _ptr(pairs), _ptr(npairs),
                                                max(1, pairs.shape[1]), _ptr(sl), _ptr(out)))
        return out

    def ssm_features_staged(self, charge: int, q_charge: Optional[np.ndarray] = None,
                            sequence_len: Optional[np.ndarray] = None) -> np.ndarray:
        """The same for the batch and results resident in the active slot (after search_staged)."""
        n = self._staged_nq
        zc = None if q_charge is None else _c(q_charge, np.int32)
        sl = None if sequence_len is None else _c(sequence_len, np.int32)
        out = np.empty((n, _lib.N_SSM_FEATURES), np.float64)
        self._check(self._lib.solo_ssm_features_staged(self._h, int(charge), _ptr(zc), _ptr(sl), _ptr(out)))
        return out

    # ------------------------------------------------------------------ mode B (lists sharded over GPUs)
    def ivf_set_owned_lists(self, charge: int, owned: Optional[np.ndarray]):
        """owned[l] != 0: list l is stored on this GPU (None: all lists)."""
        if owned is None:
            self._check(self._lib.solo_ivf_set_owned_lists(self._h, int(charge), None, 0))
        else:
            o = _c(owned, np.uint8)
            self._check(self._lib.solo_ivf_set_owned_lists(self._h, int(charge), _ptr(o), len(o)))

    def ivf_search_staged(self, charge: int, k: int, nprobe: int, d_I: int, d_D: int):
        """Search this GPU's lists for the staged batch; d_I / d_D are device pointers of (nq, k)
        int64 / float32 buffers (e.g. torch tensors' data_ptr())."""
        self._check(self._lib.solo_ivf_search_staged(self._h, int(charge), int(k), int(nprobe), C.c_void_p(d_I),
                                                     C.c_void_p(d_D)))

    def ivf_probe_staged(self, charge: int, nprobe: int, q_begin: int, nq_slice: int, d_probes: int):
        """Coarse scoring + probe selection for the staged queries [q_begin, q_begin + nq_slice): d_probes is a
        device pointer of an (nq_slice, nprobe) int32 buffer."""
        self._check(self._lib.solo_ivf_probe_staged(self._h, int(charge), int(nprobe), int(q_begin), int(nq_slice),
                                                    C.c_void_p(d_probes)))

    def ivf_scan_staged(self, charge: int, k: int, nprobe: int, d_probes: int, d_I: int = 0, d_D: int = 0,
                        d_packed: int = 0):
        """Scan this GPU's lists for every staged query with given probe rows d_probes (nq, nprobe) int32. Output:
        sorted d_I / d_D, or d_packed (nq, k) uint64 — the unsorted exchange format of mode B."""
        self._check(self._lib.solo_ivf_scan_staged(self._h, int(charge), int(k), int(nprobe), C.c_void_p(d_probes),
                                                   C.c_void_p(d_I or None), C.c_void_p(d_D or None),
                                                   C.c_void_p(d_packed or None)))

    def merge_score_staged(self, charge: int, params: SearchParams, d_parts: int, parts: int, slice_len: int,
                           q_begin: int, nq_slice: int):
        """Mode B, owner of a query slice: exact global top-k out of `parts` packed local top-k tensors
        (parts, slice_len, k), precursor window, best match — results in rows [q_begin, q_begin + nq_slice)."""
        self._check(self._lib.solo_merge_score_staged(self._h, int(charge), C.byref(params), C.c_void_p(d_parts),
                                                      int(parts), int(slice_len), int(q_begin), int(nq_slice)))
        self._staged_max_pairs = params.max_pairs

    def merge_topk_device(self, d_D_parts: int, d_I_parts: int, parts: int, nq: int, k: int, q_begin: int, nq_out: int,
                          d_D: int, d_I: int):
        self._check(self._lib.solo_merge_topk_device(self._h, C.c_void_p(d_D_parts), C.c_void_p(d_I_parts), int(parts),
                                                     int(nq), int(k), int(q_begin), int(nq_out), C.c_void_p(d_D),
                                                     C.c_void_p(d_I)))

    def score_staged_ids(self, charge: int, params: SearchParams, d_I: int, q_begin: int, nq_slice: int):
        self._check(self._lib.solo_score_staged_ids(self._h, int(charge), C.byref(params), C.c_void_p(d_I), int(q_begin),
                                                    int(nq_slice)))
        self._staged_max_pairs = params.max_pairs

    # ------------------------------------------------------------------ instrumentation
    def profile_enable(self, on: bool = True):
        self._check(self._lib.solo_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._check(self._lib.solo_profile_reset(self._h))

    def profile(self) -> dict:
        out = {}
        for s in range(self._lib.solo_profile_num_stages()):
            ms, n, u = C.c_double(), C.c_int64(), C.c_double()
            self._check(self._lib.solo_profile_get(self._h, s, C.byref(ms), C.byref(n), C.byref(u)))
            out[self._lib.solo_stage_name(s).decode()] = dict(ms=ms.value, launches=n.value, units=u.value)
        return out

    def kernel_launches(self) -> int:
        return int(self._lib.solo_kernel_launches(self._h))

"""ctypes binding of libsolo_b200.so (the C-ABI declared in include/solo_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100a device is present,
every entry point raises. The oracle under ``oracle/`` is test infrastructure and is never
imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOLO_LIB_PATH") or os.path.join(_HERE, "libsolo_b200.so")  # (override: build-variant experiments)

SOLO_OK, SOLO_EINVAL, SOLO_ECUDA, SOLO_ENOMEM, SOLO_ESTATE, SOLO_ECAPACITY = 0, -1, -2, -3, -4, -5
TOL_DA, TOL_PPM = 0, 1
N_SSM_FEATURES = 44


class SoloError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsolo_b200 error {code}: {msg}")
        self.code = code


class SearchParams(C.Structure):
    _fields_ = [("use_ann", C.c_int32), ("k", C.c_int32), ("nprobe", C.c_int32), ("tol_mode", C.c_int32),
                ("tol_value", C.c_double), ("fragment_mz_tolerance", C.c_double), ("allow_shift", C.c_int32),
                ("mz_is_f64", C.c_int32), ("max_pairs", C.c_int32), ("reserved", C.c_int32)]


class ProcessParams(C.Structure):
    _fields_ = [("min_mz", C.c_double), ("max_mz", C.c_double), ("min_mz_range", C.c_double),
                ("remove_precursor_tolerance", C.c_double), ("min_intensity", C.c_double), ("min_peaks", C.c_int32),
                ("max_peaks", C.c_int32), ("remove_precursor", C.c_int32), ("scaling", C.c_int32),
                ("resolution", C.c_int32), ("reserved", C.c_int32)]


SCALING = {None: 0, "root": 1, "sqrt": 1, "rank": 2}


class IdxannInfo(C.Structure):
    _fields_ = [("d", C.c_int32), ("metric", C.c_int32), ("is_trained", C.c_int32), ("reserved", C.c_int32),
                ("ntotal", C.c_int64), ("nlist", C.c_int64), ("nprobe", C.c_int64), ("code_size", C.c_int64),
                ("nstored", C.c_int64), ("max_list_len", C.c_int64), ("bytes_parsed", C.c_int64),
                ("fourcc", C.c_char * 8), ("quantizer_fourcc", C.c_char * 8)]


# every symbol include/solo_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
SYMBOLS = {
    "solo_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "solo_destroy": (None, [_vp]),
    "solo_last_error": (C.c_char_p, [_vp]),
    "solo_version": (C.c_char_p, []),
    "solo_set_stream": (C.c_int, [_vp, _vp]),
    "solo_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "solo_synchronize": (C.c_int, [_vp]),
    "solo_set_vectorizer": (C.c_int, [_vp, _f64, _f64, _f64, C.c_int]),
    "solo_vectorize": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _i64, C.c_int, _vp]),
    "solo_hash_slot": (C.c_int, [_vp, _i64, C.POINTER(_i32)]),
    "solo_load_library": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "solo_ivf_set_centroids": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int]),
    "solo_ivf_train": (C.c_int, [_vp, C.c_int, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "solo_ivf_train_library": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "solo_ivf_get_centroids": (C.c_int, [_vp, C.c_int, _vp]),
    "solo_ivf_add": (C.c_int, [_vp, C.c_int, _vp, _i64, C.c_int]),
    "solo_ivf_add_library": (C.c_int, [_vp, C.c_int]),
    "solo_ivf_reset": (C.c_int, [_vp, C.c_int]),
    "solo_ivf_ntotal": (C.c_int, [_vp, C.c_int, C.POINTER(_i64), C.POINTER(_i32), C.POINTER(_i32)]),
    "solo_ivf_get_assignment": (C.c_int, [_vp, C.c_int, _vp]),
    "solo_ivf_search": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "solo_debug_scan_dump": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_i32), _vp, _vp]),
    "solo_ivf_coarse": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "solo_idxann_inspect": (C.c_int, [C.c_char_p, _vp, C.c_char_p, C.c_int]),
    "solo_ivf_read_index": (C.c_int, [_vp, C.c_int, C.c_char_p, C.POINTER(_i64)]),
    "solo_ivf_write_index": (C.c_int, [_vp, C.c_int, C.c_char_p, _i64]),
    "solo_ivf_add_assigned": (C.c_int, [_vp, C.c_int, _vp, _i64, C.c_int, _vp]),
    "solo_ivf_reconstruct": (C.c_int, [_vp, C.c_int, _i64, _i64, _vp]),
    "solo_best_match_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _f64, C.c_int,
                                        C.c_int, _vp, _vp, _vp, _vp]),
    "solo_select_slot": (C.c_int, [_vp, C.c_int]),
    "solo_stage_queries": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "solo_search_staged": (C.c_int, [_vp, C.c_int, C.POINTER(SearchParams)]),
    "solo_fetch_results": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "solo_fetch_results_range": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "solo_search_batch": (C.c_int, [_vp, C.c_int, C.POINTER(SearchParams), _vp, _vp, _vp, _vp, _vp, C.c_int,
                                    _vp, _vp, _vp, _vp, _vp]),
    "solo_reserve_slot": (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int]),
    "solo_stage_queries_async": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "solo_fetch_results_async": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "solo_wait_results": (C.c_int, [_vp, C.c_int]),
    "solo_process_spectra": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "solo_splib_count": (C.c_int, [C.c_char_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.c_char_p, C.c_int]),
    "solo_splib_read": (C.c_int, [C.c_char_p, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  C.c_char_p, C.c_int]),
    "solo_mgf_count": (C.c_int, [C.c_char_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                                 C.c_char_p, C.c_int]),
    "solo_mgf_read": (C.c_int, [C.c_char_p, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                C.c_char_p, C.c_int]),
    "solo_mzml_count": (C.c_int, [C.c_char_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.c_char_p, C.c_int]),
    "solo_mzml_read": (C.c_int, [C.c_char_p, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_char_p, C.c_int]),
    "solo_mzxml_count": (C.c_int, [C.c_char_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.c_char_p, C.c_int]),
    "solo_mzxml_read": (C.c_int, [C.c_char_p, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_char_p, C.c_int]),
    "solo_ssm_feature_name": (C.c_char_p, [C.c_int]),
    "solo_ssm_features": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int,
                                    _vp, _vp]),
    "solo_ssm_features_staged": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "solo_ivf_set_owned_lists": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "solo_ivf_search_staged": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "solo_ivf_probe_staged": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "solo_ivf_scan_staged": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "solo_merge_score_staged": (C.c_int, [_vp, C.c_int, C.POINTER(SearchParams), _vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "solo_merge_topk_device": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "solo_score_staged_ids": (C.c_int, [_vp, C.c_int, C.POINTER(SearchParams), _vp, C.c_int, C.c_int]),
    "solo_profile_enable": (C.c_int, [_vp, C.c_int]),
    "solo_profile_reset": (C.c_int, [_vp]),
    "solo_profile_num_stages": (C.c_int, []),
    "solo_stage_name": (C.c_char_p, [C.c_int]),
    "solo_profile_get": (C.c_int, [_vp, C.c_int, C.POINTER(_f64), C.POINTER(_i64), C.POINTER(_f64)]),
    "solo_kernel_launches": (_i64, [_vp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libsolo_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], stdout=out)
    return LIB_PATH


def load():
    """dlopen the library and bind every declared symbol. Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SoloError(SOLO_ESTATE, f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; "
                                     f"g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

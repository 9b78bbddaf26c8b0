"""Loads the reference's UNMODIFIED src/ann_solo/spectral_library.py (and spectrum.py / config.py) with the
modules it imports replaced by the ones under test — the drop-in claim of BASELINE.json's north_star:

    import faiss                          -> the Faiss-surface module under test (ann_solo_b200.index on a GPU,
                                             an oracle-backed stand-in on the CPU)
    from ann_solo import spectrum_match   -> the get_best_match module under test
    import numexpr as ne                  -> a NumPy evaluator of the two window expressions (numexpr is absent)
    from ann_solo import reader, utils    -> in-memory reader / pass-through FDR (out of scope)
    mmh3, spectrum_utils, configargparse  -> the stubs tests/golden/make_golden.py uses

The reference source is read from where it lies (/root/reference, or $SOLO_REFERENCE_SRC); nothing is copied.
TEST INFRASTRUCTURE — never imported by the product package.
"""
import argparse
import importlib.util
import os
import sys
import types

import numpy as np

REF_SRC = os.environ.get("SOLO_REFERENCE_SRC", "/root/reference/src/ann_solo")


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "spectral_library.py"))


def numexpr_shim():
    """numexpr.evaluate for the expressions at spectral_library.py:421-427: evaluated by NumPy in the caller's
    frame (float64 promotion of the float32 library m/z, like numexpr)."""
    ne = types.ModuleType("numexpr")

    def evaluate(expr):
        frame = sys._getframe(1)
        env = dict(frame.f_globals)
        env.update(frame.f_locals)
        env["abs"] = np.abs
        return eval(expr, {"__builtins__": {}}, env)

    ne.evaluate = evaluate
    return ne


class _Saved:
    def __init__(self, names):
        self.saved = {n: sys.modules.get(n) for n in names}

    def restore(self):
        for n, m in self.saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def load_reference(faiss_mod, spectrum_match_mod, reader_mod, utils_mod, spectrum_mod=None, config_values=None):
    """Returns (reference spectral_library module, its config object, restore()). With spectrum_mod None the
    reference's own unmodified spectrum.py is used (mmh3 stubbed with sklearn's MurmurHash3_x86_32)."""
    names = ["faiss", "numexpr", "mmh3", "spectrum_utils", "spectrum_utils.spectrum", "configargparse", "ann_solo",
             "ann_solo.config", "ann_solo.spectrum", "ann_solo.reader", "ann_solo.spectrum_match", "ann_solo.utils",
             "ann_solo.spectral_library"]
    saved = _Saved(names)
    from sklearn.utils import murmurhash3_32

    mmh3 = types.ModuleType("mmh3")
    mmh3.hash = lambda key, seed=0, signed=True: int(murmurhash3_32(key, seed=seed, positive=not signed))
    sys.modules["mmh3"] = mmh3
    su = types.ModuleType("spectrum_utils")
    sus = types.ModuleType("spectrum_utils.spectrum")
    from ann_solo_b200.spectrum import MsmsSpectrum
    sus.MsmsSpectrum = MsmsSpectrum
    su.spectrum = sus
    sys.modules["spectrum_utils"] = su
    sys.modules["spectrum_utils.spectrum"] = sus
    cap = types.ModuleType("configargparse")

    class ArgParser(argparse.ArgumentParser):
        def __init__(self, *a, **k):
            for key in ("default_config_files", "args_for_setting_config_path", "formatter_class"):
                k.pop(key, None)
            super().__init__()

        def parse_args(self, args=None, namespace=None):   # configargparse also takes one string
            if isinstance(args, str):
                args = args.split()
            return super().parse_args(args, namespace)

    cap.ArgParser = ArgParser
    cap.ArgumentDefaultsHelpFormatter = argparse.ArgumentDefaultsHelpFormatter
    sys.modules["configargparse"] = cap
    sys.modules["faiss"] = faiss_mod
    sys.modules["numexpr"] = numexpr_shim()
    pkg = types.ModuleType("ann_solo")
    pkg.__path__ = [REF_SRC]
    sys.modules["ann_solo"] = pkg

    def load(name):
        spec = importlib.util.spec_from_file_location(f"ann_solo.{name}", os.path.join(REF_SRC, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"ann_solo.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        return mod

    cfg_mod = load("config")                      # the reference's own configargparse singleton, unmodified
    cfg = cfg_mod.config
    argv = ["lib.splib", "query.mgf", "out.mztab"]
    for key, val in (config_values or {}).items():
        if isinstance(val, bool):
            if val:
                argv.append(f"--{key}")
        elif val is not None:
            argv += [f"--{key}", str(val)]
    cfg.parse(" ".join(argv))
    if spectrum_mod is None:
        spectrum_mod = load("spectrum")           # the reference's own spectrum.py, unmodified
    else:
        sys.modules["ann_solo.spectrum"] = spectrum_mod
        pkg.spectrum = spectrum_mod
    for name, mod in (("reader", reader_mod), ("spectrum_match", spectrum_match_mod), ("utils", utils_mod)):
        sys.modules[f"ann_solo.{name}"] = mod
        setattr(pkg, name, mod)
    sl = load("spectral_library")                 # the module under the drop-in claim, unmodified
    return sl, cfg, saved.restore


def reader_module(library, queries):
    """ann_solo.reader stand-in: SpectralLibraryReader(filename, config_hash) -> the given in-memory library;
    read_query_file(filename) -> the given query spectra (file readers are out of scope)."""
    m = types.ModuleType("ann_solo.reader")

    class SpectralLibraryReader:
        def __new__(cls, filename, config_hash=None):
            return library

    m.SpectralLibraryReader = SpectralLibraryReader
    m.read_query_file = lambda filename: iter(queries)
    return m


def utils_module():
    """ann_solo.utils stand-in: score_ssms keeps every SSM with q = 0 (FDR / mokapot are out of scope)."""
    m = types.ModuleType("ann_solo.utils")

    def score_ssms(ssms, fdr, model, is_open):
        for s in ssms:
            s.q = 0.0
        return ssms

    m.score_ssms = score_ssms
    return m


class ReaderSurface:
    """The SpectralLibraryReader members spectral_library.py touches, over an InMemoryLibrary."""

    def __init__(self, inmem):
        self._lib = inmem
        self.spec_info = inmem.spec_info
        self.is_recreated = False

    def read_spectrum(self, spec_id, process_peaks=False):
        return self._lib.read_spectrum(spec_id, process_peaks)

    def read_all_spectra(self):
        for ident in self._lib.identifiers.tolist():
            yield self._lib.read_spectrum(ident, False)

    def charge_store(self, charge):
        return self._lib.charge_store(charge)

    def close(self):
        pass

"""Host-side mirror of the library half of the reference's src/ann_solo/reader.py for ``.splib``
files: ``SpectralLibraryReader`` keeps ``spec_info`` (reader.py:180-189), ``read_spectrum(spec_id,
process_peaks)`` (:218-246), ``read_all_spectra`` / ``read_library_file`` (:249-287) and the context
manager, but is built from ONE native pass over the file (csrc/splib_io.cu) and hands whole precursor
charges to the device as processed peak stores (``charge_store``: K0, csrc/k0_process.cu) instead of
one Python object per spectrum.

Not mirrored: the ``.spcfg`` / HDF5 cache (reader.py:147-200, :440-556; h5py and joblib stores are
out of scope, the parsed library stays in host memory), ``.sptxt`` / ``.mgf`` libraries and decoy
generation (``config.add_decoys``).
"""
from __future__ import annotations

import os
from typing import Iterator, Optional

import numpy as np

from .config import config
from .parsers import _Annotation, read_splib
from .spectrum import MsmsSpectrum, process_spectrum


def verify_extension(supported_extensions, filename: str) -> None:
    """Reference reader.py (verify_extension): FileNotFoundError for a missing or unsupported file."""
    _, ext = os.path.splitext(os.path.basename(filename))
    if ext.lower() not in supported_extensions:
        raise FileNotFoundError(f"Unrecognized file format (supported file formats: "
                                f"{', '.join(supported_extensions)})")
    if not os.path.isfile(filename):
        raise FileNotFoundError(f"File {filename} does not exist")


class SpectralLibraryReader:
    """Read spectra from a SpectraST ``.splib`` spectral library (reference reader.py:29-437)."""

    _supported_extensions = [".splib"]
    is_recreated = False

    def __init__(self, filename: str, config_hash: Optional[str] = None, engine=None) -> None:
        self._filename = filename
        self._config_hash = config_hash
        self._engine = engine
        verify_extension(self._supported_extensions, filename)
        st = read_splib(filename)
        self._store = st
        self._ids = np.array([str(i) for i in st["id"]])            # parsers.pyx:145 str(identifier)
        self._row_of = {ident: r for r, ident in enumerate(self._ids.tolist())}
        self.spec_info = {"charge": {}}
        self._rows = {}
        for z in sorted(set(st["prec_z"].tolist())):                 # file order inside every charge
            rows = np.flatnonzero(st["prec_z"] == z)
            self._rows[int(z)] = rows
            self.spec_info["charge"][int(z)] = {"id": self._ids[rows],
                                                "precursor_mz": st["prec_mz"][rows].astype(np.float32)}
        self._processed = {}

    # ------------------------------------------------------------------ reference surface
    def open(self) -> None:
        pass

    def close(self) -> None:
        pass

    def __enter__(self) -> "SpectralLibraryReader":
        return self

    def __exit__(self, exc_type, exc_value, traceback) -> None:
        pass

    def get_version(self) -> str:
        return "null"

    def _raw_spectrum(self, r: int) -> MsmsSpectrum:
        st = self._store
        b, e = st["off"][r], st["off"][r + 1]
        ann = [None if c == 0 else _Annotation(int(c)) for c in st["chg"][b:e]]
        s = MsmsSpectrum(self._ids[r], st["prec_mz"][r], int(st["prec_z"][r]), st["mz"][b:e].copy(),
                         st["inten"][b:e].copy(), annotation=ann, peptide=st["peptide"][r],
                         is_decoy=bool(st["is_decoy"][r]))
        s.index = r
        s.is_processed = False
        return s

    def read_spectrum(self, spec_id, process_peaks: bool = False) -> MsmsSpectrum:
        """Reference :218-246. With ``process_peaks`` the spectrum comes from the K0-processed store of
        its charge (the same peaks ``process_spectrum`` keeps, bit for bit), so the objects handed to
        callers and the device-resident library agree."""
        r = self._row_of[str(spec_id)]
        if not process_peaks:
            return self._raw_spectrum(r)
        z = int(self._store["prec_z"][r])
        ps = self.charge_store(z)
        i = int(np.searchsorted(self._rows[z], r))
        b, e = ps["off"][i], ps["off"][i + 1]
        st = self._store
        ann = [None if c == 0 else _Annotation(int(c)) for c in ps["chg"][b:e]]
        s = MsmsSpectrum(self._ids[r], st["prec_mz"][r], z, ps["mz"][b:e], ps["inten"][b:e], annotation=ann,
                         peptide=st["peptide"][r], is_decoy=bool(st["is_decoy"][r]))
        s.charge = ps["chg"][b:e]
        s.index = i
        s.is_valid = bool(ps["valid"][i])
        s.is_processed = True
        return s

    def read_all_spectra(self) -> Iterator[MsmsSpectrum]:
        for r in range(len(self._ids)):
            yield self._raw_spectrum(r)

    def read_library_file(self) -> Iterator[MsmsSpectrum]:
        yield from self.read_all_spectra()

    # ------------------------------------------------------------------ batched path
    def charge_store(self, charge: int) -> dict:
        """Processed peak store of one precursor charge, rows in ``spec_info['charge'][charge]['id']``
        order: ``process_spectrum(spectrum, is_library=True)`` for every spectrum in one K0 launch."""
        if charge in self._processed:
            return self._processed[charge]
        from .spectrum import default_engine
        from .synth import take_spectra
        eng = self._engine or default_engine()
        raw = dict(self._store)
        raw["valid"] = np.ones(len(self._ids), np.uint8)
        sub = take_spectra(raw, self._rows[charge])
        out = eng.process_spectra(sub, min_mz=config.min_mz, max_mz=config.max_mz, min_peaks=config.min_peaks,
                                  min_mz_range=config.min_mz_range, remove_precursor=config.remove_precursor,
                                  remove_precursor_tolerance=config.remove_precursor_tolerance,
                                  min_intensity=config.min_intensity, max_peaks=config.max_peaks_used_library,
                                  scaling=config.scaling, resolution=config.resolution)
        out["prec_mz"] = sub["prec_mz"]
        out["prec_z"] = sub["prec_z"]
        out["prec_mz32"] = self.spec_info["charge"][charge]["precursor_mz"]
        out["is_decoy"] = sub["is_decoy"]
        self._processed[charge] = out
        return out

"""Host-side mirror of the library half of the reference's src/ann_solo/reader.py for ``.splib``
files: ``SpectralLibraryReader`` keeps ``spec_info`` (reader.py:180-189), ``read_spectrum(spec_id,
process_peaks)`` (:218-246), ``read_all_spectra`` / ``read_library_file`` (:249-287) and the context
manager, but is built from ONE native pass over the file (csrc/splib_io.cu) and hands whole precursor
charges to the device as processed peak stores (``charge_store``: K0, csrc/k0_process.cu) instead of
one Python object per spectrum.

The query side: ``read_mgf`` / ``read_query_file`` (reference :868-938) over the native MGF parser
(csrc/mgf_io.cu), plus ``read_mgf_store`` = the whole file as one raw peak store for K0.

``read_mzml`` / ``read_mzxml`` (+ ``_store`` variants) do the same for mzML and mzXML (csrc/mzml_io.cu).

Not mirrored: the ``.spcfg`` / HDF5 cache (reader.py:147-200, :440-556; h5py and joblib stores are
out of scope, the parsed library stays in host memory) and decoy generation (``config.add_decoys``).
``.mgf`` libraries (MassIVE-KB style, SEQ= lines) go through the native MGF parser, ``.sptxt`` libraries
through ``parsers.read_sptxt`` (Python, like the reference's own regex-based reader).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Iterator, Optional

import numpy as np

from . import _lib
from .config import config
from .parsers import _Annotation, _raise, read_splib, read_sptxt
from .spectrum import MsmsSpectrum, process_spectrum


def verify_extension(supported_extensions, filename: str) -> None:
    """Reference reader.py (verify_extension): FileNotFoundError for a missing or unsupported file."""
    _, ext = os.path.splitext(os.path.basename(filename))
    if ext.lower() not in supported_extensions:
        raise FileNotFoundError(f"Unrecognized file format (supported file formats: "
                                f"{', '.join(supported_extensions)})")
    if not os.path.isfile(filename):
        raise FileNotFoundError(f"File {filename} does not exist")


def _sort_peaks_by_mz(st: dict) -> None:
    """Peaks of every spectrum in ascending m/z, intensity and fragment charge attached (stable). The
    reference gets this from the MsmsSpectrum constructor (spectrum_utils sorts by m/z); K0's range check and
    K5's merge both assume it, and library files are not guaranteed to be sorted."""
    mz, off = st["mz"], st["off"]
    if len(mz) < 2:
        return
    inner = np.ones(len(mz), bool)
    inner[off[1:-1][off[1:-1] < len(mz)]] = False       # first peak of every spectrum starts a new run
    if not (np.diff(mz) < 0)[inner[1:]].any():
        return
    seg = np.repeat(np.arange(len(off) - 1), np.diff(off))
    order = np.lexsort((mz, seg))
    for key in ("mz", "inten", "chg"):
        if key in st:
            st[key] = np.ascontiguousarray(st[key][order])


class SpectralLibraryReader:
    """Read spectra from a SpectraST ``.splib`` spectral library (reference reader.py:29-437)."""

    _supported_extensions = [".splib", ".sptxt", ".mgf"]   # reference reader.py:34
    is_recreated = False

    def __init__(self, filename: str, config_hash: Optional[str] = None, engine=None) -> None:
        self._filename = filename
        self._config_hash = config_hash
        self._engine = engine
        verify_extension(self._supported_extensions, filename)
        if os.path.splitext(filename)[1].lower() == ".mgf":
            st = _mgf_library_store(filename)                        # reference :283-284 read_mgf as a library
            self._ids = np.array(st["id"])
        elif os.path.splitext(filename)[1].lower() == ".sptxt":
            st = read_sptxt(filename)                                # reference :285-286 read_sptxt
            self._ids = np.array(st["id"])
        else:
            st = read_splib(filename)
            self._ids = np.array([str(i) for i in st["id"]])        # parsers.pyx:145 str(identifier)
        _sort_peaks_by_mz(st)
        self._store = st
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


# ---------------------------------------------------------------------- MGF libraries (MassIVE-KB style)
def _leading_substitute_pattern(match) -> str:
    # reference reader.py:814-835
    if match.group(1) and match.group(2):
        return "[{}]?[{}]-{:s}".format(match.group(1), match.group(2), match.group(3))
    if match.group(1):
        return "[{}]-{}".format(match.group(1), match.group(3))
    return match.group(0)


def _mgf_seq_to_proforma(peptide: str) -> str:
    """Reference reader.py:837-866: MassIVE-KB sequence ('+42.011AC+57.021K') to ProForma."""
    import re
    formatted = re.sub(r"([A-Z])([+-]?\d+\.\d+)", r"\1[\2]", peptide)
    return re.sub(r"([+-]?[\d.]+)([+-]?[\d.]+)?([A-Za-z]+)", _leading_substitute_pattern, formatted)


def _mgf_library_store(filename: str) -> dict:
    """An MGF spectral library as the raw peak store the ``.splib`` path produces: identifier = TITLE /
    SCAN / index, peptide = ProForma of SEQ, no fragment annotations (reference :906-909). Entries
    without a CHARGE line cannot be assigned to a precursor charge and are refused."""
    st = read_mgf_store(filename)
    if (st["prec_z"] <= 0).any():
        bad = int(np.flatnonzero(st["prec_z"] <= 0)[0])
        raise ValueError(f"library spectrum {st['identifier'][bad]!r} of {filename} has no (positive) CHARGE")
    st["id"] = st.pop("identifier")
    st["peptide"] = [_mgf_seq_to_proforma(q) if q else None for q in st.pop("seq")]
    st["chg"] = np.zeros(len(st["mz"]), np.uint8)
    return st


# ---------------------------------------------------------------------- query files
def read_mgf_store(filename: str) -> dict:
    """Every spectrum of an MGF file as one RAW peak store (file order): mz float32 + mz64 float64
    (ascending inside a spectrum), inten float32, off, prec_mz, prec_z (0 = no CHARGE line), rt (NaN = not
    given), is_decoy, identifier (list of str), seq (list of str, '' = not given)."""
    lib = _lib.load()
    err = C.create_string_buffer(512)
    n, npk, nid, nseq = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    path = os.fspath(filename).encode()
    rc = lib.solo_mgf_count(path, C.byref(n), C.byref(npk), C.byref(nid), C.byref(nseq), err, len(err))
    if rc:
        _raise(rc, err)
    n, npk, nid, nseq = n.value, npk.value, nid.value, nseq.value
    out = dict(prec_mz=np.empty(n, np.float64), prec_z=np.empty(n, np.int32), rt=np.empty(n, np.float64),
               is_decoy=np.empty(n, np.uint8), off=np.empty(n + 1, np.int64), mz64=np.empty(npk, np.float64),
               inten=np.empty(npk, np.float32))
    id_off, seq_off = np.empty(n + 1, np.int64), np.empty(n + 1, np.int64)
    ids, seqs = np.empty(max(nid, 1), np.uint8), np.empty(max(nseq, 1), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.solo_mgf_read(path, n, npk, nid, nseq, p(out["prec_mz"]), p(out["prec_z"]), p(out["rt"]), p(out["is_decoy"]),
                           p(out["off"]), p(out["mz64"]), p(out["inten"]), p(id_off), p(ids), p(seq_off), p(seqs), err,
                           len(err))
    if rc:
        _raise(rc, err)
    raw_ids, raw_seqs = ids.tobytes(), seqs.tobytes()
    out["identifier"] = [raw_ids[id_off[i]:id_off[i + 1]].decode() for i in range(n)]
    out["seq"] = [raw_seqs[seq_off[i]:seq_off[i + 1]].decode() for i in range(n)]
    out["mz"] = out["mz64"].astype(np.float32)
    out["valid"] = np.ones(n, np.uint8)
    return out


def read_mgf(filename: str) -> Iterator[MsmsSpectrum]:
    """Reference reader.py:868-911: one MsmsSpectrum per entry, ``index`` 1-based, ``precursor_charge``
    None without a CHARGE line, ``retention_time`` None without RTINSECONDS. SEQ is kept as written
    (the reference rewrites MassIVE-KB modifications to ProForma, :837-866)."""
    st = read_mgf_store(filename)
    for i in range(len(st["prec_mz"])):
        b, e = st["off"][i], st["off"][i + 1]
        z = int(st["prec_z"][i])
        rt = float(st["rt"][i])
        s = MsmsSpectrum(st["identifier"][i], st["prec_mz"][i], z if z != 0 else None, st["mz64"][b:e].copy(),
                         st["inten"][b:e].copy(), retention_time=None if math.isnan(rt) else rt,
                         peptide=st["seq"][i] or None, is_decoy=bool(st["is_decoy"][i]))
        s.index = i + 1
        s.is_processed = False
        yield s


def _read_xml_store(filename: str, kind: str) -> dict:
    lib = _lib.load()
    count, read = (lib.solo_mzml_count, lib.solo_mzml_read) if kind == "mzml" else (lib.solo_mzxml_count,
                                                                                     lib.solo_mzxml_read)
    err = C.create_string_buffer(512)
    n, npk, nskip = C.c_int64(), C.c_int64(), C.c_int64()
    path = os.fspath(filename).encode()
    rc = count(path, C.byref(n), C.byref(npk), C.byref(nskip), err, len(err))
    if rc:
        _raise(rc, err)
    n, npk = n.value, npk.value
    out = dict(scan_nr=np.empty(n, np.int64), index=np.empty(n, np.int32), prec_mz=np.empty(n, np.float64),
               prec_z=np.empty(n, np.int32), rt=np.empty(n, np.float64), off=np.empty(n + 1, np.int64),
               mz64=np.empty(npk, np.float64), inten=np.empty(npk, np.float32))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = read(path, n, npk, p(out["scan_nr"]), p(out["index"]), p(out["prec_mz"]), p(out["prec_z"]),
              p(out["rt"]), p(out["off"]), p(out["mz64"]), p(out["inten"]), err, len(err))
    if rc:
        _raise(rc, err)
    out["identifier"] = [str(s) for s in out["scan_nr"]]
    out["mz"] = out["mz64"].astype(np.float32)
    out["valid"] = np.ones(n, np.uint8)
    out["n_skipped"] = nskip.value
    return out


def read_mzml_store(filename: str) -> dict:
    """Every MS2 spectrum of an mzML file as one RAW peak store: mz float32 + mz64 float64 (ascending),
    inten float32, off, prec_mz, prec_z (0 = no charge cvParam), rt (as written; NaN = absent), scan_nr,
    index (position among all spectra of the file), identifier (list of str = str(scan_nr)), n_skipped
    (MS2 spectra the reference would skip with a warning)."""
    return _read_xml_store(filename, "mzml")


def read_mzxml_store(filename: str) -> dict:
    """The same for mzXML (rt in minutes from the xsd:duration, prec_z 0 = no precursorCharge)."""
    return _read_xml_store(filename, "mzxml")


def _spectra_of(st: dict) -> Iterator[MsmsSpectrum]:
    for i in range(len(st["prec_mz"])):
        b, e = st["off"][i], st["off"][i + 1]
        z = int(st["prec_z"][i])
        rt = float(st["rt"][i])
        s = MsmsSpectrum(st["identifier"][i], st["prec_mz"][i], z if z != 0 else None, st["mz64"][b:e].copy(),
                         st["inten"][b:e].copy(), retention_time=None if math.isnan(rt) else rt)
        s.index = int(st["index"][i])
        s.is_processed = False
        yield s


def read_mzml(source: str) -> Iterator[MsmsSpectrum]:
    """Reference reader.py:659-741: MS level 2 spectra only, identifier = str(scan number), ``index`` =
    position among all spectra of the file, ``precursor_charge`` None without a charge cvParam."""
    return _spectra_of(read_mzml_store(source))


def read_mzxml(source: str) -> Iterator[MsmsSpectrum]:
    """Reference reader.py:743-811."""
    return _spectra_of(read_mzxml_store(source))


def read_query_file(filename: str) -> Iterator[MsmsSpectrum]:
    """Reference reader.py:914-938."""
    verify_extension([".mgf", ".mzml", ".mzxml"], filename)
    _, ext = os.path.splitext(os.path.basename(filename))
    if ext.lower() == ".mgf":
        return read_mgf(filename)
    if ext.lower() == ".mzml":
        return read_mzml(filename)
    return read_mzxml(filename)

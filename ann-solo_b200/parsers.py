"""Host-side mirror of the reference's src/ann_solo/parsers.pyx for bulk library ingestion: the whole
``.splib`` file becomes one peak store (CSR arrays) in two passes of native code behind the C-ABI
(csrc/splib_io.cu), instead of one MsmsSpectrum + one FragmentAnnotation object per peak
(parsers.pyx:101-157). ``SplibParser`` keeps the reference's per-spectrum interface on top of it.
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np

from . import _lib
from .spectrum import MsmsSpectrum


def _raise(rc, err):
    msg = err.value.decode()
    if rc == _lib.SOLO_EINVAL:
        if msg.startswith("cannot open"):
            raise FileNotFoundError(msg)
        raise ValueError(msg)
    raise _lib.SoloError(rc, msg)


def read_splib(filename: str) -> dict:
    """Peak store of every spectrum in file order: mz/inten float32, chg uint8 (fragment charge of a/b/y
    annotations, else 0), off int64, prec_mz float64, prec_z int32, is_decoy uint8, id uint32,
    file_offset int64 (spec_info['offset'], reference reader.py:180-187), peptide (list of str)."""
    lib = _lib.load()
    err = C.create_string_buffer(512)
    n, npk, npep = C.c_int64(), C.c_int64(), C.c_int64()
    path = str(filename).encode()
    rc = lib.solo_splib_count(path, C.byref(n), C.byref(npk), C.byref(npep), err, len(err))
    if rc:
        _raise(rc, err)
    n, npk, npep = n.value, npk.value, npep.value
    out = dict(id=np.empty(n, np.uint32), prec_mz=np.empty(n, np.float64), prec_z=np.empty(n, np.int32),
               is_decoy=np.empty(n, np.uint8), file_offset=np.empty(n, np.int64), off=np.empty(n + 1, np.int64),
               mz=np.empty(npk, np.float32), inten=np.empty(npk, np.float32), chg=np.empty(npk, np.uint8))
    pep_off = np.empty(n + 1, np.int64)
    pep = np.empty(max(npep, 1), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.solo_splib_read(path, n, npk, npep, p(out["id"]), p(out["prec_mz"]), p(out["prec_z"]), p(out["is_decoy"]),
                             p(out["file_offset"]), p(out["off"]), p(out["mz"]), p(out["inten"]), p(out["chg"]),
                             p(pep_off), p(pep), err, len(err))
    if rc:
        _raise(rc, err)
    raw = pep.tobytes()
    out["peptide"] = [raw[pep_off[i]:pep_off[i + 1]].decode() for i in range(n)]
    out["valid"] = np.ones(n, np.uint8)
    return out


class _Annotation:
    __slots__ = ("charge",)

    def __init__(self, charge):
        self.charge = charge


class SplibParser:
    """Reference parsers.pyx:41-157: ``seek_first_spectrum()`` / ``read_spectrum(offset=None)``
    returning ``(spectrum, offset)`` and raising StopIteration at the end of the file."""

    def __init__(self, filename):
        self._store = read_splib(filename.decode() if isinstance(filename, bytes) else filename)
        self._row_of_offset = {int(o): i for i, o in enumerate(self._store["file_offset"])}
        self._next = 0

    def seek_first_spectrum(self):
        self._next = 0

    def read_spectrum(self, offset: int = None):
        if offset is not None and offset >= 0:
            self._next = self._row_of_offset[int(offset)]
        st, i = self._store, self._next
        if i >= len(st["id"]):
            raise StopIteration
        self._next = i + 1
        b, e = st["off"][i], st["off"][i + 1]
        ann = [None if c == 0 else _Annotation(int(c)) for c in st["chg"][b:e]]
        spectrum = MsmsSpectrum(str(st["id"][i]), st["prec_mz"][i], int(st["prec_z"][i]), st["mz"][b:e].copy(),
                                st["inten"][b:e].copy(), annotation=ann, peptide=st["peptide"][i],
                                is_decoy=bool(st["is_decoy"][i]))
        return spectrum, int(st["file_offset"][i])


# ---------------------------------------------------------------------- SpectraST text libraries
_SPTXT_ENTRY = re.compile(rb"(?<![a-zA-Z])Name:\s?(?:(?!((?<![a-zA-Z])Name:\s?)).|\n)*", re.IGNORECASE)


def sptxt_annotation_charge(annotation: str) -> int:
    """Fragment charge as reference reader._parse_fragment_annotation (reader.py:565-597) yields it for the
    ``.sptxt`` path: ions a / b / y / p only; without '^' the charge is 1 (also for neutral losses: the
    reference takes abs(-1)); with '^' the leading integer behind it (1 when there is none). 0 = no
    annotation (None in the reference)."""
    if not annotation or annotation[0] not in "abyp":
        return 0
    index_charge = annotation[1:].split("/", 1)[0].split("^")
    if len(index_charge) == 1:
        return 1
    m = re.search(r"^\d+", index_charge[1])
    return abs(int(m.group(0))) if m else 1


def sptxt_seq_to_proforma(peptide: str, modifications) -> str:
    """Reference reader._sptxt_seq_to_proforma (reader.py:300-322): '[name]' inserted behind residue idx of
    every 'idx,aa,name' entry of the Mods= field. A residue is one upper-case letter with what SpectraST
    hangs on it: a lower-case / bracketed terminal prefix ('n[43]A') and bracketed masses behind it
    ('C[160]'); pyteomics' parser, which the reference calls, only knows plain and modX sequences."""
    residues = re.findall(r"[a-z]*(?:\[[^\]]*\])?[A-Z](?:\[[^\]]*\])*", peptide)
    if "".join(residues) != peptide:
        residues = [peptide]
    for shift, modification in enumerate(modifications or ()):
        idx, _, name = modification.split(",")
        residues.insert(int(idx) + shift + 1, "[" + name + "]")
    return "".join(residues)


def read_sptxt(filename: str) -> dict:
    """A SpectraST ``.sptxt`` library as the raw peak store ``read_splib`` produces (reference
    reader.py:324-418): entries start at 'Name:', identifier = 1-based position, peptide / charge from the
    Name line, PrecursorMZ: (else Parent=), decoy <=> 'decoy' anywhere in the metadata, Mods= for the
    ProForma peptide, tab-separated peak lines (m/z, intensity, annotation) behind 'NumPeaks:'."""
    data = open(filename, "rb").read()
    out = dict(id=[], peptide=[], prec_z=[], prec_mz=[], is_decoy=[], off=[0], mz=[], inten=[], chg=[])
    for ident, match in enumerate(_SPTXT_ENTRY.finditer(data), 1):
        raw = "\n".join(match.group(0).decode("utf-8").splitlines())
        tokens = re.split(r"Num\s?Peaks:\s?[0-9]+\n", raw.strip(), flags=re.IGNORECASE)
        if len(tokens) < 2:
            raise ValueError(f"{filename}: entry {ident} has no NumPeaks line")
        meta, peaks = tokens[0], tokens[1]
        peptide_charge = meta.split("\n", 1)[0].split("/")
        peptide = peptide_charge[0].split(" ")[-1].strip()
        charge = int(peptide_charge[1].strip())
        pm = re.search(r"PrecursorMZ:\s?[0-9]+.[0-9]+", meta, re.IGNORECASE) or \
            re.search(r"Parent=\s?[0-9]+.[0-9]+", meta, re.IGNORECASE)
        if pm is None:
            raise ValueError(f"{filename}: entry {ident} has neither PrecursorMZ nor Parent")
        mods = re.search(r"Mods=.+?(?=[\s\n])", meta, re.IGNORECASE)
        mods = str(mods.group(0)).split("/")[1:] if mods else None
        n0 = len(out["mz"])
        for line in peaks.strip().split("\n"):
            cols = line.split("\t")
            if len(cols) < 2 or not cols[0].strip():
                continue
            out["mz"].append(np.float32(float(cols[0])))
            out["inten"].append(np.float32(float(cols[1])))
            out["chg"].append(sptxt_annotation_charge(cols[2]) if len(cols) > 2 else 0)
        order = np.argsort(np.array(out["mz"][n0:], np.float32), kind="stable")   # MsmsSpectrum orders by m/z
        for key in ("mz", "inten", "chg"):
            seg = out[key][n0:]
            out[key][n0:] = [seg[i] for i in order]
        out["id"].append(str(ident))
        out["peptide"].append(sptxt_seq_to_proforma(peptide, mods))
        out["prec_z"].append(charge)
        out["prec_mz"].append(float(re.search(r"[0-9]+.[0-9]+", pm.group(0)).group(0)))
        out["is_decoy"].append(bool(re.search("decoy", meta, re.IGNORECASE)))
        out["off"].append(len(out["mz"]))
    n = len(out["id"])
    return dict(id=out["id"], peptide=out["peptide"], prec_z=np.array(out["prec_z"], np.int32),
                prec_mz=np.array(out["prec_mz"], np.float64), is_decoy=np.array(out["is_decoy"], np.uint8),
                off=np.array(out["off"], np.int64), mz=np.array(out["mz"], np.float32),
                inten=np.array(out["inten"], np.float32), chg=np.array(out["chg"], np.uint8),
                valid=np.ones(n, np.uint8))

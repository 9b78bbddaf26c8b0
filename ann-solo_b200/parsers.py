"""Host-side mirror of the reference's src/ann_solo/parsers.pyx for bulk library ingestion: the whole
``.splib`` file becomes one peak store (CSR arrays) in two passes of native code behind the C-ABI
(csrc/splib_io.cu), instead of one MsmsSpectrum + one FragmentAnnotation object per peak
(parsers.pyx:101-157). ``SplibParser`` keeps the reference's per-spectrum interface on top of it.
"""
from __future__ import annotations

import ctypes as C

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

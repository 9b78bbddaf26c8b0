"""TEST INFRASTRUCTURE ONLY (oracle): the SpectraST binary ``.splib`` layout as the reference's parser
reads it (reference parsers.pyx:41-186) — a pure-Python reader restating that parser field by field,
and a writer that produces files of the same layout for tests (no .splib file ships with the
reference; **parity unpinned** beyond the parser source itself).

Layout: int32 version, int32 sub-version, line (file name), int32 k, k preamble lines; then per
spectrum: uint32 id, line "n.PEPTIDE.c/z ...", float64 precursor m/z, line (status), uint32 num_peaks,
num_peaks x (float64 m/z, float64 intensity, line annotation, line info), line comment.
"""
from __future__ import annotations

import struct
from typing import List

import numpy as np


def write_splib(path: str, spectra: List[dict], preamble=("### synthetic library", "### for tests")) -> List[int]:
    """spectra: dicts with id, peptide, charge, prec_mz, mz, intensity, annotations (list of str),
    decoy (bool), optional mods (text behind the charge in the name line). Returns the byte offsets."""
    offsets = []
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", 5, 0))
        f.write(b"synthetic.splib\n")
        f.write(struct.pack("<i", len(preamble)))
        for line in preamble:
            f.write(line.encode() + b"\n")
        for s in spectra:
            offsets.append(f.tell())
            f.write(struct.pack("<I", s["id"]))
            f.write(f"n.{s['peptide']}.c/{s['charge']}{s.get('mods', '')}\n".encode())
            f.write(struct.pack("<d", s["prec_mz"]))
            f.write(b"Normal\n")
            f.write(struct.pack("<I", len(s["mz"])))
            for m, i, a in zip(s["mz"], s["intensity"], s["annotations"]):
                f.write(struct.pack("<dd", float(m), float(i)))
                f.write(a.encode() + b"\n")
                f.write(b"info\n")
            remark = " Remark=DECOY_x" if s.get("decoy") else " Remark=_NONE_"
            f.write(f"Spec=Consensus{remark} Nreps=1/1\n".encode())
    return offsets


def parse_annotation(raw: bytes) -> int:
    """Peak charge as reference parse_annotation (parsers.pyx:160-186) yields it; 0 = no annotation."""
    if not raw or raw[:1] not in (b"a", b"b", b"y"):
        return 0
    q = 1
    while q < len(raw) and raw[q:q + 1].isdigit():
        q += 1
    if q == 1:
        return 0
    slash = raw.find(b"/", q)
    if slash == q:
        return 1
    if raw[q:q + 1] == b"^":
        digits = raw[q + 1:slash if slash >= 0 else len(raw)]
        k = 0
        while k < len(digits) and digits[k:k + 1].isdigit():
            k += 1
        return int(digits[:k]) if k else 0
    return 0


def read_splib(path: str) -> dict:
    buf = open(path, "rb").read()
    pos = 8

    def line():
        nonlocal pos
        e = buf.find(b"\n", pos)
        e = len(buf) if e < 0 else e
        out = buf[pos:e]
        pos = min(e + 1, len(buf))
        return out

    line()
    (k,) = struct.unpack_from("<i", buf, pos)
    pos += 4
    for _ in range(k):
        line()
    out = dict(id=[], peptide=[], prec_charge=[], prec_mz=[], is_decoy=[], file_offset=[], off=[0], mz=[], inten=[],
               chg=[])
    while pos < len(buf):
        out["file_offset"].append(pos)
        (ident,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        name = line()
        p0 = name.find(b".") + 1
        p1 = name.find(b".", p0)
        out["peptide"].append(name[p0:p1].decode())
        c0 = name.find(b"/", p1) + 1
        c1 = c0
        while c1 < len(name) and name[c1:c1 + 1].isdigit():
            c1 += 1
        out["prec_charge"].append(int(name[c0:c1]))
        (pm,) = struct.unpack_from("<d", buf, pos)
        pos += 8
        line()
        (npk,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        for _ in range(npk):
            m, i = struct.unpack_from("<dd", buf, pos)
            pos += 16
            out["chg"].append(parse_annotation(line()))
            line()
            out["mz"].append(np.float32(m))
            out["inten"].append(np.float32(i))
        out["is_decoy"].append(b" Remark=DECOY_" in line())
        out["id"].append(ident)
        out["prec_mz"].append(pm)
        out["off"].append(len(out["mz"]))
    return dict(id=np.array(out["id"], np.uint32), peptide=out["peptide"],
                prec_z=np.array(out["prec_charge"], np.int32), prec_mz=np.array(out["prec_mz"], np.float64),
                is_decoy=np.array(out["is_decoy"], np.uint8), file_offset=np.array(out["file_offset"], np.int64),
                off=np.array(out["off"], np.int64), mz=np.array(out["mz"], np.float32),
                inten=np.array(out["inten"], np.float32), chg=np.array(out["chg"], np.uint8))

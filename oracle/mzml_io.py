"""TEST INFRASTRUCTURE ONLY (oracle): an mzML writer for tests and an independent reader (stdlib
xml.etree + base64 + zlib) restating what the reference takes from an mzML spectrum
(reference reader.py:659-741 via pyteomics.mzml, which is absent: **parity unpinned**).
"""
from __future__ import annotations

import base64
import math
import xml.etree.ElementTree as ET
import zlib
from typing import List

import numpy as np

NS = "{http://psi.hupo.org/ms/mzml}"


def _binary(values, bits: int, compress: bool, kind: str) -> str:
    raw = np.asarray(values, "<f8" if bits == 64 else "<f4").tobytes()
    if compress:
        raw = zlib.compress(raw)
    text = base64.b64encode(raw).decode()
    acc_type = ("MS:1000523", "64-bit float") if bits == 64 else ("MS:1000521", "32-bit float")
    acc_comp = ("MS:1000574", "zlib compression") if compress else ("MS:1000576", "no compression")
    acc_kind = ("MS:1000514", "m/z array") if kind == "mz" else ("MS:1000515", "intensity array")
    return (f'<binaryDataArray encodedLength="{len(text)}">'
            f'<cvParam cvRef="MS" accession="{acc_type[0]}" name="{acc_type[1]}" value=""/>'
            f'<cvParam cvRef="MS" accession="{acc_comp[0]}" name="{acc_comp[1]}" value=""/>'
            f'<cvParam cvRef="MS" accession="{acc_kind[0]}" name="{acc_kind[1]}" value="" unitCvRef="MS" '
            f'unitAccession="MS:1000040" unitName="m/z"/>'
            f'<binary>{text}</binary></binaryDataArray>')


def write_mzml(path: str, spectra: List[dict]) -> None:
    """spectra: dicts with id, ms_level, mz, intensity and optionally prec_mz, charge, possible_charge,
    rt, bits (32/64), zlib (bool), extra_precursor (a second precursor that must be ignored)."""
    out = ['<?xml version="1.0" encoding="utf-8"?>',
           '<mzML xmlns="http://psi.hupo.org/ms/mzml" version="1.1.0">',
           '<cvList count="1"><cv id="MS" fullName="PSI-MS" URI="x"/></cvList>',
           '<run id="run1">', f'<spectrumList count="{len(spectra)}" defaultDataProcessingRef="dp">']
    for i, s in enumerate(spectra):
        n = len(s["mz"])
        out.append(f'<spectrum index="{i}" id="{s["id"]}" defaultArrayLength="{n}">')
        out.append(f'<cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="{s["ms_level"]}"/>')
        out.append('<cvParam cvRef="MS" accession="MS:1000127" name="centroid spectrum" value=""/>')
        out.append('<scanList count="1"><cvParam cvRef="MS" accession="MS:1000795" name="no combination" value=""/><scan>')
        if "rt" in s:
            out.append(f'<cvParam cvRef="MS" accession="MS:1000016" name="scan start time" value="{s["rt"]!r}" '
                       f'unitCvRef="UO" unitAccession="UO:0000031" unitName="minute"/>')
        out.append('</scan></scanList>')
        if "prec_mz" in s:
            out.append('<precursorList count="1"><precursor><selectedIonList count="1"><selectedIon>')
            out.append(f'<cvParam cvRef="MS" accession="MS:1000744" name="selected ion m/z" value="{s["prec_mz"]!r}"/>')
            if "charge" in s:
                out.append(f'<cvParam cvRef="MS" accession="MS:1000041" name="charge state" value="{s["charge"]}"/>')
            if "possible_charge" in s:
                out.append(f'<cvParam cvRef="MS" accession="MS:1000633" name="possible charge state" '
                           f'value="{s["possible_charge"]}"/>')
            out.append('</selectedIon></selectedIonList><activation><cvParam cvRef="MS" accession="MS:1000422" '
                       'name="beam-type collision-induced dissociation" value=""/></activation></precursor>')
            if s.get("extra_precursor"):
                out.append('<precursor><selectedIonList count="1"><selectedIon><cvParam cvRef="MS" '
                           'accession="MS:1000744" name="selected ion m/z" value="999.9"/><cvParam cvRef="MS" '
                           'accession="MS:1000041" name="charge state" value="7"/></selectedIon></selectedIonList>'
                           '</precursor>')
            out.append('</precursorList>')
        out.append('<binaryDataArrayList count="2">')
        out.append(_binary(s["mz"], s.get("bits", 64), s.get("zlib", False), "mz"))
        out.append(_binary(s["intensity"], 32 if s.get("bits", 64) == 32 else s.get("int_bits", 32), s.get("zlib", False),
                           "intensity"))
        out.append('</binaryDataArrayList></spectrum>')
    out += ['</spectrumList>', '</run>', '</mzML>']
    with open(path, "w") as f:
        f.write("\n".join(out))


def read_mzml(path: str) -> List[dict]:
    out = []
    root = ET.parse(path).getroot()
    for index, sp in enumerate(root.iter(NS + "spectrum")):
        cv = {c.get("accession"): c.get("value") for c in sp.findall(NS + "cvParam")}
        if int(cv.get("MS:1000511", -1)) != 2:
            continue
        sid = sp.get("id")
        try:
            if "scan=" in sid:
                scan = int(sid[sid.find("scan=") + 5:])
            elif "index=" in sid:
                scan = int(sid[sid.find("index=") + 6:])
            else:
                raise ValueError
        except ValueError:
            continue
        ion = sp.find(f"{NS}precursorList/{NS}precursor/{NS}selectedIonList/{NS}selectedIon")
        if ion is None:
            continue
        icv = {c.get("accession"): c.get("value") for c in ion.findall(NS + "cvParam")}
        scan_el = sp.find(f"{NS}scanList/{NS}scan")
        scv = {c.get("accession"): c.get("value") for c in scan_el.findall(NS + "cvParam")} if scan_el is not None else {}
        arrays = {}
        for bda in sp.iter(NS + "binaryDataArray"):
            bcv = {c.get("accession") for c in bda.findall(NS + "cvParam")}
            raw = base64.b64decode(bda.find(NS + "binary").text or "")
            if "MS:1000574" in bcv and raw:
                raw = zlib.decompress(raw)
            a = np.frombuffer(raw, "<f8" if "MS:1000523" in bcv else "<f4").astype(np.float64)
            arrays["mz" if "MS:1000514" in bcv else "inten"] = a
        order = np.argsort(arrays["mz"], kind="stable")
        z = int(icv["MS:1000041"]) if "MS:1000041" in icv else int(icv.get("MS:1000633", 0))
        out.append(dict(identifier=str(scan), index=index, prec_mz=float(icv["MS:1000744"]), prec_z=z,
                        rt=float(scv["MS:1000016"]) if "MS:1000016" in scv else math.nan,
                        mz=arrays["mz"][order], inten=arrays["inten"].astype(np.float32)[order]))
    return out


# ---------------------------------------------------------------------- mzXML
def write_mzxml(path: str, scans: List[dict]) -> None:
    """scans: dicts with num, ms_level, mz, intensity and optionally prec_mz, charge, rt (seconds),
    precision (32/64), zlib (bool), children (list of scans nested inside this one, as mzXML does for
    the MS2 scans of a survey scan)."""
    def emit(s, out):
        prec = s.get("precision", 32)
        inter = np.empty(2 * len(s["mz"]), ">f8" if prec == 64 else ">f4")
        inter[0::2] = s["mz"]
        inter[1::2] = s["intensity"]
        raw = inter.tobytes()
        comp = "none"
        if s.get("zlib"):
            raw = zlib.compress(raw)
            comp = "zlib"
        rt = f' retentionTime="PT{s["rt"]!r}S"' if "rt" in s else ""
        out.append(f'<scan num="{s["num"]}" msLevel="{s["ms_level"]}" peaksCount="{len(s["mz"])}"{rt}>')
        if "prec_mz" in s:
            z = f' precursorCharge="{s["charge"]}"' if "charge" in s else ""
            out.append(f'<precursorMz precursorIntensity="1234.5"{z}>{s["prec_mz"]!r}</precursorMz>')
        out.append(f'<peaks precision="{prec}" byteOrder="network" contentType="m/z-int" compressionType="{comp}" '
                   f'compressedLen="{len(raw) if s.get("zlib") else 0}">{base64.b64encode(raw).decode()}</peaks>')
        for child in s.get("children", ()):
            emit(child, out)
        out.append("</scan>")

    out = ['<?xml version="1.0" encoding="ISO-8859-1"?>', '<mzXML xmlns="http://sashimi.sourceforge.net/schema_revision/mzXML_3.2">',
           '<msRun scanCount="0">']
    for s in scans:
        emit(s, out)
    out += ["</msRun>", "</mzXML>"]
    with open(path, "w") as f:
        f.write("\n".join(out))


def read_mzxml(path: str) -> List[dict]:
    ns = "{http://sashimi.sourceforge.net/schema_revision/mzXML_3.2}"
    out = []
    root = ET.parse(path).getroot()
    for index, sc in enumerate(root.iter(ns + "scan")):       # document order, nested scans included
        if int(sc.get("msLevel", -1)) != 2:
            continue
        pm = sc.find(ns + "precursorMz")
        if pm is None:
            continue
        pk = sc.find(ns + "peaks")
        raw = base64.b64decode(pk.text or "")
        if pk.get("compressionType") == "zlib" and raw:
            raw = zlib.decompress(raw)
        a = np.frombuffer(raw, ">f8" if pk.get("precision") == "64" else ">f4").astype(np.float64)
        mz, inten = a[0::2], a[1::2]
        order = np.argsort(mz, kind="stable")
        rt = sc.get("retentionTime")
        out.append(dict(identifier=str(int(sc.get("num"))), index=index, prec_mz=float(pm.text),
                        prec_z=int(pm.get("precursorCharge", 0)), rt=float(rt[2:-1]) / 60.0 if rt else math.nan,
                        mz=mz[order], inten=inten.astype(np.float32)[order]))
    return out

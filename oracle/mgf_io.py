"""TEST INFRASTRUCTURE ONLY (oracle): pure-Python restatement of how the reference reads MGF query
files (reference reader.py:868-911 through pyteomics.mgf, which is absent: **parity unpinned**), and a
writer for MassIVE-KB style entries like the reference's own test builds
(src/tests/query_reader_test.py:41-66: BEGIN IONS / SEQ= / PEPMASS= / CHARGE=2+ / peaks / END IONS).
"""
from __future__ import annotations

import math
import re
from typing import List

import numpy as np


def write_mgf(path: str, spectra: List[dict], header=("COM=synthetic", "# a comment")) -> None:
    """spectra: dicts with prec_mz, mz, intensity and optionally title, scan, charge (str as written),
    rt, seq, decoy, extra (raw lines), peak_charge (third column)."""
    lines = list(header)
    for s in spectra:
        lines.append("BEGIN IONS")
        if "seq" in s:
            lines.append(f"SEQ={s['seq']}")
        if "title" in s:
            lines.append(f"TITLE={s['title']}")
        if "scan" in s:
            lines.append(f"SCANS={s['scan']}")
        lines.append(f"PEPMASS={s['prec_mz']!r}" + (" 12345.6" if s.get("pepmass_intensity") else ""))
        if "charge" in s:
            lines.append(f"CHARGE={s['charge']}")
        if "rt" in s:
            lines.append(f"RTINSECONDS={s['rt']!r}")
        if s.get("decoy"):
            lines.append("DECOY=1")
        lines += list(s.get("extra", ()))
        for k, (m, i) in enumerate(zip(s["mz"], s["intensity"])):
            lines.append(f"{float(m)!r} {float(i)!r}" + (" 1+" if s.get("peak_charge") and k % 2 else ""))
        lines.append("END IONS")
        lines.append("")
    with open(path, "w") as f:
        f.write("\n".join(lines))


def read_mgf(path: str) -> List[dict]:
    out, cur = [], None
    for raw in open(path):
        line = raw.strip()
        if not line:
            continue
        if cur is None:
            if line.upper() == "BEGIN IONS":
                cur = dict(params={}, mz=[], inten=[])
            continue
        if line.upper() == "END IONS":
            p = cur["params"]
            ident = p.get("title", p.get("scan", p.get("scans", str(len(out) + 1))))
            z = 0
            if "charge" in p:
                m = re.match(r"\s*([+-]?)(\d+)([+-]?)", p["charge"])
                if m:
                    z = int(m.group(2)) * (-1 if "-" in (m.group(1), m.group(3)) else 1)
            mz = np.array(cur["mz"], np.float64)
            inten = np.array(cur["inten"], np.float64).astype(np.float32)
            order = np.argsort(mz, kind="stable")
            out.append(dict(identifier=ident, prec_mz=float(p["pepmass"].split()[0]), prec_z=z,
                            rt=float(p["rtinseconds"]) if "rtinseconds" in p else math.nan,
                            is_decoy="decoy" in p, seq=p.get("seq", ""), mz=mz[order], inten=inten[order]))
            cur = None
            continue
        if line[0] in "#;!/":
            continue
        if "=" in line:
            k, v = line.split("=", 1)
            cur["params"][k.strip().lower()] = v.strip()
            continue
        f = line.split()
        cur["mz"].append(float(f[0]))
        cur["inten"].append(float(f[1]))
    assert cur is None, "BEGIN IONS without END IONS"
    return out

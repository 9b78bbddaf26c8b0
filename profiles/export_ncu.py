#!/usr/bin/env python
"""Turns the ncu reports of profiles/profile.sh into what is committed (run here, no GPU needed):

    python profiles/export_ncu.py r2

  profiles/<tag>_kernels_raw.csv   the counters quoted in DESIGN.md / the summary, one row per profiled launch
  profiles/<tag>_launches.csv      the launch list (copied)
  profiles/traffic.json            DRAM bytes per launch of the dominant kernels, with the build's git hash — what
                                   bench.py puts into roofline.traffic
  profiles/<tag>_sass_scan_tc.txt  cuobjdump excerpt: the tcgen05 / TMEM / TMA instructions of the scan kernel
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_kernels.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(header) if h in ("ID", "Kernel Name") or any(h == k or h.startswith(k) for k in KEEP)]
    out = os.path.join(ROOT, "profiles", f"{tag}_kernels_raw.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([header[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in data:
            w.writerow([re.sub(r"\(.*", "", r[i]) if header[i] == "Kernel Name" else r[i] for i in cols])
    print("wrote", out, len(data), "launches")
    # DRAM traffic per launch of the kernels bench.py reports a roofline for
    hi = {h: i for i, h in enumerate(header)}
    unit = {h: units[i] for h, i in hi.items()}

    def to_bytes(v, u):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return float(v.replace(",", "")) * scale

    acc = {}
    for r in data:
        full = re.sub(r"^void\s+", "", r[hi["Kernel Name"]])
        name = re.sub(r"[<(].*", "", full)
        key = name
        if name == "scan_tc_kernel":
            key = "k2_coarse" if re.search(r"scan_tc_kernel<\(int\)1|scan_tc_kernel<1", full) or ", 7>" in full or "(int)7>" in full else "k3_list_scan"
        b = to_bytes(r[hi["dram__bytes_read.sum"]], unit["dram__bytes_read.sum"]) + to_bytes(r[hi["dram__bytes_write.sum"]], unit["dram__bytes_write.sum"])
        acc.setdefault(key, []).append(b)
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    traffic = {"c2": {"git": git, "source": f"profiles/{tag}_kernels_raw.csv (ncu --set full, charge-2 launches of one C2 step)",
                      "k3_list_scan_dram_bytes_per_launch": int(sum(acc.get("k3_list_scan", [0])) / max(1, len(acc.get("k3_list_scan", [])))),
                      "per_kernel_dram_bytes_per_launch": {k: int(sum(v) / len(v)) for k, v in sorted(acc.items())},
                      "launches_profiled": {k: len(v) for k, v in sorted(acc.items())}}}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(json.dumps(traffic, indent=1))
    src = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if os.path.isfile(src):
        with open(src) as f, open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as g:
            g.write(f.read())
    # SASS evidence
    so = os.path.join(ROOT, "ann-solo_b200", "libsolo_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out_lines, fn, counts = [], None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
        if fn and "scan_tc" in fn and re.search(r"UTCHMMA|UTCBAR|LDTM|UTMALDG|UTCATOMSWS|SYNCS|LDGSTS", line):
            op = re.search(r"(UTCHMMA[.\w]*|UTCBAR[.\w]*|LDTM[.\w]*|UTMALDG[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|LDGSTS[.\w]*)", line).group(1)
            counts.setdefault(fn, {}).setdefault(op, 0)
            counts[fn][op] += 1
            if len(out_lines) < 120 and ("UTCHMMA" in line or "LDTM" in line or "UTMALDG" in line):
                out_lines.append(f"{fn[:60]}: {line.strip()}")
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_scan_tc.txt"), "w") as f:
        f.write(f"cuobjdump -sass ann-solo_b200/libsolo_b200.so (git {git}): tensor-core / TMEM / TMA instructions per scan_tc kernel\n\n")
        for k, v in sorted(counts.items()):
            f.write(f"{k}\n   " + "  ".join(f"{op} x{n}" for op, n in sorted(v.items())) + "\n")
        f.write("\nfirst occurrences:\n" + "\n".join(out_lines) + "\n")
    print("kernels with tcgen05 SASS:", len(counts))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Writes profiles/current.json from an ``ncu --set full`` report of scripts/profile_kernels.py
(N items per launch): warp instructions and DRAM bytes per item of the first-tier thal kernels,
stamped with the hash of the library sources.  bench.py reports these figures only while the
sources still hash to the same value -- a kernel edit silently invalidates them otherwise.

usage: python scripts/make_current_profile.py report.ncu-rep n_items capture_name"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from polyoligo_b200 import build as po_build

rep, n, capture = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    return float(r[col[name]].replace(",", ""))


def dram(r):
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        unit = rows[1][col[name]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        tot += num(r, name) * scale
    return tot


dimer = hp = None
for r in rows[2:]:
    k = r[col["Kernel Name"]]
    rec = {"inst": num(r, "smsp__inst_executed.sum"), "dram": dram(r), "ms": num(r, "gpu__time_duration.sum"),
           "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "registers": num(r, "launch__registers_per_thread")}
    if k.startswith("void k_thal<256,") and dimer is None:
        dimer = rec
    if k.startswith("void k_thal<128,"):
        if hp is None:
            hp = {"inst": 0.0, "dram": 0.0, "ms": 0.0, "parts": []}
        if len(hp["parts"]) < 2:          # fill + finish of one call
            hp["inst"] += rec["inst"]; hp["dram"] += rec["dram"]; hp["ms"] += rec["ms"]; hp["parts"].append(rec)
out = {"source_hash": po_build.source_hash(), "capture": capture, "items_per_launch": n,
       "dimer": {"warp_instructions_per_pair": dimer["inst"] / n, "dram_bytes_per_pair": dimer["dram"] / n,
                 "kernel_ms": dimer["ms"], "issue_active_pct": dimer["issue_active_pct"], "registers": dimer["registers"]},
       "hairpin": {"warp_instructions_per_oligo": hp["inst"] / n, "dram_bytes_per_oligo": hp["dram"] / n,
                   "kernel_ms": hp["ms"], "issue_active_pct": [p["issue_active_pct"] for p in hp["parts"]]}}
with open(os.path.join(ROOT, "profiles", "current.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))

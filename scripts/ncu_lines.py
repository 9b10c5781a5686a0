"""Executed warp instructions of one kernel per source line of one file.
usage: python scripts/ncu_lines.py report.ncu-rep kernel_substr file.cuh lo hi"""
import csv, collections, subprocess, sys, io
rep, kern_sub, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = kern = hdr2 = None
agg = collections.Counter(); text = {}; tot = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2 and r[0] == 'Function Name': kern = r[1]; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No': hdr2 = r; continue
    if hdr2 is None or kern is None or kern_sub not in kern: continue
    try:
        ln = int(r[0]); inst = float(r[hdr2.index('Instructions Executed')] or 0)
    except (ValueError, IndexError):
        continue
    tot += inst
    if cur == fname and lo <= ln <= hi:
        agg[ln] += inst; text[ln] = r[1][:110]
for ln in sorted(agg):
    print("%5d %6.2f%%  %s" % (ln, 100 * agg[ln] / tot, text[ln]))

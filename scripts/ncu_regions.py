"""Aggregate executed warp instructions of one kernel by source-line regions.
usage: python scripts/ncu_regions.py report.ncu-rep kernel_substr file:lo-hi=name ..."""
import csv, collections, subprocess, sys, io
rep, kern_sub = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    spec, name = a.split("=")
    f, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None; kern = None; hdr2 = None
agg = collections.Counter(); smp = collections.Counter(); perfile = collections.Counter()
tot = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2 and r[0] == 'Function Name': kern = r[1]; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No': hdr2 = r; continue
    if hdr2 is None or kern is None or kern_sub not in kern: continue
    try:
        ln = int(r[0]); inst = float(r[hdr2.index('Instructions Executed')] or 0); s = float(r[hdr2.index('# Samples')] or 0)
    except (ValueError, IndexError):
        continue
    tot += inst
    name = None
    for f, lo, hi, nm in regions:
        if cur == f and lo <= ln <= hi:
            name = nm; break
    if name is None:
        name = "other:" + str(cur)
    agg[name] += inst; smp[name] += s
print("kernel contains %r: total warp-inst %.4g" % (kern_sub, tot))
ts = sum(smp.values())
for k, v in agg.most_common():
    print("%6.2f%% inst %6.2f%% smp  %s" % (100 * v / tot, 100 * smp[k] / max(ts, 1), k))

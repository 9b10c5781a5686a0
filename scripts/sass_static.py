"""Static SASS instruction count of one kernel by source file:line regions (code footprint).
usage: python scripts/sass_static.py lib.so kernel_mangled_substr [file:lo-hi=name ...]"""
import collections, re, subprocess, sys, tempfile, os, glob
lib, sub = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    spec, name = a.split("=")
    f, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cubin = [c for c in glob.glob(d + "/*.cubin") if "tables" not in c][0]
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
inside = False; cur = ("?", 0); cnt = collections.Counter(); per_line = collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        inside = sub in m.group(1); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        name = None
        for f, lo, hi, nm in regions:
            if cur[0] == f and lo <= cur[1] <= hi: name = nm; break
        cnt[name or "other:" + cur[0]] += 1
        per_line[cur] += 1
tot = sum(cnt.values())
print("static SASS instructions: %d (%d KB)" % (tot, tot * 16 // 1024))
for k, v in cnt.most_common(): print("%6d %5.1f%%  %s" % (v, 100 * v / tot, k))
print("top lines:")
for (f, l), v in per_line.most_common(25): print("%6d  %s:%d" % (v, f, l))

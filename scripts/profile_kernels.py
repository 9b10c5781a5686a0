"""Launches each thal kernel a few times on a config-4 shaped batch (for ncu captures)."""
import sys
sys.path.insert(0, ".")
import time
import polyoligo_b200 as po
from polyoligo_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = po.Context(0)
a, b = synth.config4(n)
for _ in range(reps):
    t = time.time(); ctx.thal_batch(po.THAL_ANY, a, b); t1 = time.time() - t
    t = time.time(); ctx.thal_batch(po.THAL_HAIRPIN, a); t2 = time.time() - t
print("het %.2f M/s  hairpin %.2f M/s (incl. pageable copies)" % (n / t1 / 1e6, n / t2 / 1e6))

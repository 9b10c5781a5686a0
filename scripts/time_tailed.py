"""KASP-tailed shape (BASELINE 8d secondary): reporter (21 nt) + 18-25 nt primer vs 18-25 nt primer."""
import sys, time
import numpy as np
sys.path.insert(0, "tests")
import polyoligo_b200 as po
from helpers import rand_seqs, VIC, FAM
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
rng = np.random.default_rng(5)
core = rand_seqs(rng, n, 18, 25)
com = rand_seqs(rng, n, 18, 25)
tailed = [(VIC if k % 2 else FAM) + s for k, s in enumerate(core)]
a, b = po.pack(tailed), po.pack(com)
ctx = po.Context(0)
ctx.thal_batch(po.THAL_ANY, a[:1000], b[:1000])
for rep in range(2):
    t0 = time.perf_counter(); out = ctx.thal_batch(po.THAL_ANY, a, b); dt = time.perf_counter() - t0
print("tailed heterodimers: %d in %.3f s = %.2f M/s (host buffers)" % (n, dt, n / dt / 1e6))
w = a["w"]; 

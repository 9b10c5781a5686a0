"""One-off soak: GPU thal vs the CPU oracle, bit-exact, on a few million random inputs."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import polyoligo_b200 as po
from polyoligo_b200 import synth
from oracle.oracle import Oracle
from helpers import rand_seqs, VIC, FAM
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ctx, orc = po.Context(0), Oracle()
threads = 16
def check(name, got, ref):
    ok = all(np.array_equal(got[f], ref[f]) for f in ("tm", "dg", "dh", "ds")) and \
        np.array_equal(got["flags"] & 1, ref["structure_found"])
    print("%-28s %9d items  bit-exact: %s" % (name, len(got), ok), flush=True)
    return ok
ok = True
a, b = synth.config4(n, seed=424242)
sa, sb = synth.to_strings(a), synth.to_strings(b)
t0 = time.time()
ok &= check("heterodimer (config 4)", ctx.thal_batch(po.THAL_ANY, a, b), orc.thal_batch(sa, sb, 1, nthreads=threads))
ok &= check("hairpin (config 4)", ctx.thal_batch(po.THAL_HAIRPIN, a), orc.thal_batch(sa, None, 4, nthreads=threads))
m = n // 4
ok &= check("END1", ctx.thal_batch(po.THAL_END1, a[:m], b[:m]), orc.thal_batch(sa[:m], sb[:m], 2, nthreads=threads))
ok &= check("END2", ctx.thal_batch(po.THAL_END2, a[:m], b[:m]), orc.thal_batch(sa[:m], sb[:m], 3, nthreads=threads))
ok &= check("homodimer", ctx.thal_batch(po.THAL_ANY, b[:m]), orc.thal_batch(sb[:m], sb[:m], 1, nthreads=threads))
rng = np.random.default_rng(9)
core, com = rand_seqs(rng, m, 18, 26), rand_seqs(rng, m, 18, 25)
tailed = [(VIC if k % 2 else FAM) + s for k, s in enumerate(core)]
ok &= check("KASP-tailed heterodimer", ctx.thal_batch(po.THAL_ANY, po.pack(tailed), po.pack(com)),
            orc.thal_batch(tailed, com, 1, nthreads=threads))
print("total %.0f s, all bit-exact: %s" % (time.time() - t0, ok))
sys.exit(0 if ok else 1)

"""Stage timing of the config-5 array pipeline (host wall clock per C-ABI call)."""
import sys, time, collections
import numpy as np
from polyoligo_b200 import kasp_pipeline as kp, lib_primer3 as lp, primer3_bindings as p3b, _native
from polyoligo_b200.thermoanalysis import get_context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
workers = int(sys.argv[3]) if len(sys.argv) > 3 else 1
p3b.resetP3Globals(); lp.PRIMER3_GLOBALS.clear()
lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0, PRIMER_MIN_TM=55.0,
               PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0, PRIMER_MAX_POLY_X=100,
               PRIMER_SALT_MONOVALENT=50.0, PRIMER_DNA_CONC=50.0, PRIMER_PAIR_MAX_DIFF_TM=6.0,
               PRIMER_PRODUCT_SIZE_RANGE=[[90, 150], [100, 200], [50, 250]], PRIMER_NUM_RETURN=100)
lp.set_tm_delta(5.0)
ctx = get_context(0)
T = collections.OrderedDict()
for name in ("kasp_enumerate", "oligo_props_batch", "p3_design_fixed", "annotate_mutations", "rank_pairs"):
    fn = getattr(ctx, name)
    def wrap(fn=fn, name=name):
        def w(*a, **k):
            t0 = time.perf_counter(); r = fn(*a, **k); T[name] = T.get(name, 0) + time.perf_counter() - t0; return r
        return w
    setattr(ctx, name, wrap())
dyes = ("GAAGGTCGGAGTCAACGGATT", "GAAGGTGACCAAGTTCATGCT")
kp.design_markers_arrays(kp.synth_config5(n, seed=1), n_primers=10, reporters=dyes, chunk=chunk, workers=workers)   # same-size warm-up
T.clear()
data = kp.synth_config5(n)
t0 = time.perf_counter()
res = kp.design_markers_arrays(data, n_primers=10, reporters=dyes, chunk=chunk, workers=workers)
tot = time.perf_counter() - t0
print("total %.3f s  %.0f markers/s" % (tot, n / tot))
for k, v in T.items(): print("  %-22s %.3f s" % (k, v))
print("  %-22s %.3f s" % ("numpy glue", tot - sum(T.values())))

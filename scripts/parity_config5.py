"""One-off: config-5 array pipeline, GPU context vs CPU oracle context on N markers (bit-exact arrays)."""
import sys, time
import numpy as np
from polyoligo_b200 import kasp_pipeline as kp, lib_primer3 as lp, primer3_bindings as p3b
from oracle.cpu_pipeline import CpuContext
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 99
p3b.resetP3Globals(); lp.PRIMER3_GLOBALS.clear()
lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0, PRIMER_MIN_TM=55.0,
               PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0, PRIMER_MAX_POLY_X=100,
               PRIMER_SALT_MONOVALENT=50.0, PRIMER_DNA_CONC=50.0, PRIMER_PAIR_MAX_DIFF_TM=6.0,
               PRIMER_PRODUCT_SIZE_RANGE=[[90, 150], [100, 200], [50, 250]], PRIMER_NUM_RETURN=100)
lp.set_tm_delta(5.0)
dyes = ("GAAGGTCGGAGTCAACGGATT", "GAAGGTGACCAAGTTCATGCT")
data = kp.synth_config5(n, seed=seed)
gpu = kp.design_markers_arrays(data, n_primers=10, reporters=dyes, chunk=64)
t0 = time.perf_counter()
cpu = kp.design_markers_arrays(data, n_primers=10, reporters=dyes, chunk=64, ctx=CpuContext())
print("cpu %.1f s" % (time.perf_counter() - t0))
ok = all(np.array_equal(gpu[k], cpu[k]) for k in ("seg", "dir", "goodness", "qbits", "sel", "n_sel")) and \
    all(gpu[k].tobytes() == cpu[k].tobytes() for k in ("f", "r", "a"))
print("markers %d, merged pairs %d, selected %d, identical: %s" % (n, len(gpu["f"]), int(gpu["n_sel"].sum()), ok))
sys.exit(0 if ok else 1)

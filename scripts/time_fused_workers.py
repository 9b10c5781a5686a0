"""Config-5 design through po_kasp_design_batch: one host thread vs two (blocks of 4096 markers on two
contexts / streams), for a full-size set and for an 8-GPU shard of it (host wall clock)."""
import sys, time
sys.path.insert(0, ".")
import bench
from polyoligo_b200 import kasp_pipeline as kp
bench.set_kasp_globals()
kp.design_markers_fused(kp.synth_config5(8192, seed=1), n_primers=10, reporters=bench.DYES)
for n in [int(a) for a in sys.argv[1:]] or [12500, 100000]:
    data = kp.synth_config5(n, seed=kp.CONFIG5_SEED)
    for workers in (1, 2, 3):
        best = 1e9
        for _ in range(2):
            t = time.perf_counter()
            kp.design_markers_fused(data, n_primers=10, reporters=bench.DYES, workers=workers)
            best = min(best, time.perf_counter() - t)
        print("n %6d workers %d  %.3f s  %.0f markers/s" % (n, workers, best, n / best), flush=True)

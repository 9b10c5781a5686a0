"""Times the config-5 design of N markers: array pipeline vs the single fused C call (host wall clock)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bench
from polyoligo_b200 import kasp_pipeline as kp, sharding
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
bench.set_kasp_globals()
data = kp.synth_config5(n, seed=kp.CONFIG5_SEED)
kp.design_markers_fused(kp.synth_config5(2048, seed=1), n_primers=10, reporters=bench.DYES)
kp.design_markers_arrays(kp.synth_config5(2048, seed=1), n_primers=10, reporters=bench.DYES)
t = time.perf_counter(); rec = kp.marker_records(data["templates"], data["marker_idx"], data["ref_len"], data["alt"]); t1 = time.perf_counter() - t
t = time.perf_counter(); mb = kp.pack_maps(data["maps"]); t2 = time.perf_counter() - t
print("host packing: marker records %.3f s, maps %.3f s" % (t1, t2))
for name, fn in (("fused", lambda: kp.design_markers_fused(data, n_primers=10, reporters=bench.DYES)),
                 ("arrays", lambda: kp.design_markers_arrays(data, n_primers=10, reporters=bench.DYES))):
    ctx = kp.get_context(None)
    i0 = ctx.thal_item_count(); l0 = ctx.launch_count()
    t = time.perf_counter(); r = fn(); dt = time.perf_counter() - t
    print("%-7s %.3f s  %.0f markers/s  launches %d thal items %d" % (name, dt, n / dt, ctx.launch_count() - l0, ctx.thal_item_count() - i0))

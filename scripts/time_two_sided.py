"""Two-sided picking (both primers free, target in the middle) for many templates per call."""
import sys, time
import numpy as np
from polyoligo_b200 import primer3_bindings as p3b
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(1)
p3b.resetP3Globals()
p3b.setP3Globals({"PRIMER_MIN_SIZE": 18, "PRIMER_OPT_SIZE": 20, "PRIMER_MAX_SIZE": 25, "PRIMER_MIN_TM": 55.0,
                  "PRIMER_OPT_TM": 60.0, "PRIMER_MAX_TM": 65.0, "PRIMER_MAX_POLY_X": 100, "PRIMER_PAIR_MAX_DIFF_TM": 6.0,
                  "PRIMER_PRODUCT_SIZE_RANGE": [[40, 75], [40, 150]], "PRIMER_NUM_RETURN": 100})
args = [{"SEQUENCE_TEMPLATE": "".join(np.array(list("ACGT"))[rng.integers(0, 4, 500)]), "SEQUENCE_TARGET": [249, 1]}
        for _ in range(n)]
p3b.design_primers_batch(args[:8])
t0 = time.perf_counter()
res = p3b.design_primers_batch(args)
dt = time.perf_counter() - t0
print("%d two-sided designs (HRM size ranges, depth 100) in %.2f s = %.1f designs/s; pairs returned %d"
      % (n, dt, n / dt, sum(r["PRIMER_PAIR_NUM_RETURNED"] for r in res)))

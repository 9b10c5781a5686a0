#!/bin/bash
# round-2 development loop on the GPU box: thal parity (tests + soak), then kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_thal_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_thal.log
timeout 900 python scripts/soak_parity.py ${SOAK:-300000} 2>&1 | tail -9 | tee gpurun_out/soak.log
for lib in polyoligo_b200/libpolyoligo_b200.so $ALT_LIBS; do
  echo "== $lib"
  POLYOLIGO_B200_LIB=$PWD/$lib timeout 600 python bench.py --pairs 3000000 --steps 3 --warmup 3 --no-cpu-baseline --markers 0 2> gpurun_out/bench_3m.err | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['kernel_ms'], 'value %.4g het/s %.4g hp/s %.4g'%(d['value'], d['heterodimers_per_s'], d['hairpins_per_s']))
except Exception as e:
    print('bench failed', e)
"
done
tail -3 gpurun_out/bench_3m.err

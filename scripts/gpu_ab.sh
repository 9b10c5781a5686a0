#!/bin/bash
# A/B library builds on the same box: kernel times from bench.py CUDA events (3M config-4 pairs)
mkdir -p gpurun_out
for lib in $ALT_LIBS; do
  echo "== $lib"
  POLYOLIGO_B200_LIB=$PWD/$lib timeout 600 python bench.py --pairs ${PAIRS:-3000000} --steps 3 --warmup 3 --no-cpu-baseline --markers 0 2> gpurun_out/bench_ab.err | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['kernel_ms'], 'value %.4g het/s %.4g hp/s %.4g'%(d['value'], d['heterodimers_per_s'], d['hairpins_per_s']))
except Exception as e:
    print('bench failed', e)
"
done
tail -3 gpurun_out/bench_ab.err

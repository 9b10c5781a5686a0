#!/bin/bash
# A/B two library builds on the same box: kernel times from bench.py CUDA events
for lib in "$@"; do
  echo "== $lib"
  POLYOLIGO_B200_LIB=$lib python bench.py --pairs 3000000 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['kernel_ms'], 'het/s %.3g hp/s %.3g'%(d['heterodimers_per_s'], d['hairpins_per_s']))"
done

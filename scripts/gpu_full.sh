#!/bin/bash
# full GPU check: the whole gpu test suite, smoke, then the default bench and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -c 700 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err

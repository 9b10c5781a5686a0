#!/bin/bash
# GPU box script: tests, smoke, bench, then ncu (only after the plain run exited 0).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
# ncu: launch list on a reduced batch (same code path), then one full capture of the dimer kernel
python bench.py --steps 1 --warmup 3 --pairs 400000 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --pairs 400000 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
python bench.py --steps 1 --warmup 3 --pairs 400000 --no-cpu-baseline > gpurun_out/plain_small2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_thal -s 8 -c 2 -o gpurun_out/prof_r1 \
    python bench.py --steps 1 --warmup 3 --pairs 400000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out

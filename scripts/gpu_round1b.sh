#!/bin/bash
# GPU box script (round 1, second half): tests, smoke, default bench, reference arm, then ncu
# (each ncu command only after the same command exited 0 without it; config-5 part disabled under ncu).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 4500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
SMALL="--steps 1 --warmup 3 --pairs 400000 --no-cpu-baseline --markers 0"
python bench.py $SMALL > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
python bench.py $SMALL > gpurun_out/plain_small2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_thal -s 28 -c 5 -f -o gpurun_out/prof_r1_v10 \
    python bench.py $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -12

#!/bin/bash
# full default bench (10^7 pairs), the reference arm, then launch list + full capture on a reduced batch
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
python bench.py --steps 1 --warmup 3 --pairs 400000 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --pairs 400000 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"

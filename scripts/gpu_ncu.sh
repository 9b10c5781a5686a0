#!/bin/bash
# one ncu --set full capture of the first-tier thal kernels on a 400k-pair batch (after a plain run exited 0)
mkdir -p gpurun_out
TAG=${TAG:-r2}
python scripts/profile_kernels.py 400000 2 > gpurun_out/plain_profile.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_profile.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_thal -s 7 -c 7 -f -o gpurun_out/prof_$TAG \
    python scripts/profile_kernels.py 400000 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1
head -c 6000 gpurun_out/${TAG}_ncu_full_summary.txt

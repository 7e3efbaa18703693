#!/bin/bash
# Round-2 ncu captures (run under gpurun on ONE B200); summaries go to profiles/ with tools/summarize_ncu.py.
set -x
mkdir -p gpurun_out
C3="--streams 11234 --queries 10 --places 10000 --seq-len 10"     # one full stream group of config 3
# 1. launch list of the bench command itself (shares of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.json 2> gpurun_out/r02_launches.err
# 2. full captures: hidden + output tensor-core kernels (3rd step), matching kernel, binning kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:output_tc_kernel -s 4 -c 2 -f -o gpurun_out/r02_k2k3 \
    python tools/profile_step.py --steps 3 $C3 > gpurun_out/r02_k2k3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:seqmatch_topk -s 2 -c 1 -f -o gpurun_out/r02_k4 \
    python tools/profile_step.py --steps 3 $C3 > gpurun_out/r02_k4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bin_pow2 -s 2 -c 1 -f -o gpurun_out/r02_k1 \
    python tools/profile_step.py --steps 3 --events 1000000000 > gpurun_out/r02_k1.log 2>&1
# 3. config 2 (the round-1 workload) for continuity
timeout 600 ncu --set full --clock-control none --import-source on -k regex:output_tc_kernel -s 4 -c 2 -f -o gpurun_out/r02_k2k3_cfg2 \
    python tools/profile_step.py --steps 3 --streams 1000 --queries 16 --places 1000 --seq-len 2 > gpurun_out/r02_k2k3_cfg2.log 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# Full ncu capture of one kernel: bash profiles/gpu_prof.sh <encode|decode> <tag> [extra bench args]
K=$1; TAG=$2; shift 2
CMD="python bench.py --steps 3 --warmup 3 --preheat 0 --no-cpu --no-batched $@"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fri_$K -s 3 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log

#!/bin/bash
# round-1 late experiments: cleaned-up decode (VIMNMX clamp, hoisted loads) vs 512-thread CTAs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_b.log 2>&1; tail -3 gpurun_out/pytest_b.log
bash profiles/variants2.sh "base 0 2" "t512 4x4 1"

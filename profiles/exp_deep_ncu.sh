#!/bin/bash
# ncu --set full of the main transform kernels at 16384^2 u16: depth 9 against depth 16 (deep-tree variant)
mkdir -p gpurun_out
for D in 9 16; do
  CMD="python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --depth $D --divisor 1 --reps 2 --sets 2"
  for K in encode decode; do
    ncu --set full --clock-control none --import-source on -k regex:fri_${K}_kernel -s 2 -c 1 -o gpurun_out/prof_deep_d${D}_$K -f $CMD > gpurun_out/ncu_deep_d${D}_$K.log 2>&1
  done
done
ls -la gpurun_out/prof_deep_*

#!/bin/bash
# Small groups / small CTAs for the RGB kernels: 128-thread CTAs (8 per SM) instead of 256 (4 per SM)
mkdir -p gpurun_out
out=gpurun_out/exp_group_small.txt; : > $out
for g in 4x4 4x2 2x4 3x3; do
  echo "== group $g" >> $out
  FRI_GROUP=$g python profiles/exp_b2b.py --reps 400 --tag "g$g" >> $out 2>&1
done
for g in 4x4 4x2; do
  echo "== batch32 group $g" >> $out
  FRI_GROUP=$g python profiles/exp_b2b.py --frames 32 --reps 20 --sets 2 --tag "b32g$g" >> $out 2>&1
done
cut -c1-330 $out

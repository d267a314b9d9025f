#!/bin/bash
# Group-shape sweep (FRI_GROUP=AxB): wave quantisation of the single-frame launch vs per-group efficiency.
mkdir -p gpurun_out
out=gpurun_out/exp_group.txt; : > $out
for g in 4x4 5x3 3x5 4x3 3x4 6x2 3x6 6x3 7x2 4x4; do
  echo "== group $g" >> $out
  FRI_GROUP=$g python profiles/exp_b2b.py --reps 400 --tag "g$g" >> $out 2>&1
done
for g in 4x4 5x3 3x5 4x3; do
  echo "== batch32 group $g" >> $out
  FRI_GROUP=$g python profiles/exp_b2b.py --frames 32 --reps 20 --sets 2 --tag "b32g$g" >> $out 2>&1
done
cat $out

#!/bin/bash
# A/B of the decoder's hand-off prefetch (FRI_DEC_HANDOFF = distance in CTAs, FRI_DEC_HANDOFF_TILES = tiles pulled).
mkdir -p gpurun_out
out=gpurun_out/exp_handoff.txt; : > $out
for rep in 1 2; do
for cfg in "0 8" "592 8" "592 16" "296 8" "444 8" "740 8" "592 4" "1184 8"; do
  set -- $cfg
  echo "== handoff=$1 tiles=$2 rep=$rep" >> $out
  FRI_DEC_HANDOFF=$1 FRI_DEC_HANDOFF_TILES=$2 python profiles/exp_b2b.py --reps 400 --tag "h$1t$2" >> $out 2>&1
done; done
echo "== batch 32" >> $out
for cfg in "0 8" "592 8" "592 16"; do
  set -- $cfg
  echo "== batch32 handoff=$1 tiles=$2" >> $out
  FRI_DEC_HANDOFF=$1 FRI_DEC_HANDOFF_TILES=$2 python profiles/exp_b2b.py --frames 32 --reps 20 --sets 2 --tag "b32h$1t$2" >> $out 2>&1
done
for cfg in "0 8" "592 8" "740 8"; do
  set -- $cfg
  echo "== gray handoff=$1 tiles=$2" >> $out
  FRI_DEC_HANDOFF=$1 FRI_DEC_HANDOFF_TILES=$2 python profiles/exp_b2b.py --shape 4096x4096x1 --reps 400 --sets 8 --tag "g$1t$2" >> $out 2>&1
done
cat $out

#!/bin/bash
# 1-channel group-shape sweep with back-to-back launch timing (profiles/exp_b2b.py)
for cfg in ${GRAY_CFGS:-"8x4 4" "4x4 2" "8x2 2"}; do
  set -- $cfg
  FRI_GROUP=$1 FRI_TILES_PER_WARP=$2 python profiles/exp_b2b.py --shape 4096x4096x1 --divisor 1 --tag "g$1-t$2" --reps 200
  FRI_GROUP=$1 FRI_TILES_PER_WARP=$2 python profiles/exp_b2b.py --shape 512x512x1 --frames 256 --divisor 1 --tag "g$1-t$2" --reps 50
done

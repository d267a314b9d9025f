#!/bin/bash
# Tuning sweep over CTA group shapes (FRI_GROUP) and tiles per warp; prints encode/decode GB/s.
mkdir -p gpurun_out
for cfg in "${@}"; do
  IFS=, read grp tpw <<< "$cfg"
  for mode in "single" "batch"; do
    if [ $mode == single ]; then ARGS="--steps 200"; else ARGS="--steps 40 --shape 3840x2160x3 --frames 8"; fi
    FRI_GROUP=$grp FRI_TILES_PER_WARP=$tpw python bench.py $ARGS --warmup 5 --no-cpu --no-batched --preheat 0.3 > gpurun_out/sweep.log 2>&1
    python - "$cfg" $mode <<PY
import json, sys
try:
    d = json.loads(open("gpurun_out/sweep.log").read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], "enc %.0f GB/s" % d["roofline_encode"]["achieved"], "dec %.0f GB/s" % d["roofline_decode"]["achieved"],
          "threads", d["launch"]["threads"], "smem", d["launch"]["smem_bytes"], "groups", d["launch"]["n_groups"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e, open("gpurun_out/sweep.log").read()[-600:])
PY
  done
done

#!/bin/bash
# Runs the single-frame and batched bench for each library variant given as argument (tag or "base").
for tag in "$@"; do
  if [ $tag == base ]; then unset FRI_CUDA_LIB; else export FRI_CUDA_LIB=$PWD/frave_b200/libfri_cuda_$tag.so; fi
  for mode in single batch; do
    if [ $mode == single ]; then ARGS="--steps 200"; else ARGS="--steps 40 --shape 3840x2160x3 --frames 8"; fi
    python bench.py $ARGS --warmup 5 --no-cpu --no-batched --preheat 0.3 > gpurun_out/var.log 2>&1
    python - $tag $mode <<PY
import json, sys
try:
    d = json.loads(open("gpurun_out/var.log").read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], "enc %.0f GB/s" % d["roofline_encode"]["achieved"], "dec %.0f GB/s" % d["roofline_decode"]["achieved"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e, open("gpurun_out/var.log").read()[-600:])
PY
  done
done

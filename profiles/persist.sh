#!/bin/bash
FRI_PERSISTENT=1 timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for pm in 0 1; do
  export FRI_PERSISTENT=$pm
  for mode in single batch; do
    if [ $mode == single ]; then ARGS="--steps 200"; else ARGS="--steps 40 --shape 3840x2160x3 --frames 8"; fi
    python bench.py $ARGS --warmup 5 --no-cpu --preheat 0.3 > gpurun_out/var.log 2>&1
    python - $pm $mode <<PY
import json, sys
try:
    d = json.loads(open("gpurun_out/var.log").read().strip().splitlines()[-1])
    print("persistent", sys.argv[1], sys.argv[2], "enc %.0f GB/s %.1f us" % (d["roofline_encode"]["achieved"], 1e3*d["roofline_encode"]["avg_launch_ms"]), "dec %.0f GB/s %.1f us" % (d["roofline_decode"]["achieved"], 1e3*d["roofline_decode"]["avg_launch_ms"]), "value %.0f" % d["value"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e, open("gpurun_out/var.log").read()[-800:])
PY
  done
done

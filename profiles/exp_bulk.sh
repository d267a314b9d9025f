#!/bin/bash
# bulk-copy (TMA) vs cp.async staging of the encoder, by launch size
run() {
  python bench.py --warmup 5 --no-cpu --no-batched --no-e2e --preheat 0.3 $2 > gpurun_out/var.log 2>&1
  python - "$1" <<PY
import json, sys
try:
    d = json.loads(open("gpurun_out/var.log").read().strip().splitlines()[-1])
    print(sys.argv[1], "enc %.0f GB/s %.1f us" % (d["roofline_encode"]["achieved"], 1e3*d["roofline_encode"]["avg_launch_ms"]), "dec %.0f GB/s" % d["roofline_decode"]["achieved"], "i16 enc %.1f us" % (1e3*d["int16_arrays"]["encode"]["avg_launch_ms"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/var.log").read()[-600:])
PY
}
for rep in 1 2; do
for b in 0 1; do
  FRI_STAGE_BULK=$b run "bulk=$b 1x4096^2" "--steps 200"
  FRI_STAGE_BULK=$b run "bulk=$b 2x4K" "--steps 100 --shape 3840x2160x3 --frames 2"
  FRI_STAGE_BULK=$b run "bulk=$b 4x4K" "--steps 60 --shape 3840x2160x3 --frames 4"
  FRI_STAGE_BULK=$b run "bulk=$b 8x4K" "--steps 40 --shape 3840x2160x3 --frames 8"
  FRI_STAGE_BULK=$b run "bulk=$b 32x4K" "--steps 10 --shape 3840x2160x3 --frames 32"
done; done

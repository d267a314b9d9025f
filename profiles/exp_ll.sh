#!/bin/bash
# encode: first-wave CTAs issue the look-ahead prefetch after their own staging completes
# (FRI_ENC_LATE_LOOKAHEAD_CTAS; default = resident CTAs)
run() {
  python bench.py --warmup 5 --no-cpu --no-batched --no-e2e --preheat 0.3 $2 > gpurun_out/var.log 2>&1
  python - "$1" <<PY
import json, sys
d = json.loads(open("gpurun_out/var.log").read().strip().splitlines()[-1])
print(sys.argv[1], "enc %.0f GB/s %.1f us" % (d["roofline_encode"]["achieved"], 1e3*d["roofline_encode"]["avg_launch_ms"]), "i16 enc %.1f us" % (1e3*d["int16_arrays"]["encode"]["avg_launch_ms"]))
PY
}
for rep in 1 2; do for n in 0 default; do
  if [ $n == default ]; then unset FRI_ENC_LATE_LOOKAHEAD_CTAS; else export FRI_ENC_LATE_LOOKAHEAD_CTAS=$n; fi
  run "late_lookahead_first=$n 1x4096^2" "--steps 200"
  run "late_lookahead_first=$n 8x4K" "--steps 40 --shape 3840x2160x3 --frames 8"
done; done

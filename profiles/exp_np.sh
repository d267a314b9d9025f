#!/bin/bash
# decode: CTAs of the first wave without the own-run L2 prefetch (FRI_DEC_NO_PREFETCH_CTAS; default = resident CTAs)
run() {
  python bench.py --warmup 5 --no-cpu --no-batched --no-e2e --preheat 0.3 $2 > gpurun_out/var.log 2>&1
  python - "$1" <<PY
import json, sys
d = json.loads(open("gpurun_out/var.log").read().strip().splitlines()[-1])
print(sys.argv[1], "dec %.0f GB/s %.1f us" % (d["roofline_decode"]["achieved"], 1e3*d["roofline_decode"]["avg_launch_ms"]), "i16 dec %.1f us" % (1e3*d["int16_arrays"]["decode"]["avg_launch_ms"]))
PY
}
for rep in 1 2; do for n in 0 default; do
  if [ $n == default ]; then unset FRI_DEC_NO_PREFETCH_CTAS; else export FRI_DEC_NO_PREFETCH_CTAS=$n; fi
  run "no_prefetch_first=$n 1x4096^2" "--steps 200"
  run "no_prefetch_first=$n 8x4K" "--steps 40 --shape 3840x2160x3 --frames 8"
  run "no_prefetch_first=$n 16x1080p" "--steps 40 --shape 1920x1080x3 --frames 16"
done; done

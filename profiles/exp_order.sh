run() {
  python bench.py --warmup 5 --no-cpu --no-batched --no-e2e --preheat 0.3 $2 > gpurun_out/var.log 2>&1
  python - "$1" <<PY
import json, sys
d = json.loads(open("gpurun_out/var.log").read().strip().splitlines()[-1])
print(sys.argv[1], "enc %.0f GB/s %.1f us" % (d["roofline_encode"]["achieved"], 1e3*d["roofline_encode"]["avg_launch_ms"]), "dec %.0f GB/s %.1f us" % (d["roofline_decode"]["achieved"], 1e3*d["roofline_decode"]["avg_launch_ms"]))
PY
}
for rep in 1 2; do for o in 0 1 2; do
  FRI_ORDER=$o run "order=$o 1x4096^2" "--steps 200"
  FRI_ORDER=$o run "order=$o 8x4K" "--steps 40 --shape 3840x2160x3 --frames 8"
done; done

for d in 1 4 5; do
python bench.py --steps 200 --warmup 5 --no-cpu --no-batched --preheat 0.3 --divisor $d > gpurun_out/div_$d.log 2>&1
python - $d <<PY
import json,sys
d=json.loads(open("gpurun_out/div_%s.log"%sys.argv[1]).read().strip().splitlines()[-1])
print("divisor",sys.argv[1],"enc %.0f GB/s %.1f us"%(d["roofline_encode"]["achieved"],1e3*d["roofline_encode"]["avg_launch_ms"]),"dec %.0f GB/s %.1f us"%(d["roofline_decode"]["achieved"],1e3*d["roofline_decode"]["avg_launch_ms"]))
PY
done

for b in 1 2 4 6 8; do
FRI_BANDS=$b python bench.py --steps 20 --warmup 3 --no-cpu --no-batched --preheat 0.2 > gpurun_out/var.log 2>&1
python - $b <<PY
import json,sys
d=json.loads(open("gpurun_out/var.log").read().strip().splitlines()[-1])
print("bands",sys.argv[1],{k:round(v) for k,v in d["e2e"]["variants_mpix_s"].items()})
PY
done

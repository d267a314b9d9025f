#!/bin/bash
# One GPU-box pass: parity tests, the headline bench, a batched bench and a compact ncu launch list.
# usage: bash profiles/gpu_check.sh [tag]   (outputs under gpurun_out/)
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
python bench.py --steps 300 --warmup 5 --no-cpu --no-batched > gpurun_out/bench_$TAG.log 2>&1
python bench.py --steps 60 --warmup 5 --no-cpu --no-batched --shape 3840x2160x3 --frames 8 > gpurun_out/bench8_$TAG.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/bench_$TAG.log", "gpurun_out/bench8_$TAG.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "MPix/s %.0f" % d["value"], "enc %.0f GB/s %.1f us" % (d["roofline_encode"]["achieved"], 1e3 * d["roofline_encode"]["avg_launch_ms"]),
              "dec %.0f GB/s %.1f us" % (d["roofline_decode"]["achieved"], 1e3 * d["roofline_decode"]["avg_launch_ms"]), "e2e %.0f" % d["e2e"]["value"], d["clocks"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f).read()[-2000:])
PY
CMD="python bench.py --steps 3 --warmup 3 --preheat 0 --no-cpu --no-batched"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active --clock-control none -k regex:fri_ -s 6 -c 4 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
python profiles/ncu_launches.py gpurun_out/launches_$TAG.csv

#!/bin/bash
# Round-1 evidence pass (one GPU): parity suite, bench lines, ncu launch list + full captures,
# CTA timeline, the other configs, PCIe context.  Outputs under gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
T=r1f
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$T.log; tail -3 gpurun_out/pytest_$T.log
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_$T.err
python bench.py --steps 60 --warmup 5 --no-cpu --no-batched --shape 3840x2160x3 --frames 8 > gpurun_out/bench8_$T.json 2>> gpurun_out/bench_$T.err
python profiles/pcie.py > gpurun_out/pcie_$T.txt 2>&1
python profiles/extra_configs.py > gpurun_out/other_$T.jsonl 2>&1
python -m frave_b200.build --variant trace -DFRI_TRACE=1 > /dev/null 2>&1; python profiles/trace.py > gpurun_out/timeline_$T.txt 2>&1
CMD="python bench.py --steps 3 --warmup 3 --preheat 0 --no-cpu --no-batched --no-e2e"
$CMD > gpurun_out/plain_$T.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:fri_ -c 40 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_$T.log 2>&1
for K in encode decode; do
  ncu --set full --clock-control none --import-source on -k regex:fri_$K -s 3 -c 1 -o gpurun_out/prof_${T}_$K -f $CMD > gpurun_out/ncu_${T}_$K.log 2>&1
done
ls -la gpurun_out/*$T*

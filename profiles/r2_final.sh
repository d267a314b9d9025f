#!/bin/bash
# Round-2 evidence pass (one GPU): parity suite, bench lines (default, reference arm, batch256), the other
# BASELINE shapes with back-to-back timing, CTA timeline, ncu launch list + full captures.  Outputs under
# gpurun_out/ (copied to profiles/r2_* afterwards).
mkdir -p gpurun_out
T=r2f
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$T.log; tail -3 gpurun_out/pytest_$T.log
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench20_$T.json 2>> gpurun_out/bench_$T.err; echo "bench20 rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_$T.err
python bench.py --workload batch256 --no-cpu > gpurun_out/bench_batch256_$T.json 2>> gpurun_out/bench_$T.err; echo "batch256 rc=$?"
{
python profiles/exp_b2b.py --shape 4096x4096x3 --tag "configs[1] 4096x4096x3 u8"
python profiles/exp_b2b.py --shape 4096x4096x3 --half --tag "configs[1] int16 coefficient arrays"
python profiles/exp_b2b.py --shape 3840x2160x3 --frames 32 --reps 30 --tag "configs[2] per-GPU share at 8 GPUs: 32 x 4K"
python profiles/exp_b2b.py --shape 3840x2160x3 --frames 8 --reps 60 --tag "8 x 4K"
python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --divisor 1 --reps 20 --sets 2 --tag "configs[3] 16384^2 u16 depth 9"
python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --depth 16 --divisor 1 --reps 20 --sets 2 --tag "configs[3] depth 16"
python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --depth 20 --divisor 1 --reps 20 --sets 2 --tag "configs[3] depth 20"
python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --depth 24 --divisor 1 --reps 20 --sets 2 --tag "configs[3] depth 24"
python profiles/exp_b2b.py --shape 1920x1080x3 --frames 16 --reps 100 --tag "configs[4] 16 x 1080p"
python profiles/exp_b2b.py --shape 4096x4096x1 --divisor 1 --reps 200 --tag "4096^2 x 1 u8"
python profiles/exp_b2b.py --shape 512x512x1 --frames 256 --divisor 1 --reps 50 --tag "configs[0] shape batched: 256 x 512^2 x 1"
python profiles/exp_b2b.py --shape 512x512x1 --frames 1 --divisor 1 --reps 200 --tag "configs[0] single 512^2 x 1"
} > gpurun_out/other_$T.jsonl 2>&1
python profiles/emit_prof.py > gpurun_out/emit_$T.txt 2>&1
python profiles/codec_timing.py > gpurun_out/codec_$T.jsonl 2>&1
python profiles/pcie.py > gpurun_out/pcie_$T.txt 2>&1
python -m frave_b200.build --variant trace -DFRI_TRACE=1 > /dev/null 2>&1; python profiles/trace.py > gpurun_out/timeline_$T.txt 2>&1; python profiles/trace.py 4096x4096x1 >> gpurun_out/timeline_$T.txt 2>&1
CMD="python bench.py --steps 3 --warmup 3 --preheat 0 --no-cpu --no-batched --no-e2e"
$CMD > gpurun_out/plain_$T.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:fri_ -c 60 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_$T.log 2>&1
for K in encode decode; do
  ncu --set full --clock-control none --import-source on -k regex:fri_$K -s 3 -c 1 -o gpurun_out/prof_${T}_$K -f $CMD > gpurun_out/ncu_${T}_$K.log 2>&1
done
GCMD="python profiles/exp_b2b.py --shape 4096x4096x1 --divisor 1 --reps 5"
for K in encode decode; do
  ncu --set full --clock-control none --import-source on -k regex:fri_$K -s 3 -c 1 -o gpurun_out/prof_${T}_gray_$K -f $GCMD > gpurun_out/ncu_${T}_gray_$K.log 2>&1
done
ls -la gpurun_out/*$T*

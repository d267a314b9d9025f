python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
{
python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --divisor 1 --reps 20 --sets 2 --tag "configs[3] 16384^2 u16 depth 9"
python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --depth 16 --divisor 1 --reps 20 --sets 2 --tag "configs[3] depth 16"
python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --depth 20 --divisor 1 --reps 20 --sets 2 --tag "configs[3] depth 20"
python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --depth 24 --divisor 1 --reps 20 --sets 2 --tag "configs[3] depth 24"
} > gpurun_out/other_deep_final.jsonl 2>&1
cat gpurun_out/other_deep_final.jsonl | cut -c1-260
python profiles/codec_timing.py 2>/dev/null | tail -1 | cut -c1-400

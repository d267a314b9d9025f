for D in 16 24; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fri_ -s 6 -c 8 --csv --log-file gpurun_out/deep_$D.csv python profiles/exp_b2b.py --shape 16384x16384x1 --sample-bytes 2 --depth $D --divisor 1 --reps 3 --sets 2 > gpurun_out/deep_$D.log 2>&1
python profiles/ncu_launches.py gpurun_out/deep_$D.csv
done

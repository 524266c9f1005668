set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_r_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_r_tests.log
timeout 300 python tools/prof_filters.py --out gpurun_out/r2_r_filters.json > gpurun_out/r2_r_filters.txt 2>&1
RGIE_SCALE_COL=0 timeout 300 python tools/prof_filters.py > gpurun_out/r2_r_filters_nocol.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_r_filters_ncu.csv python tools/prof_filters.py --reps 1 > gpurun_out/r2_r_filters_ncu.log 2>&1
for v in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_r_bench_$v.json 2>> gpurun_out/r2_r_bench.err
done
echo done

set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s_tests.log
for v in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_s_bench_$v.json 2>> gpurun_out/r2_s_bench.err
done
echo done

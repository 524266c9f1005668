set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_l_tests.log
for mb in 32 64 32 64; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --micro-batch $mb > gpurun_out/r2_l_bench_mb${mb}_$RANDOM.json 2>> gpurun_out/r2_l_bench.err
done
echo done

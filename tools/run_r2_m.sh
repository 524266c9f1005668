set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_regressor_gpu.py -m gpu -q -x > gpurun_out/r2_m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_m_tests.log
for v in 0 8 16 0 8 4; do
  RGIE_STEM_SUB=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_m_bench_sub${v}_$RANDOM.json 2>> gpurun_out/r2_m_bench.err
done
echo done

set -x
mkdir -p gpurun_out
ok=1
for v in 1 2; do
  RGIE_LEAN_DMA=$v timeout 400 python -m pytest tests/test_regressor_gpu.py tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r2_g_dma${v}_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_g_dma${v}_test.log
  grep -q "rc=0" gpurun_out/r2_g_dma${v}_test.log || ok=0
done
if [ $ok = 1 ]; then
  for v in 0 1 2 0 1 2; do
    RGIE_LEAN_DMA=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_g_prof_dma$v.json > gpurun_out/r2_g_bench_dma${v}_$RANDOM.json 2>> gpurun_out/r2_g_bench.err
  done
fi
timeout 1200 python -m pytest tests/test_engine_gpu.py -m gpu -q -s > gpurun_out/r2_g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_g_tests.log
echo done

set -x
mkdir -p gpurun_out
RGIE_PATCH_2CTA=1 timeout 300 python -m pytest tests/test_regressor_gpu.py tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r2_f_p2_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_f_p2_test.log
if grep -q "rc=0" gpurun_out/r2_f_p2_test.log; then
  for v in 0 1 0 1; do
    RGIE_PATCH_2CTA=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_f_prof_p2$v.json > gpurun_out/r2_f_bench_p2${v}_$RANDOM.json 2>> gpurun_out/r2_f_bench.err
  done
  timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_f_tests.log
else
  RGIE_PATCH_2CTA=0 timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_f_tests.log
fi
echo done

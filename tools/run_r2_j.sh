set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -s -k "fp32_tensor_core" > gpurun_out/r2_j_gemm32.log 2>&1; echo "rc=$?" >> gpurun_out/r2_j_gemm32.log
if grep -q "rc=0" gpurun_out/r2_j_gemm32.log; then
  timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_j_tests.log
  timeout 600 python bench.py --precision fp32 --batch 32 --micro-batch 32 --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_j_prof_fp32tc.json > gpurun_out/r2_j_bench_fp32tc.json 2> gpurun_out/r2_j_bench_fp32tc.err
  RGIE_FP32_SIMT=1 timeout 900 python bench.py --precision fp32 --batch 32 --micro-batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_j_bench_fp32simt.json 2> gpurun_out/r2_j_bench_fp32simt.err
  timeout 300 python tools/prof_midu.py > gpurun_out/r2_j_midu.log 2>&1
else
  RGIE_FP32_SIMT=1 timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_j_tests_simt.log 2>&1; echo "rc=$?" >> gpurun_out/r2_j_tests_simt.log
fi
timeout 300 python tools/prof_filters.py --out gpurun_out/r2_j_filters.json > gpurun_out/r2_j_filters.log 2>&1
echo done

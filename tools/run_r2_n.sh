set -x
mkdir -p gpurun_out
RGIE_GEMM_B2B=1 timeout 300 python -m pytest tests/test_regressor_gpu.py -m gpu -q -x -k "bf16_tcgen05 or golden" > gpurun_out/r2_n_b2b_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_n_b2b_test.log
if grep -q "rc=0" gpurun_out/r2_n_b2b_test.log; then
  for v in 0 1 0 1; do
    RGIE_GEMM_B2B=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_n_prof_b2b$v.json > gpurun_out/r2_n_bench_b2b${v}_$RANDOM.json 2>> gpurun_out/r2_n_bench.err
  done
fi
echo done

set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_e_tests.log
for v in 0 1 0 1; do
  RGIE_GEMM_B2B=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_e_prof_b2b$v.json > gpurun_out/r2_e_bench_b2b${v}_$RANDOM.json 2>> gpurun_out/r2_e_bench.err
done
timeout 600 python bench.py --images 192 --sweep-steps 5 > gpurun_out/r2_e_sweep192.json 2> gpurun_out/r2_e_sweep192.err
echo done

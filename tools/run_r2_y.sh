set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_regressor_gpu.py -m gpu -q -x > gpurun_out/r2_y_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_y_tests.log
grep -q "rc=0" gpurun_out/r2_y_tests.log || exit 0
for v in 1 0 1 0; do
RGIE_CONV3_HSHARE=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_y_prof_h3$v.json > gpurun_out/r2_y_bench_h3${v}_$RANDOM.json 2>> gpurun_out/r2_y_bench.err
done
echo done

set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_filters_gpu.py tests/test_engine_gpu.py tests/test_resize_update_gpu.py tests/test_callers_gpu.py -m gpu -q -x > gpurun_out/r2_q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_q_tests.log
timeout 300 python tools/prof_filters.py --out gpurun_out/r2_q_filters.json > gpurun_out/r2_q_filters.txt 2>&1
RGIE_SHARP_MARCH=0 RGIE_SCALE_TAB=0 timeout 300 python tools/prof_filters.py > gpurun_out/r2_q_filters_old.txt 2>&1
RGIE_RESIZE_ROWS=16 timeout 300 python tools/prof_filters.py > gpurun_out/r2_q_filters_rows16.txt 2>&1
for v in 1 0 1 0; do
RGIE_FUSED_PREFIX=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_q_bench_prefix${v}_$RANDOM.json 2>> gpurun_out/r2_q_bench.err
done
RGIE_RESIZE_ROWS=16 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_q_bench_rows16.json 2>> gpurun_out/r2_q_bench.err
timeout 600 python -m pytest tests/test_imaginaire_gpu.py -m gpu -q -x -s > gpurun_out/r2_q_tests_imag.log 2>&1; echo "rc=$?" >> gpurun_out/r2_q_tests_imag.log
timeout 600 python bench.py --latent --steps 20 --warmup 3 > gpurun_out/r2_q_bench_latent.json 2> gpurun_out/r2_q_bench_latent.err
echo done

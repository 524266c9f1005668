set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_imaginaire_gpu.py tests/test_resize_update_gpu.py tests/test_engine_gpu.py -m gpu -q -s > gpurun_out/r2_t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t_tests.log
timeout 600 python bench.py --latent --steps 20 --warmup 3 > gpurun_out/r2_t_bench_latent.json 2> gpurun_out/r2_t_bench_latent.err
RGIE_LATENT_GRAPHS=0 timeout 600 python bench.py --latent --steps 20 --warmup 3 > gpurun_out/r2_t_bench_latent_eager.json 2> gpurun_out/r2_t_bench_latent_eager.err
timeout 300 python tools/prof_filters.py --out gpurun_out/r2_t_filters.json > gpurun_out/r2_t_filters.txt 2>&1
for v in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t_bench_$v.json 2>> gpurun_out/r2_t_bench.err
RGIE_RESIZE_THREADS=256 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t_bench_rs256_$v.json 2>> gpurun_out/r2_t_bench.err
done
echo done

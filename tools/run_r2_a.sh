set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
(python -c "import kornia; print(kornia.__version__)" 2>&1 | tail -1) > gpurun_out/r2_kornia.txt
tools/microbench/tma_overlap > gpurun_out/r2_tma_overlap.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_tests_a.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_a.log
RGIE_ZZ16=1 timeout 600 python -m pytest tests/test_regressor_gpu.py tests/test_emonet_gpu.py -m gpu -q -x > gpurun_out/r2_tests_zz16.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_zz16.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/r2_a_prof.json > gpurun_out/r2_a_bench.json 2> gpurun_out/r2_a_bench.err
RGIE_ZZ16=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_a_prof_zz16.json > gpurun_out/r2_a_bench_zz16.json 2> gpurun_out/r2_a_bench_zz16.err
timeout 300 python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_a_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2200 --csv --log-file gpurun_out/r2_a_launches.csv python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_a_ncu.log 2>&1
echo done

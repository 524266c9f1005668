set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_zz_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_zz_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/r2_zz_prof.json > gpurun_out/r2_zz_bench.json 2> gpurun_out/r2_zz_bench.err
timeout 900 python bench.py > gpurun_out/r2_zz_bench_default.json 2> gpurun_out/r2_zz_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_zz_bench_reference.json 2> gpurun_out/r2_zz_bench_reference.err
timeout 600 python bench.py --latent --steps 20 --warmup 3 > gpurun_out/r2_zz_bench_latent.json 2> gpurun_out/r2_zz_bench_latent.err
timeout 300 python tools/prof_filters.py --out gpurun_out/r2_zz_filters.json > gpurun_out/r2_zz_filters.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_zz_smoke.log 2>&1
timeout 300 python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_zz_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 2200 --csv --log-file gpurun_out/r2_zz_launches.csv python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_zz_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_conv1_pool_kernel|conv3_hshare_kernel" -c 2 -o gpurun_out/r2_zz_full_n64b python bench.py --batch 32 --micro-batch 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_zz_ncu_full.log 2>&1
echo done

# One gpurun call that regenerates the evidence of a build: GPU tests, bench lines (20 steps with the per-launch table, default
# 100 steps, reference arm, configs[2]), filter table, smoke, ncu launch list and one --set full capture.  TAG names the files.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG:-r2_zz}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG:-r2_zz}_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/${TAG:-r2_zz}_prof.json > gpurun_out/${TAG:-r2_zz}_bench.json 2> gpurun_out/${TAG:-r2_zz}_bench.err
timeout 900 python bench.py > gpurun_out/${TAG:-r2_zz}_bench_default.json 2> gpurun_out/${TAG:-r2_zz}_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG:-r2_zz}_bench_reference.json 2> gpurun_out/${TAG:-r2_zz}_bench_reference.err
timeout 600 python bench.py --latent --steps 20 --warmup 3 > gpurun_out/${TAG:-r2_zz}_bench_latent.json 2> gpurun_out/${TAG:-r2_zz}_bench_latent.err
timeout 300 python tools/prof_filters.py --out gpurun_out/${TAG:-r2_zz}_filters.json > gpurun_out/${TAG:-r2_zz}_filters.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG:-r2_zz}_smoke.log 2>&1
timeout 300 python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG:-r2_zz}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 2200 --csv --log-file gpurun_out/${TAG:-r2_zz}_launches.csv python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG:-r2_zz}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_conv1_pool_kernel|conv3_hshare_kernel" -c 2 -o gpurun_out/${TAG:-r2_zz}_full_n64b python bench.py --batch 32 --micro-batch 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG:-r2_zz}_ncu_full.log 2>&1
echo done

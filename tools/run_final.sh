# The short evidence call of a final build: GPU tests, one 20-step bench line with the per-launch table, smoke, and the ncu
# launch list (time, DRAM bytes, tensor-pipe activity per launch) that tools/summarize_launches.py turns into
# profiles/<TAG>_step_B32.json (stamped with the source hash bench.py matches).  TAG names the files.
set -x
T=${TAG:-r2_final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_gpu_tests.log
timeout 200 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/${T}_per_gemm.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 120 python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 2200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_gpu_tests.log; tail -1 gpurun_out/${T}_smoke.log
echo done

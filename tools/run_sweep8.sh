set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_s8_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 8 --images 1024 > gpurun_out/r2_s8_sweep1024.json 2> gpurun_out/r2_s8_sweep1024.err
timeout 900 $TR bench.py --gpus 8 --images 4096 > gpurun_out/r2_s8_sweep4096.json 2> gpurun_out/r2_s8_sweep4096.err
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_s8_bench_n8.json 2> gpurun_out/r2_s8_bench_n8.err
echo done

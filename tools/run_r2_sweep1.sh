set -x
mkdir -p gpurun_out
timeout 900 python bench.py --images 1024 > gpurun_out/r2_s1_sweep1024.json 2> gpurun_out/r2_s1_sweep1024.err
echo done

set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_imaginaire_gpu.py -m gpu -q -s -k "configs2 and fp32" > gpurun_out/r2_u_graph1.log 2>&1
RGIE_LATENT_GRAPHS=0 timeout 600 python -m pytest tests/test_imaginaire_gpu.py -m gpu -q -s -k "configs2 and fp32" > gpurun_out/r2_u_graph0.log 2>&1
echo done

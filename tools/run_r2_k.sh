set -x
mkdir -p gpurun_out
CMD="python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r2_k_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_patch_2cta_kernel -s 14 -c 2 -o gpurun_out/r2_k_patch2 $CMD > gpurun_out/r2_k_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_sm100_kernel<128" -s 40 -c 3 -o gpurun_out/r2_k_n128 $CMD > gpurun_out/r2_k_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_sm100_kernel<256, 3, 9" -s 42 -c 4 -o gpurun_out/r2_k_dma $CMD > gpurun_out/r2_k_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done

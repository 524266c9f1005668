set -x
mkdir -p gpurun_out
CMD="python bench.py --batch 32 --micro-batch 32 --steps 2 --warmup 3 --no-cpu-baseline"
RGIE_GEMM_B2B=1 timeout 300 $CMD > gpurun_out/r2_d_plain.log 2>&1 || exit 1
export RGIE_GEMM_B2B=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_patch_kernel -s 14 -c 3 -o gpurun_out/r2_d_patch $CMD > gpurun_out/r2_d_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_b2b_kernel -s 8 -c 4 -o gpurun_out/r2_d_b2b $CMD > gpurun_out/r2_d_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_sm100_kernel<256, 3, 8" -s 28 -c 4 -o gpurun_out/r2_d_lean $CMD > gpurun_out/r2_d_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"maxpool|pack_crops|crop_grad" -s 8 -c 4 -o gpurun_out/r2_d_pool $CMD > gpurun_out/r2_d_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done

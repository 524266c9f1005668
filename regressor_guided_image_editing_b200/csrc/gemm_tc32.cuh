// fp32-accurate row-shifted GEMM on tcgen05 (gemm_tc32.cu): bf16x3 operand split, six partial products per fp32 product.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace rgie {

struct GemmPlanTc32 {
  GemmDesc d;                       // A, A2, res, mask, D: fp32;  Wt: three bf16 planes [3][n_pad][K] (split_weights_bf16x3)
  CUtensorMap tmA, tmA2, tmW;
  int num_m_tiles, num_n_tiles, w_rows, grid;
};

void split_weights_bf16x3(const float* w, size_t n, __nv_bfloat16* planes);
int build_gemm_tc32(const GemmDesc& d, GemmPlanTc32* p);
int run_gemm_tc32(const GemmPlanTc32& p, cudaStream_t st);

}  // namespace rgie

// Pre-built launch plan for the tcgen05 row-shifted GEMM (tensor maps are encoded once, reused every step).
#pragma once
#include "common.cuh"

namespace rgie {

struct GemmPlanSm100 {
  GemmDesc d;
  CUtensorMap tmA, tmA2, tmB;      // A, optional second operand, weights
  CUtensorMap tmD, tmR;            // output / residual boxes of the TMA-store epilogue (epi >= 1 / >= 2)
  int epi;                          // epilogue variant (gemm_sm100.cu: SmemLayout)
  int bn;
  int num_m_tiles, num_n_tiles;
  int grid;
};

int build_gemm_sm100(const GemmDesc& d, GemmPlanSm100* p);
int run_gemm_sm100(const GemmPlanSm100& p, cudaStream_t st);
int gemm_sm100_num_sms();

}  // namespace rgie

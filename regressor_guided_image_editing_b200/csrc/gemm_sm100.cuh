// Pre-built launch plan for the tcgen05 row-shifted GEMM (tensor maps are encoded once, reused every step).
#pragma once
#include "common.cuh"

namespace rgie {

struct GemmPlanSm100 {
  GemmDesc d;
  CUtensorMap tmA, tmA2, tmB;      // A as 128-row tiles / 136-row slabs, weights
  int n_groups;                     // runs of consecutive row offsets (one A slab load serves the whole run)
  int group_w[kMaxTaps], group_tap0[kMaxTaps];
  long group_off[kMaxTaps];
  int use_base_offset;
  int bn;
  int num_m_tiles, num_n_tiles;
  int grid;
};

int build_gemm_sm100(const GemmDesc& d, GemmPlanSm100* p);
int run_gemm_sm100(const GemmPlanSm100& p, cudaStream_t st);
int gemm_sm100_num_sms();

}  // namespace rgie

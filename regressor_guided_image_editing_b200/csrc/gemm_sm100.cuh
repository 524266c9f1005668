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
  // patch-tile variant (gemm_sm100.cu: gemm_patch_kernel): taps on a (dy, dx) grid, 16 x 8 pixel output tiles
  int patch, patch_ny, patch_nx, patch_dy0, patch_dx0, patch_wt, patch_tap[4][4];
  int patch_2cta;                   // 1: gemm_patch_2cta_kernel (cta_group::2, a CTA pair per pair of horizontally adjacent patches)
  // special == 1: conv_hshare_kernel (conv1 input gradient with the horizontal taps as the N dimension)
  int special, hs_wt, hs_dy0, hs_dx0, hs_col0;
  long hs_lines;
  int num_m_tiles, num_n_tiles;
  int grid;
  // b2b == 1: gemm_b2b_kernel -- `d` (256-wide 1x1 stage) and the next layer's 1x1 reduction `d2` (256 -> 64) in one launch
  int b2b;
  GemmDesc d2;
  CUtensorMap tmW2;
};

int build_gemm_sm100(const GemmDesc& d, GemmPlanSm100* p);
int build_conv_hshare_sm100(const GemmDesc& d, const void* Wh, int dy0, int dx0, GemmPlanSm100* p);
// back-to-back fusion of two consecutive 1x1 ops (the second one reads exactly what the first one writes)
bool gemm_b2b_eligible(const GemmDesc& d1, const GemmDesc& d2);
int build_gemm_b2b_sm100(const GemmDesc& d1, const GemmDesc& d2, GemmPlanSm100* p);
int run_gemm_sm100(const GemmPlanSm100& p, cudaStream_t st);
int make_tensor_map_2d(CUtensorMap* map, const void* base, int dtype, uint64_t inner, uint64_t rows, uint32_t box_inner,
                       uint32_t box_rows, int swizzle_bytes, uint64_t row_elems);
int gemm_sm100_num_sms();

}  // namespace rgie

// Pre-built launch plan for the tcgen05 row-shifted GEMM (tensor maps are encoded once, reused every step).
#pragma once
#include "common.cuh"

namespace rgie {

// conv1 + 3x3/2 max-pool in one launch (gemm_conv1_pool_kernel): a tile is the 16 x 8 conv patch that holds 7 x 3 complete
// pooling windows; tiles of one image overlap by 2 lines / 2 columns so that no window straddles two tiles.
struct StemPoolParams {
  int n_img, H0, Hp;            // conv output H0 x H0 per image, pooled Hp x Hp
  int lines_per_img;            // lines of one image in the conv's [lines, P] source grid
  int TY, TX;                   // tiles per image
  FastDiv fd_img, fd_tx;        // dividers by TY * TX and TX
  void* P1; Geom g1;            // pooled output rows (padded layout of the next stage), 64 channels, bf16
  uint8_t* arg;                 // [n, Hp, Hp, 64]: argmax code 0..8 | 0x10 when the maximum is > 0
};

struct GemmPlanSm100 {
  GemmDesc d;
  CUtensorMap tmA, tmA2, tmB;      // A, optional second operand, weights
  CUtensorMap tmD, tmR;            // output / residual boxes of the TMA-store epilogue (epi >= 1 / >= 2)
  int epi;                          // epilogue variant (gemm_sm100.cu: SmemLayout)
  int bn;
  // patch-tile variant (gemm_sm100.cu: gemm_patch_kernel): taps on a (dy, dx) grid, 16 x 8 pixel output tiles
  int patch, patch_ny, patch_nx, patch_dy0, patch_dx0, patch_wt, patch_tap[4][4];
  int patch_2cta;                   // 1: gemm_patch_2cta_kernel (cta_group::2, a CTA pair per pair of horizontally adjacent patches)
  // special == 1: conv_hshare_kernel (conv1 input gradient with the horizontal taps as the N dimension); 2: conv3_hshare_kernel
  int special, hs_wt, hs_dy0, hs_dx0, hs_col0;
  long hs_lines;
  int num_m_tiles, num_n_tiles;
  int grid;
  // b2b == 1: gemm_b2b_kernel -- `d` (256-wide 1x1 stage) and the next layer's 1x1 reduction `d2` (256 -> 64) in one launch
  int b2b;
  GemmDesc d2;
  CUtensorMap tmW2;
  // pool == 1: gemm_conv1_pool_kernel (`d` = conv1 forward over the packed 16-channel crops; its output is never written)
  int pool;
  StemPoolParams sp;
};

int build_gemm_sm100(const GemmDesc& d, GemmPlanSm100* p);
int build_conv_hshare_sm100(const GemmDesc& d, const void* Wh, int dy0, int dx0, GemmPlanSm100* p);
// 3x3 64 -> 64 with the horizontal taps as the N dimension (conv3_hshare_kernel); wh_buf: 192 x 192 bf16 device scratch of the caller
int build_conv3_hshare_sm100(const GemmDesc& d, void* wh_buf, GemmPlanSm100* p);
// back-to-back fusion of two consecutive 1x1 ops (the second one reads exactly what the first one writes)
bool gemm_b2b_eligible(const GemmDesc& d1, const GemmDesc& d2);
int build_gemm_b2b_sm100(const GemmDesc& d1, const GemmDesc& d2, GemmPlanSm100* p);
// conv1 forward (4 vertical taps over the 16-channel packed crops) fused with the 3x3 stride-2 max-pool that follows it
int build_conv1_pool_sm100(const GemmDesc& conv1, void* P1, const Geom& g1, uint8_t* arg, int Hp, GemmPlanSm100* p);
int run_gemm_sm100(const GemmPlanSm100& p, cudaStream_t st);
int make_tensor_map_2d(CUtensorMap* map, const void* base, int dtype, uint64_t inner, uint64_t rows, uint32_t box_inner,
                       uint32_t box_rows, int swizzle_bytes, uint64_t row_elems);
int gemm_sm100_num_sms();

}  // namespace rgie

// tcgen05 / TMEM / TMA backend of the row-shifted GEMM (common.cuh: GemmDesc) -- sm_100a only.
//
//   D[map(m), n] = epi( sum_t sum_c A[m + row_off[t], c] * Wt[n, t*Cin + c] + sum_c A2[m, c] * Wt[n, ntaps*Cin + c] )
//                                                                                     bf16 x bf16 -> fp32 (TMEM)
//
// One persistent CTA per SM (gemm_sm100_kernel), 320 threads:
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d, 128B swizzle, zero fill for out-of-range rows = conv padding)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (cta_group::1, M=128, N=BN, K=16 per instruction)
//   warps 2..9  : epilogue       (tcgen05.ld 32x32b -> bias / residual / ReLU / ReLU-mask -> row-remapped 32B stores
//                                 + 1 sign bit per stored element).  The residual operand is prefetched a whole tile ahead
//                                 in registers (its HBM latency is never exposed); ReLU masks are 1 bit per element.
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) between TMA and MMA; 2 TMEM accumulator buffers
// (tmem_full/tmem_empty) between MMA and epilogue so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Variants in this file (selection: build_gemm_sm100; measurements: profiles/README.md):
//   * 640 threads, 16 epilogue warps (epilogue_lean_role, setmaxnreg register hand-over) for 256-wide bf16 tiles; output by
//     32-byte global stores or, where destination rows = source rows, through shared-memory boxes + one TMA store per part;
//   * gemm_sm100_2cta_kernel: CTA pairs (cta_group::2, M = 256 across a TPC) for contractions >= 768;
//   * lean role with DMA threads (EPI 9): output AND residual move by TMA through in-place half boxes;
//   * gemm_b2b_kernel (opt-in): a 256-wide 1x1 stage and the next layer's 256 -> 64 reduction in one launch;
//   * gemm_patch_kernel (16 x 8 pixel patch tiles, resident weights) for the 64-channel multi-tap convolutions and
//     conv_hshare_kernel (horizontal taps as N) for the conv1 input gradient.
#include <stdlib.h>
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "gemm_sm100.cuh"

namespace rgie {

namespace {

// EPI selects the epilogue: 0 = row-per-thread global accesses (any destination mapping, fp32 or bf16 output; 8 or 16 warps);
// 8 = lean 16-warp role whose output leaves through four 128 x 64 SWIZZLE_128B boxes and TMA stores; 9 = 8 + the residual
// operand arrives by TMA as well (in-place half boxes of 128 x 32, SWIZZLE_64B, served by two DMA threads).
constexpr int EPI_BOX_BYTES = BM * 32 * 2;
template <int BN, int STAGES, int EPI>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int A_BYTES = A_STAGE_BYTES;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_BYTES;
  static constexpr int OB_OFF = B_OFF + STAGES * B_STAGE_BYTES;                 // output boxes (EPI 8: 4 x 16 KB, EPI 9: 8 x 8 KB)
  static constexpr int BAR_OFF = OB_OFF + (EPI >= 8 ? 4 * BM * 128 : 0);       // full[S], empty[S], tfull[2], tempty[2], rfull[8], ready[8]
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * STAGES + 20) * 8;
  static constexpr int BIAS_OFF = TMEM_PTR_OFF + 16;                // whole bias vector (<= MAX_BIAS floats)
  static constexpr int TOTAL = BIAS_OFF + MAX_BIAS * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;   // + slack for manual 1024 B alignment
  static_assert(DYN_BYTES <= 232448, "shared memory plan exceeds 227 KB");
};

// Tile -> pixel-row mapping of the epilogue.  Flat tiles (enabled == 0): tile mt covers rows m_begin + mt*128 + r.
// Patch tiles (gemm_patch_kernel): tile mt = (ht, wt) covers the 16 x 8 pixel patch at line ht*16, column wt*8 of the
// [lines, P] pixel grid; accumulator row r is pixel (r >> 3, r & 7) of the patch.
struct PatchMap {
  int enabled, P, WT;
  FastDiv fd_wt;
};

// ===================== epilogue role (8 warps; one TMEM lane = one output row per thread) =====================
// warps 2..9: lane quarter q = warp & 3 (hardware restriction on tcgen05.ld), column half = (warp - 2) >> 2.
// A thread walks the 32-column chunks of its row / column half, tile after tile.  Operands it has to read:
//   residual (bf16 [m, n])   two register buffers; a buffer is refilled with the chunk two positions ahead (crossing
//                            into the next tile) right after it has been consumed, so the loads never wait for a tile
//                            boundary
//   accumulator (TMEM)       tcgen05.ld of chunk c+1 is in flight while chunk c is processed
//   ReLU mask as bits        one 32-bit word per chunk, loaded one tile ahead (blocked layout: bits_index)
//   ReLU mask as activations loaded at use (MiDU head only; the regressor uses bits)
// All global accesses of activations are 256-bit (one full 32 B sector per thread per instruction).
// TSPLIT = false: the NEW/4 warp groups split the COLUMNS of every tile (2 accumulator buffers).
// TSPLIT = true : the warp groups take alternate TILES (all columns; NACC accumulator buffers, one arrive group per
//                 buffer): with narrow N tiles the epilogue of a tile is a latency chain, not a throughput problem, so
//                 two tiles in flight halve its cost.
// TWO = true    : CTA-pair patch kernel (cta_group::2): the tile sequence of a CTA is unchanged (CTA b walks tiles b, b + grid,
//                 ...; CTAs 2c and 2c + 1 walk M-adjacent tiles in lockstep), only the accumulator-free arrive goes to the
//                 LEADER CTA's barrier (remotely for the peer).
template <int BN, int NEW, bool TSPLIT = false, int NACC = 2, bool TWO = false>
__device__ __forceinline__ void epilogue_role(const GemmDesc& d, const float* sbias, const uint32_t tmem_base,
                                              const uint32_t tfull0, const uint32_t tempty0, const int warp, const int lane,
                                              const int num_tiles, const int num_n_tiles, const FastDiv fd_nt,
                                              const PatchMap pm) {
  constexpr int NSPLIT = NEW / 4;                  // column parts: 4 warps (one per TMEM lane quarter) share a part
  constexpr int CH = BN >= 32 ? 32 : 16;
  auto tfull_bar = [&](int a) { return tfull0 + 8u * a; };
  auto tempty_bar = [&](int a) { return tempty0 + 8u * a; };
  const int q = warp & 3;
  const int part = (warp - 2) >> 2;
  const int row = q * 32 + lane;
  constexpr int NCH = BN / CH;                     // chunks per tile
  constexpr int CSPLIT = TSPLIT ? 1 : NSPLIT;      // column parts
  constexpr int TSTEP = TSPLIT ? NSPLIT : 1;       // tile stride between the tiles of one warp group
  constexpr int CPW = NCH >= CSPLIT ? NCH / CSPLIT : 1;      // chunks per warp
  const int c_begin = TSPLIT ? 0 : part * CPW;
  const bool active = c_begin < NCH;
  const int tile0 = blockIdx.x + (TSPLIT ? part * (int)gridDim.x : 0);
  const int tile_step = TSTEP * (int)gridDim.x;
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(d.res);
  const __nv_bfloat16* mask = reinterpret_cast<const __nv_bfloat16*>(d.mask);
  const uint32_t* mbits = d.mask_bits;
  const bool has_bias = d.bias != nullptr;
  const long res_lim = d.res_rows < d.m_end ? d.res_rows : d.m_end;
  // pixel row of this thread in m-tile mt (d.m_end = not a row)
  auto row_m = [&](int mt) -> long {
    if (!pm.enabled) return d.m_begin + (long)mt * BM + row;
    const int ht = (int)pm.fd_wt.div((uint32_t)mt), wt = mt - ht * pm.WT;
    const int w = wt * 8 + (row & 7);
    return w < pm.P ? (long)(ht * 16 + (row >> 3)) * pm.P + w : d.m_end;
  };

  // operands of the NEXT tile (prefetched while the current one is processed)
  constexpr int NB = CPW >= 2 ? 2 : 1;             // residual buffers (chunk ci lives in buffer ci % NB; CPW is 1 or even)
  uint32_t rbuf[NB][CH / 2];                       // residual: CH bf16 per chunk
  uint32_t bits_nxt[CPW];
  const __nv_bfloat16* nres = nullptr;             // residual row of the next tile (null: nothing to read)
  const __nv_bfloat16* cres = nullptr;             // residual row of the current tile
  // seq-th tile of this warp group -> (mt, nt); false past the end
  auto decode = [&](int seq, int& mt, int& nt) -> bool {
    const int tile = tile0 + seq * tile_step;
    if (tile >= num_tiles) return false;
    mt = (int)fd_nt.div((uint32_t)tile);
    nt = tile - mt * num_n_tiles;
    return true;
  };
  auto locate = [&](int seq) {
    cres = nres;
    nres = nullptr;
#pragma unroll
    for (int i = 0; i < CPW; ++i) bits_nxt[i] = 0u;
    int mt, nt;
    if (decode(seq, mt, nt) && active) {
      const long m = row_m(mt);
      if (res != nullptr && m < res_lim) nres = res + m * d.ld_res + nt * BN + c_begin * CH;
      if (mbits != nullptr && m < d.m_end) {
        const int w0 = (nt * BN) / 32 + c_begin;
#pragma unroll
        for (int i = 0; i < CPW; ++i) bits_nxt[i] = __ldg(mbits + bits_index(m, w0 + i, d.ld_mb));
      }
    }
  };
  // fetch chunk position cj (>= CPW: chunk cj - CPW of the next tile) into its buffer
  auto res_fetch = [&](int cj) {
    const __nv_bfloat16* p = cj < CPW ? cres : nres;
    const int cc = cj < CPW ? cj : cj - CPW;
    if (p != nullptr) {
#pragma unroll
      for (int j = 0; j < CH / 16; ++j) ldg256(p + cc * CH + j * 16, rbuf[cj % NB] + 8 * j);
    }
  };
  locate(0);               // nres = first tile
  {
    const __nv_bfloat16* first = nres;
    cres = first;
#pragma unroll
    for (int cj = 0; cj < NB; ++cj) res_fetch(cj);
  }

  int it = TSPLIT ? part : 0;                      // index of the tile in this CTA's sequence
  int mt, nt;
  for (int seq = 0; decode(seq, mt, nt); ++seq, it += TSTEP) {
    const int acc = it % NACC;
    const uint32_t acc_phase = (it / NACC) & 1;
    const long m = row_m(mt);
    long dest = -1;
    if (m < d.m_end) dest = map_row(d.src, d.dst_kind, d.dst, m);
    const bool live = active && dest >= 0;
    const bool use_res = live && res != nullptr && m < d.res_rows;
    const bool use_mask = live && mask != nullptr;
    const __nv_bfloat16* mask_p = mask + m * d.ld_mask + nt * BN;
    uint32_t bits_cur[CPW], bits_out[CPW];
#pragma unroll
    for (int i = 0; i < CPW; ++i) { bits_cur[i] = bits_nxt[i]; bits_out[i] = 0u; }
    locate(seq + 1);                               // cres = this tile, nres / mask words = next tile
    mbar_wait(tfull_bar(acc), acc_phase);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
    if (active) {
      uint32_t racc[2][CH];
      tmem_ld<CH>(taddr + (uint32_t)(c_begin * CH), racc[0]);
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        const int c = c_begin + ci;
        tmem_ld_wait();
        if (ci + 1 < CPW) tmem_ld<CH>(taddr + (uint32_t)((c + 1) * CH), racc[(ci + 1) & 1]);
        const uint32_t* r = racc[ci & 1];
        const int n0 = nt * BN + c * CH;
        float v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
        if (use_res) {
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) { v[2 * j] += bf16_lo(rbuf[ci % NB][j]); v[2 * j + 1] += bf16_hi(rbuf[ci % NB][j]); }
        }
        res_fetch(ci + NB);                        // refill this buffer with the chunk NB positions ahead
        if (live) {
          if (has_bias) {
            const float4* b4 = reinterpret_cast<const float4*>(sbias + n0);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 b = b4[j];
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (d.relu) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
            if (d.D_bits != nullptr) {
              uint32_t w = 0u;
#pragma unroll
              for (int j = CH - 1; j >= 0; --j) w = push_positive_bit(w, v[j]);
              bits_out[ci] = w;
            }
          } else if (d.D_bits != nullptr) {
            uint32_t w = 0u;
#pragma unroll
            for (int j = 0; j < CH; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
            bits_out[ci] = w;
          }
          if (use_mask) {
#pragma unroll
            for (int jj = 0; jj < CH / 16; ++jj) {
              uint32_t k8[8];
              ldg256(mask_p + c * CH + jj * 16, k8);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                const uint32_t lo = k8[j] & 0xFFFFu, hi = k8[j] >> 16;
                if (!(lo != 0 && lo < 0x8000u)) { v[jj * 16 + 2 * j] = 0.f; bits_out[ci] &= ~(1u << (jj * 16 + 2 * j)); }
                if (!(hi != 0 && hi < 0x8000u)) { v[jj * 16 + 2 * j + 1] = 0.f; bits_out[ci] &= ~(1u << (jj * 16 + 2 * j + 1)); }
              }
            }
          }
          if (mbits != nullptr) {
            const uint32_t w = bits_cur[ci];
            bits_out[ci] &= w;
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = keep_if_bit(v[j], w, j);
          }
          if (d.d_fp32) {
            float* o = reinterpret_cast<float*>(d.D) + dest * d.ldd + n0;
#pragma unroll
            for (int j = 0; j < CH / 8; ++j) stg256(o + 8 * j, reinterpret_cast<const uint32_t*>(v) + 8 * j);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(d.D) + dest * d.ldd + n0;
            uint32_t pk[CH / 2];
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < CH / 16; ++j) stg256(o + 16 * j, pk + 8 * j);
          }
        }
      }
      if (live && d.D_bits != nullptr) {
        const int w0 = (nt * BN) / 32 + c_begin;
#pragma unroll
        for (int i = 0; i < CPW; ++i) d.D_bits[bits_index(dest, w0 + i, d.ld_db)] = bits_out[i];
      }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (TWO) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0)); else mbar_arrive(tempty_bar(acc));
    }
  }
}

// ===================== lean epilogue for 16 epilogue warps (BN = 256, bf16 output, no activation mask) =====================
// The HBM-bound 1x1 expansions are limited by the latency chain of their epilogue (tcgen05.ld -> math -> store) at two
// epilogue warps per scheduler.  This variant runs FOUR warps per scheduler: warps 4..19, each a TMEM lane quarter
// (warp & 3) x a 64-column part, walking its part in 16-column chunks so that the whole role fits in 120 registers
// (setmaxnreg moves registers from the producer warpgroup).  Same arithmetic and the same bit-mask layout as
// epilogue_role.
#ifndef RGIE_LEAN_NRB
#define RGIE_LEAN_NRB 4
#endif
// TWO: CTA-pair kernel -- `num_tiles` counts PAIRS of M-adjacent tiles, cluster c walks pairs c, c + clusters, ...; the CTA
// of rank r owns tile 2 * pair_m + r, and the accumulator-free arrive goes to the leader CTA's barrier.
// STORE (BN = 256, DST_SAME): the output does not leave through 32-byte row-per-thread global stores (32 distinct lines =
// 32 L1TEX wavefronts per warp instruction; ncu: LSU data pipe 73-82 % busy on the HBM-bound launches) but through a
// 128 x 64 SWIZZLE_128B box per column part in shared memory (two st.shared.v4 per chunk, conflict-free) and ONE TMA store
// per part and tile.  The four warps of a part meet at a named barrier before the store and before the box is reused.
template <int BN, bool TWO = false, bool STORE = false>
__device__ __forceinline__ void epilogue_lean_role(const GemmDesc& d, const float* sbias, const uint32_t tmem_base,
                                                   const uint32_t tfull0, const uint32_t tempty0, const int warp,
                                                   const int lane, const int num_tiles, const int num_n_tiles,
                                                   const FastDiv fd_nt, const CUtensorMap* tmD = nullptr,
                                                   const uint32_t ob_smem = 0) {
  static_assert(!STORE || (BN == 256 && !TWO), "the store variant is instantiated for single-CTA 256-wide tiles");
  const int t_first = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_stride = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int t_rank = TWO ? (int)cluster_ctarank() : 0;
  static_assert(BN == 256 || BN == 128, "lean epilogue: 4 parts of 64 (BN = 256) or 32 (BN = 128) columns");
  constexpr int CH = 16, PW = BN / 4, CPW = PW / CH, NW = PW / 32;   // columns, chunks and bit words per part
  auto tfull_bar = [&](int a) { return tfull0 + 8u * a; };
  auto tempty_bar = [&](int a) { return tempty0 + 8u * a; };
  const int q = warp & 3;
  const int part = (warp - 4) >> 2;
  const int row = q * 32 + lane;
  const int col0 = part * PW;
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(d.res);
  const uint32_t* mbits = d.mask_bits;
  const bool has_bias = d.bias != nullptr;
  const long res_lim = d.res_rows < d.m_end ? d.res_rows : d.m_end;

  constexpr int NRB = RGIE_LEAN_NRB < CPW ? RGIE_LEAN_NRB : CPW;   // residual: chunk ci lives in buffer ci % NRB, refilled NRB chunks ahead
  static_assert(NRB == 2 || NRB == 4, "NRB");
  uint32_t rbuf[NRB][CH / 2];
  uint32_t bits_nxt[2];
  const __nv_bfloat16* nres = nullptr;
  const __nv_bfloat16* cres = nullptr;
  auto locate = [&](int tile) {
    cres = nres;
    nres = nullptr;
    bits_nxt[0] = bits_nxt[1] = 0u;
    if (tile < num_tiles) {
      const int tq = (int)fd_nt.div((uint32_t)tile), nt = tile - tq * num_n_tiles;
      const int mt = TWO ? 2 * tq + t_rank : tq;
      const long m = d.m_begin + (long)mt * BM + row;
      if (res != nullptr && m < res_lim) nres = res + m * d.ld_res + nt * BN + col0;
      if (mbits != nullptr && m < d.m_end) {
        const int w0 = (nt * BN + col0) / 32;
        bits_nxt[0] = __ldg(mbits + bits_index(m, w0, d.ld_mb));
        if (NW == 2) bits_nxt[1] = __ldg(mbits + bits_index(m, w0 + 1, d.ld_mb));
      }
    }
  };
  auto res_fetch = [&](int cj) {
    const __nv_bfloat16* p = cj < CPW ? cres : nres;
    const int cc = cj < CPW ? cj : cj - CPW;
    if (p != nullptr) ldg256(p + cc * CH, rbuf[cj & (NRB - 1)]);
  };
  locate(t_first);
  cres = nres;
#pragma unroll
  for (int cj = 0; cj < NRB; ++cj) res_fetch(cj);

  // STORE: this thread's row inside its part's box, the box leader, the named barrier of the part (ids 1..4)
  const uint32_t box = ob_smem + (uint32_t)(part * (BM * 128));
  const uint32_t box_row = box + (uint32_t)row * 128u;
  const uint32_t swz = (uint32_t)(row & 7);
  const bool box_leader = STORE && q == 0 && lane == 0;
  const int bar_id = 1 + part;

  int it = 0;
  for (int tile = t_first; tile < num_tiles; tile += t_stride, ++it) {
    const int tq = (int)fd_nt.div((uint32_t)tile), nt = tile - tq * num_n_tiles;
    const int mt = TWO ? 2 * tq + t_rank : tq;
    const int acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1;
    const long m = d.m_begin + (long)mt * BM + row;
    long dest = -1;
    if (m < d.m_end) dest = map_row(d.src, d.dst_kind, d.dst, m);
    const bool live = dest >= 0;
    const bool use_res = live && res != nullptr && m < d.res_rows;
    const uint32_t bits_cur0 = bits_nxt[0], bits_cur1 = bits_nxt[1];
    uint32_t bits_out0 = 0u, bits_out1 = 0u;
    locate(tile + t_stride);
    mbar_wait(tfull_bar(acc), acc_phase);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + col0);
    if (STORE && it > 0) {
      if (box_leader) bulk_wait_read0();            // the previous tile's TMA store has finished reading the box
      named_bar_sync(bar_id, 128);
    }
#pragma unroll
    for (int ci = 0; ci < CPW; ++ci) {
      uint32_t r[CH];
      tmem_ld<CH>(taddr + (uint32_t)(ci * CH), r);
      tmem_ld_wait();
      const int n0 = nt * BN + col0 + ci * CH;
      float v[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
      if (use_res) {
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) {
          v[2 * j] += bf16_lo(rbuf[ci & (NRB - 1)][j]); v[2 * j + 1] += bf16_hi(rbuf[ci & (NRB - 1)][j]);
        }
      }
      res_fetch(ci + NRB);
      if (live) {
        if (has_bias) {
          const float4* b4 = reinterpret_cast<const float4*>(sbias + n0);
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) {
            const float4 b = b4[j];
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        uint32_t wout = 0u;
        if (d.relu) {
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
          if (d.D_bits != nullptr) {
#pragma unroll
            for (int j = CH - 1; j >= 0; --j) wout = push_positive_bit(wout, v[j]);
          }
        } else if (d.D_bits != nullptr) {
#pragma unroll
          for (int j = 0; j < CH; ++j) wout |= (v[j] > 0.f ? 1u : 0u) << j;
        }
        if (mbits != nullptr) {
          const uint32_t w = ((ci < 2 ? bits_cur0 : bits_cur1) >> (16 * (ci & 1))) & 0xFFFFu;
          wout &= w;
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = keep_if_bit(v[j], w, j);
        }
        if (ci < 2) bits_out0 |= wout << (16 * (ci & 1)); else bits_out1 |= wout << (16 * (ci & 1));
        uint32_t pk[CH / 2];
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        if (STORE) {
          sts128(box_row + (((uint32_t)(2 * ci) ^ swz) << 4), pk);
          sts128(box_row + (((uint32_t)(2 * ci + 1) ^ swz) << 4), pk + 4);
        } else {
          stg256(reinterpret_cast<__nv_bfloat16*>(d.D) + dest * d.ldd + n0, pk);
        }
      } else if (STORE) {
        const uint32_t z[4] = {0u, 0u, 0u, 0u};       // pad rows are zero by construction and stay zero
        sts128(box_row + (((uint32_t)(2 * ci) ^ swz) << 4), z);
        sts128(box_row + (((uint32_t)(2 * ci + 1) ^ swz) << 4), z);
      }
    }
    if (STORE) {
      fence_async_smem();
      named_bar_sync(bar_id, 128);
      if (box_leader) {
        tma_store_2d(tmD, box, nt * BN + col0, (int)(d.m_begin + (long)mt * BM));
        bulk_commit();
      }
    }
    if (live && d.D_bits != nullptr) {
      const int w0 = (nt * BN + col0) / 32;
      d.D_bits[bits_index(dest, w0, d.ld_db)] = bits_out0;
      if (NW == 2) d.D_bits[bits_index(dest, w0 + 1, d.ld_db)] = bits_out1;
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (TWO) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0)); else mbar_arrive(tempty_bar(acc));
    }
  }
  if (box_leader) bulk_wait0();
}

// ===================== lean epilogue with a DMA thread (EPI 9): BN = 256, DST_SAME, bf16 output =====================
// The 16-warp lean role above still fetches its residual operand with row-per-thread 32-byte loads: 32 distinct lines per
// warp instruction, the other half of the L1TEX/LSU load that the TMA-store output path removed (ncu on the fused kernel:
// epilogue warps stalled on the release of their address registers behind the LSU queue).  Here every byte of the epilogue
// moves by TMA.  A 64-column part is split into two half boxes of 128 rows x 32 columns (64-byte rows, SWIZZLE_64B, 8 KB); a
// half box holds, in turn, the RESIDUAL of a tile (TMA load), then -- written in place by the thread that owns the row --
// the OUTPUT of that tile (TMA store).  One otherwise idle thread of the producer warpgroup (warp 2) is the DMA engine of all
// eight half boxes: for each box in turn it waits for `ready` (the 4 warps of the part have written it), issues the TMA
// store, waits until the PREVIOUS store has finished reading its box and refills that box with the next tile's residual
// (`rfull`: the 4 warps wait for it before they touch the box again).  The residual of half box (p, h) is therefore loaded
// while the other half of the tile is being processed, and the epilogue warps execute no global memory instruction at all
// besides the 1-bit mask words; no named barriers either.
template <int BN>
__device__ __forceinline__ void epilogue_lean_dma_role(const GemmDesc& d, const float* sbias, const uint32_t tmem_base,
                                                       const uint32_t tfull0, const uint32_t tempty0, const uint32_t rfull0,
                                                       const uint32_t ready0, const int warp, const int lane,
                                                       const int num_tiles, const int num_n_tiles, const FastDiv fd_nt,
                                                       const uint32_t ob_smem) {
  static_assert(BN == 256, "EPI 9 is instantiated for 256-wide tiles");
  constexpr int CH = 16, PW = 64, CPW = 4;
  auto tfull_bar = [&](int a) { return tfull0 + 8u * a; };
  auto tempty_bar = [&](int a) { return tempty0 + 8u * a; };
  const int q = warp & 3;
  const int part = (warp - 4) >> 2;
  const int row = q * 32 + lane;
  const int col0 = part * PW;
  const bool has_res = d.res != nullptr;
  const uint32_t* mbits = d.mask_bits;
  const bool has_bias = d.bias != nullptr;
  // this thread's 64-byte row inside a half box: 16-byte piece j lives at piece (j ^ ((row >> 1) & 3))  (SWIZZLE_64B)
  const uint32_t row_off = (uint32_t)row * 64u;
  const uint32_t swz = (uint32_t)((row >> 1) & 3);
  auto box_addr = [&](int h) { return ob_smem + (uint32_t)((part * 2 + h) * EPI_BOX_BYTES); };
  auto rfull_bar = [&](int h) { return rfull0 + 8u * (part * 2 + h); };
  auto ready_bar = [&](int h) { return ready0 + 8u * (part * 2 + h); };

  uint32_t bits_nxt[2];
  auto bits_fetch = [&](int tile) {
    bits_nxt[0] = bits_nxt[1] = 0u;
    if (mbits != nullptr && tile < num_tiles) {
      const int mt = (int)fd_nt.div((uint32_t)tile), nt = tile - mt * num_n_tiles;
      const long m = d.m_begin + (long)mt * BM + row;
      if (m < d.m_end) {
        const int w0 = (nt * BN + col0) / 32;
        bits_nxt[0] = __ldg(mbits + bits_index(m, w0, d.ld_mb));
        bits_nxt[1] = __ldg(mbits + bits_index(m, w0 + 1, d.ld_mb));
      }
    }
  };
  bits_fetch(blockIdx.x);

  int it = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int mt = (int)fd_nt.div((uint32_t)tile), nt = tile - mt * num_n_tiles;
    const int acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1;
    const long m = d.m_begin + (long)mt * BM + row;
    long dest = -1;
    if (m < d.m_end) dest = map_row(d.src, d.dst_kind, d.dst, m);
    const bool live = dest >= 0;
    const bool use_res = live && has_res && m < d.res_rows;
    const uint32_t bits_cur[2] = {bits_nxt[0], bits_nxt[1]};
    uint32_t bits_out[2] = {0u, 0u};
    bits_fetch(tile + gridDim.x);
    mbar_wait(tfull_bar(acc), acc_phase);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + col0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // the box is ours again: the TMA store of the previous tile has read it and (if there is one) this tile's residual is in it
      mbar_wait(rfull_bar(h), (uint32_t)(it & 1));
      const uint32_t brow = box_addr(h) + row_off;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int ci = 2 * h + cc;
        uint32_t r[CH];
        tmem_ld<CH>(taddr + (uint32_t)(ci * CH), r);
        const uint32_t a0 = brow + (((uint32_t)(2 * cc) ^ swz) << 4), a1 = brow + (((uint32_t)(2 * cc + 1) ^ swz) << 4);
        uint32_t rs[8];
        if (use_res) { lds128(a0, rs); lds128(a1, rs + 4); }
        tmem_ld_wait();
        if (ci == CPW - 1) {
          // the accumulator is in registers: hand the TMEM buffer back before the arithmetic of the last chunk
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (live) {
          const int n0 = nt * BN + col0 + ci * CH;
          float v[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
          if (use_res) {
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) { v[2 * j] += bf16_lo(rs[j]); v[2 * j + 1] += bf16_hi(rs[j]); }
          }
          if (has_bias) {
            const float4* b4 = reinterpret_cast<const float4*>(sbias + n0);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 b = b4[j];
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          uint32_t wout = 0u;
          if (d.relu) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
            if (d.D_bits != nullptr) {
#pragma unroll
              for (int j = CH - 1; j >= 0; --j) wout = push_positive_bit(wout, v[j]);
            }
          } else if (d.D_bits != nullptr) {
#pragma unroll
            for (int j = 0; j < CH; ++j) wout |= (v[j] > 0.f ? 1u : 0u) << j;
          }
          if (mbits != nullptr) {
            const uint32_t w = (bits_cur[h] >> (16 * cc)) & 0xFFFFu;
            wout &= w;
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = keep_if_bit(v[j], w, j);
          }
          bits_out[h] |= wout << (16 * cc);
          uint32_t pk[CH / 2];
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          sts128(a0, pk);
          sts128(a1, pk + 4);
        } else {
          const uint32_t z[4] = {0u, 0u, 0u, 0u};       // pad rows are zero by construction and stay zero
          sts128(a0, z);
          sts128(a1, z);
        }
      }
      fence_async_smem();                 // generic-proxy writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) mbar_arrive(ready_bar(h));
    }
    if (live && d.D_bits != nullptr) {
      const int w0 = (nt * BN + col0) / 32;
      d.D_bits[bits_index(dest, w0, d.ld_db)] = bits_out[0];
      d.D_bits[bits_index(dest, w0 + 1, d.ld_db)] = bits_out[1];
    }
  }
}

// A DMA thread of EPI 9 (lane 0 of warp 2 serves parts 0-1, lane 0 of warp 3 parts 2-3).  Box k = part * 2 + half.  Per box:
// wait until the part has written it, store it, wait until the store has read it (a few hundred ns), refill it with the next
// tile's residual -- which then has the processing time of the OTHER half of the tile to arrive.
template <int BN>
__device__ __forceinline__ void epilogue_dma_thread(const GemmDesc& d, const CUtensorMap* tmD, const CUtensorMap* tmR,
                                                    const uint32_t rfull0, const uint32_t ready0, const int part0,
                                                    const int num_tiles, const int num_n_tiles, const FastDiv fd_nt,
                                                    const uint32_t ob_smem) {
  const bool has_res = d.res != nullptr;
  auto box_of = [&](int k) { return ob_smem + (uint32_t)(k * EPI_BOX_BYTES); };
  // hand box k to the epilogue warps for `tile`: with this tile's residual in it, or as it is when there is none
  auto refill = [&](int k, int tile) {
    if (tile >= num_tiles) return;
    const uint32_t bar = rfull0 + 8u * k;
    if (has_res) {
      const int mt = (int)fd_nt.div((uint32_t)tile), nt = tile - mt * num_n_tiles;
      mbar_expect_tx(bar, EPI_BOX_BYTES);
      tma_load_2d(box_of(k), tmR, nt * BN + (k >> 1) * 64 + (k & 1) * 32, (int)(d.m_begin + (long)mt * BM), bar);
    } else {
      mbar_arrive(bar);
    }
  };
#pragma unroll 1
  for (int s = 0; s < 4; ++s) refill((part0 + (s & 1)) * 2 + (s >> 1), blockIdx.x);
  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int mt = (int)fd_nt.div((uint32_t)tile), nt = tile - mt * num_n_tiles;
    const int m0 = (int)(d.m_begin + (long)mt * BM);
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      const int k = (part0 + (s & 1)) * 2 + (s >> 1);          // half 0 of both parts, then half 1 of both parts
      mbar_wait(ready0 + 8u * k, (uint32_t)(it & 1));
      tma_store_2d(tmD, box_of(k), nt * BN + (k >> 1) * 64 + (k & 1) * 32, m0);
      bulk_commit();
      bulk_wait_read0();                                        // the store has read the box
      refill(k, tile + (int)gridDim.x);
    }
  }
  bulk_wait0();
}

template <int BN, int STAGES, int EPI, int NEW>
__global__ void __launch_bounds__(num_threads(NEW), 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                  const __grid_constant__ CUtensorMap tmR, const GemmDesc d, const int num_m_tiles, const int num_n_tiles,
                  const FastDiv fd_nt) {
  using L = SmemLayout<BN, STAGES, EPI>;
  constexpr uint32_t TMEM_COLS = tmem_cols(BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto rfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 4 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int kb_per_tap = d.Cin / BK;
  const int num_kb = d.ntaps * kb_per_tap;                       // k-blocks of the row-shifted taps
  const int num_kb2 = d.A2 != nullptr ? d.Cin2 / BK : 0;         // k-blocks of the second operand (tiles below a2_rows)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NEW);   // one arrive per epilogue warp
    }
    for (int a = 0; a < 8; ++a) mbar_init(rfull_bar(a), 1);
    if (EPI == 9)
      for (int a = 0; a < 8; ++a) mbar_init(rfull_bar(8 + a), 4);     // ready[8]: the 4 warps of a part
    if (EPI >= 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmD) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmR) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (d.bias != nullptr) {
    float* sb = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    for (int i = threadIdx.x; i < d.Cout; i += num_threads(NEW)) sb[i] = d.bias[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // NEW == 16: register re-allocation between warpgroups -- 640 threads launch with 96 registers each; the producer
  // warpgroup (warps 0..3) keeps 32, which lets the 16 epilogue warps grow to 112 (the pool is the CTA's own allocation: 128 x 32 + 512 x 112 = 640 x 96).
  // Every warp of a warpgroup executes the SAME setmaxnreg instruction (.sync.aligned), and it dominates the role code
  // so that ptxas allocates each role against its own budget.
  if (NEW == 16 && warp >= 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    if constexpr (NEW == 16 && EPI == 9)
      epilogue_lean_dma_role<BN>(d, reinterpret_cast<const float*>(smem + L::BIAS_OFF), tmem_base, tfull_bar(0), tempty_bar(0),
                                 rfull_bar(0), rfull_bar(8), warp, lane, num_tiles, num_n_tiles, fd_nt, smem_base + L::OB_OFF);
    else if constexpr (NEW == 16)
      epilogue_lean_role<BN, false, EPI == 8>(d, reinterpret_cast<const float*>(smem + L::BIAS_OFF), tmem_base, tfull_bar(0),
                                              tempty_bar(0), warp, lane, num_tiles, num_n_tiles, fd_nt, &tmD,
                                              smem_base + L::OB_OFF);
  } else {
  if (NEW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = (int)fd_nt.div((uint32_t)tile);
        const int nt = tile - mt * num_n_tiles;
        const long m0 = d.m_begin + (long)mt * BM;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), L::A_BYTES + L::B_STAGE_BYTES);
          tma_load_2d(smem_base + L::A_OFF + stage * L::A_BYTES, &tmA, cb * BK, (int)(m0 + d.row_off[tap]), full_bar(stage));
          tma_load_2d(smem_base + L::B_OFF + stage * L::B_STAGE_BYTES, &tmB, kb * BK, nt * BN, full_bar(stage));
          if (++cb == kb_per_tap) { cb = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (m0 < d.a2_rows) {
          for (int kb = 0; kb < num_kb2; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), L::A_BYTES + L::B_STAGE_BYTES);
            tma_load_2d(smem_base + L::A_OFF + stage * L::A_BYTES, &tmA2, kb * BK, (int)m0, full_bar(stage));
            tma_load_2d(smem_base + L::B_OFF + stage * L::B_STAGE_BYTES, &tmB, (num_kb + kb) * BK, nt * BN, full_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const long m0 = d.m_begin + (long)((int)fd_nt.div((uint32_t)tile)) * BM;
        const int kb_total = num_kb + (m0 < d.a2_rows ? num_kb2 : 0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc(smem_base + L::A_OFF + stage * L::A_BYTES);
          const uint64_t bdesc = make_smem_desc(smem_base + L::B_OFF + stage * L::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr>>4) start field
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));          // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));              // accumulator complete -> epilogue
      }
    }
  } else if constexpr (NEW == 16) {
    static_assert(EPI == 0 || EPI == 8 || EPI == 9, "16 epilogue warps: lean epilogue (row-per-thread, TMA-store or DMA) only");
    // warps 2, 3: idle members of the producer warpgroup -- with EPI 9 their lane 0 is a DMA thread of the epilogue
    if constexpr (EPI == 9) {
      if (lane == 0)
        epilogue_dma_thread<BN>(d, &tmD, &tmR, rfull_bar(0), rfull_bar(8), (warp - 2) * 2, num_tiles, num_n_tiles, fd_nt,
                                smem_base + L::OB_OFF);
    }
  } else {
    static_assert(EPI == 0, "8 epilogue warps: the row-per-thread role");
    epilogue_role<BN, NEW, false, 2>(d, reinterpret_cast<const float*>(smem + L::BIAS_OFF), tmem_base, tfull_bar(0),
                                           tempty_bar(0), warp, lane, num_tiles, num_n_tiles, fd_nt,
                                           PatchMap{0, 0, 0, FastDiv{1, 0, 0}});
  }
  }

  // ===================== teardown =====================
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ===============================================================================================================
// CTA-pair kernel (cta_group::2): a cluster of two CTAs (the two SMs of a TPC) computes a 256 x 256 output tile = two
// M-adjacent 128-row tiles with the same 256 output channels.  Per k-block each CTA loads its own 128 x 64 A tile and HALF
// of the 256 x 64 weight tile (32 KB instead of 48 KB per SM); the leader CTA's single thread issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), each CTA's TMEM receives its own 128 accumulator rows, and each CTA
// runs the 16-warp lean epilogue on them.
//   full[s]   (leader's): leader expects the bytes of BOTH CTAs; the peer's TMA loads signal it (cta_group::2 loads)
//   empty[s]  (each CTA's own): multicast tcgen05.commit of the leader
//   tfull[a]  (each CTA's own): multicast tcgen05.commit
//   tempty[a] (leader's): 32 arrivals = 16 epilogue warps x 2 CTAs (the peer arrives remotely)
// ===============================================================================================================
template <int BN, int STAGES>
struct Smem2 {
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_STAGE_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;              // full[S], empty[S], tfull[2], tempty[2]
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * STAGES + 4) * 8;
  static constexpr int BIAS_OFF = (TMEM_PTR_OFF + 16 + 15) & ~15;
  static constexpr int TOTAL = BIAS_OFF + MAX_BIAS * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static_assert(DYN_BYTES <= 232448, "shared memory plan exceeds 227 KB");
};

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1)
gemm_sm100_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                       const __grid_constant__ CUtensorMap tmBh, const GemmDesc d, const int num_m_tiles,
                       const int num_n_tiles, const FastDiv fd_nt) {
  using L = Smem2<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = tmem_cols(BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_first = (int)(blockIdx.x >> 1), pair_stride = (int)(gridDim.x >> 1);
  const int num_pairs = ((num_m_tiles + 1) >> 1) * num_n_tiles;
  const int kb_per_tap = d.Cin / BK;
  const int num_kb = d.ntaps * kb_per_tap;
  const int num_kb2 = d.A2 != nullptr ? d.Cin2 / BK : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBh) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (d.bias != nullptr) {
    float* sb = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    for (int i = threadIdx.x; i < d.Cout; i += 640) sb[i] = d.bias[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised and its TMEM is allocated before anything crosses over
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    epilogue_lean_role<BN, true>(d, reinterpret_cast<const float*>(smem + L::BIAS_OFF), tmem_base, tfull_bar(0), tempty_bar(0),
                                 warp, lane, num_pairs, num_n_tiles, fd_nt);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      // ===================== TMA producer (both CTAs) =====================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int pair = pair_first; pair < num_pairs; pair += pair_stride) {
          const int tq = (int)fd_nt.div((uint32_t)pair), nt = pair - tq * num_n_tiles;
          const long m0_pair = d.m_begin + (long)(2 * tq) * BM;
          const long m0 = m0_pair + (long)rank * BM;
          const int n0 = nt * BN + (int)rank * (BN / 2);
          int tap = 0, cb = 0;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t lfull = mapa_rank(full_bar(stage), 0);
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
            tma_load_2d_2cta(smem_base + L::A_OFF + stage * A_STAGE_BYTES, &tmA, cb * BK, (int)(m0 + d.row_off[tap]), lfull);
            tma_load_2d_2cta(smem_base + L::B_OFF + stage * L::B_HALF_BYTES, &tmBh, kb * BK, n0, lfull);
            if (++cb == kb_per_tap) { cb = 0; ++tap; }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (m0_pair < d.a2_rows) {
            for (int kb = 0; kb < num_kb2; ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1);
              const uint32_t lfull = mapa_rank(full_bar(stage), 0);
              if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
              tma_load_2d_2cta(smem_base + L::A_OFF + stage * A_STAGE_BYTES, &tmA2, kb * BK, (int)m0, lfull);
              tma_load_2d_2cta(smem_base + L::B_OFF + stage * L::B_HALF_BYTES, &tmBh, (num_kb + kb) * BK, n0, lfull);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (one thread of the leader CTA) =====================
      if (lane == 0 && rank == 0) {
        constexpr uint32_t idesc = make_idesc(2 * BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int pair = pair_first; pair < num_pairs; pair += pair_stride, ++it) {
          const int acc = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          const long m0_pair = d.m_begin + (long)(2 * (int)fd_nt.div((uint32_t)pair)) * BM;
          const int kb_total = num_kb + (m0_pair < d.a2_rows ? num_kb2 : 0);
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < kb_total; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint64_t adesc = make_smem_desc(smem_base + L::A_OFF + stage * A_STAGE_BYTES);
            const uint64_t bdesc = make_smem_desc(smem_base + L::B_OFF + stage * L::B_HALF_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2cta(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit_2cta(tfull_bar(acc));
        }
      }
    }
  }

  // ===================== teardown: nobody leaves while the peer may still signal into this CTA =====================
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ===============================================================================================================
// Back-to-back GEMM: a 256-wide stage-1 tile (1x1 expansion + residual / second operand + ReLU / masks, exactly the lean
// kernel's work) is ALSO the complete K = 256 operand row block of the next layer's 1x1 reduction (256 -> 64), so the
// second GEMM runs on the tile while it is still on chip and the next layer never re-reads the 256-channel tensor from HBM:
//   stage 1: acc1[128 x 256] = A * W1^T (+ A2 * W1b^T)      smem operands, TMEM columns [0, 256), ONE buffer
//            epilogue 1 = the DMA-thread lean role (epilogue_lean_dma_role): eight half boxes of 128 x 32 (SWIZZLE_64B); a
//            half box receives the tile's residual by TMA load, is overwritten in place with the bf16 output, and is then
//            (a) the source of the TMA store that writes D and (b) a K-major operand tile (K = 32) of stage 2 as it stands.
//   stage 2: acc2[128 x 64] += box_k (smem, SWIZZLE_64B descriptor) * W2[:, 32 k ..]^T for the eight boxes;  W2 (64 x 256,
//            32 KB, SWIZZLE_128B) resident in smem, TMEM columns 256 + 64 * buf (two buffers)
//            epilogue 2 (the same 16 warps, 16 columns each): bias2 / ReLU / bit mask -> bf16 -> global D2 (+ sign bits)
// acc1 has ONE buffer (2 x 256 + 2 x 64 columns do not fit in TMEM), but it is released as soon as the last tcgen05.ld of a
// tile has completed (t1empty); the MMAs of the next tile (K = 64: 4 instructions) run underneath the rest of the epilogue.
// MMA thread:  S1(0);  for t: [stage 2 of the four half-0 boxes of t as they become ready]; S1(t+1) once t1empty(t);
// [stage 2 of the half-1 boxes of t]; commit t2full(t).   DMA threads (lane 0 of warps 2, 3; two parts each), per box:
// wait ready -> TMA store -> wait for the store's read AND for stage 2's read (bfree) -> refill with the next tile's residual.
// Epilogue warps per tile t: E1(t), then E2(t-1).
// History (profiles/README.md): round 1 fed stage 2 from TMEM with tcgen05.st and released acc1 after the whole epilogue
// (1.59 ms per fused launch against 0.70 + 0.41); round 2's first rewrite used 128 x 64 boxes with register-prefetched
// residuals (time-neutral: the tile time was set by the LSU-bound epilogue, to which epilogue 2 was added).
// ===============================================================================================================
struct B2bParams {
  const float* bias2;
  void* D2; int ldd2;
  int relu2;
  const uint32_t* mask_bits2; int ld_mb2;
  uint32_t* D2_bits; int ld_db2;
};
constexpr int B2B_N2 = 64;
constexpr int B2B_STAGES = 2;
struct SmemB2b {
  static constexpr int B_STAGE_BYTES = 256 * BK * 2;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = B2B_STAGES * A_STAGE_BYTES;
  static constexpr int W2_OFF = B_OFF + B2B_STAGES * B_STAGE_BYTES;          // 4 k-blocks of [64 rows x 128 B]
  static constexpr int W2_BYTES = B2B_N2 * 256 * 2;
  static constexpr int OB_OFF = W2_OFF + W2_BYTES;                           // 8 half boxes of 128 rows x 64 B
  static constexpr int BAR_OFF = OB_OFF + 8 * EPI_BOX_BYTES;   // full[S], empty[S], t1full, t1empty, t2full[2], t2empty[2], rfull[8], ready[8], bfree[8], w2full
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * B2B_STAGES + 32) * 8;
  static constexpr int BIAS_OFF = (TMEM_PTR_OFF + 16 + 15) & ~15;
  static constexpr int BIAS2_OFF = BIAS_OFF + 256 * 4;
  static constexpr int TOTAL = BIAS2_OFF + B2B_N2 * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static_assert(OB_OFF % 1024 == 0 && W2_OFF % 1024 == 0, "operand tiles must be 1024-byte aligned");
  static_assert(DYN_BYTES <= 232448, "shared memory plan exceeds 227 KB");
};

// K-major SWIZZLE_64B operand tile (rows of 64 B, 8-row groups 512 B apart): layout type 4
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  desc |= (uint64_t)1 << 16;
  desc |= (uint64_t)(512 >> 4) << 32;
  desc |= (uint64_t)1 << 46;
  desc |= (uint64_t)4 << 61;
  return desc;
}

__global__ void __launch_bounds__(640, 1)
gemm_b2b_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmW2,
                const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR, const GemmDesc d,
                const B2bParams q2, const int num_tiles) {
  using L = SmemB2b;
  constexpr int BN = 256, N2 = B2B_N2, STAGES = B2B_STAGES;
  constexpr uint32_t ACC2_COL = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t t1full = bar_base + 8u * (2 * STAGES);
  const uint32_t t1empty = bar_base + 8u * (2 * STAGES + 1);
  auto t2full = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  auto t2empty = [&](int b) { return bar_base + 8u * (2 * STAGES + 4 + b); };
  auto rfull = [&](int k) { return bar_base + 8u * (2 * STAGES + 6 + k); };
  auto ready = [&](int k) { return bar_base + 8u * (2 * STAGES + 14 + k); };
  auto bfree = [&](int k) { return bar_base + 8u * (2 * STAGES + 22 + k); };
  const uint32_t w2full = bar_base + 8u * (2 * STAGES + 30);
  auto box_of = [&](int k) { return smem_base + L::OB_OFF + (uint32_t)(k * EPI_BOX_BYTES); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_per_tap = d.Cin / BK;
  const int num_kb = d.ntaps * kb_per_tap;
  const int num_kb2 = d.A2 != nullptr ? d.Cin2 / BK : 0;
  const bool has_res = d.res != nullptr;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmD) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmR) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(t1full, 1);
    mbar_init(t1empty, 16);
    for (int b = 0; b < 2; ++b) {
      mbar_init(t2full(b), 1);
      mbar_init(t2empty(b), 16);
    }
    for (int k = 0; k < 8; ++k) {
      mbar_init(rfull(k), 1);
      mbar_init(ready(k), 4);
      mbar_init(bfree(k), 1);
    }
    mbar_init(w2full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    float* sb = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    for (int i = threadIdx.x; i < 256; i += 640) sb[i] = d.bias != nullptr ? d.bias[i] : 0.f;
    float* sb2 = reinterpret_cast<float*>(smem + L::BIAS2_OFF);
    for (int i = threadIdx.x; i < N2; i += 640) sb2[i] = q2.bias2 != nullptr ? q2.bias2[i] : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ===================== epilogue warps: TMEM lane quarter (warp & 3) x 64-column part =====================
    constexpr int CH = 16, CPW = 4;
    const int qd = warp & 3;
    const int part = (warp - 4) >> 2;
    const int row = qd * 32 + lane;
    const int col0 = part * 64;
    const float* sbias = reinterpret_cast<const float*>(smem + L::BIAS_OFF);
    const float* sbias2 = reinterpret_cast<const float*>(smem + L::BIAS2_OFF);
    const uint32_t* mbits = d.mask_bits;
    const int m_end = (int)d.m_end;                    // all row indices are < 2^31 (checked where the plan is built)
    const int res_rows = (int)(d.res_rows < d.m_end ? d.res_rows : d.m_end);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t row_off = (uint32_t)row * 64u;      // this thread's 64-byte row inside a half box (SWIZZLE_64B)
    const uint32_t swz = (uint32_t)((row >> 1) & 3);
    int dest_prev = -1;
    uint32_t wmask2_prev = 0xFFFFFFFFu;

    // wraw: the word holding this row's 16 stage-2 mask bits, loaded one tile earlier and not touched until here
    auto stage2 = [&](int it_prev, int dest, uint32_t wraw) {      // epilogue 2 of the tile handled one iteration ago
      const int b = it_prev & 1;
      mbar_wait(t2full(b), (uint32_t)((it_prev >> 1) & 1));
      tcgen05_fence_after();
      const int n0 = part * CH;
      const uint32_t wmask = (wraw >> (16 * (part & 1))) & 0xFFFFu;
      uint32_t wout = 0u;
#pragma unroll
      for (int h = 1; h >= 0; --h) {                 // two halves of 8 columns, upper half first (bit order of wout)
        uint32_t r[8];
        tmem_ld<8>(lane_addr + ACC2_COL + (uint32_t)(b * N2 + n0 + 8 * h), r);
        tmem_ld_wait();
        if (h == 0) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t2empty(b));    // the accumulator is in registers: stage 2 of tile it_prev + 2 may overwrite it
        }
        if (dest >= 0) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 3; j >= 0; --j) {
            const int c = 8 * h + 2 * j;
            float hi = __uint_as_float(r[2 * j + 1]) + sbias2[n0 + c + 1], lo = __uint_as_float(r[2 * j]) + sbias2[n0 + c];
            if (q2.relu2) {
              hi = fmaxf(hi, 0.f); lo = fmaxf(lo, 0.f);
              wout = push_positive_bit(push_positive_bit(wout, hi), lo);
            } else {
              wout = (wout << 2) | (hi > 0.f ? 2u : 0u) | (lo > 0.f ? 1u : 0u);
            }
            hi = keep_if_bit(hi, wmask, c + 1); lo = keep_if_bit(lo, wmask, c);
            pk[j] = pack_bf16x2(lo, hi);
          }
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(q2.D2) + (long)dest * q2.ldd2 + n0 + 8 * h) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      if (dest >= 0 && q2.D2_bits != nullptr)
        reinterpret_cast<uint16_t*>(q2.D2_bits + bits_index(dest, n0 / 32, q2.ld_db2))[part & 1] = (uint16_t)(wout & wmask);
    };

    uint32_t bits_nxt0 = 0u, bits_nxt1 = 0u;
    auto prefetch_bits = [&](int tile) {
      bits_nxt0 = bits_nxt1 = 0u;
      if (tile >= num_tiles) return;
      const int m = (int)d.m_begin + tile * BM + row;
      if (mbits != nullptr && m < m_end) {
        bits_nxt0 = __ldg(mbits + bits_index(m, col0 / 32, d.ld_mb));
        bits_nxt1 = __ldg(mbits + bits_index(m, col0 / 32 + 1, d.ld_mb));
      }
    };
    prefetch_bits(blockIdx.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m = (int)d.m_begin + tile * BM + row;
      int dest = -1;                                       // DST_SAME: the destination row is the source row (pad rows: none)
      if (m < m_end) dest = (int)map_row(d.src, d.dst_kind, d.dst, m);
      const bool live = dest >= 0;
      const bool use_res = live && has_res && m < res_rows;
      const uint32_t bits_cur[2] = {bits_nxt0, bits_nxt1};
      uint32_t bits_out[2] = {0u, 0u};
      uint32_t wmask2 = 0xFFFFFFFFu;
      if (live && q2.mask_bits2 != nullptr) wmask2 = __ldg(q2.mask_bits2 + bits_index(dest, (part * CH) / 32, q2.ld_mb2));
      prefetch_bits(tile + gridDim.x);
      mbar_wait(t1full, (uint32_t)(it & 1));
      tcgen05_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = part * 2 + h;
        // the box is ours again: the TMA store AND stage 2 of the previous tile have read it, this tile's residual is in it
        mbar_wait(rfull(k), (uint32_t)(it & 1));
        const uint32_t brow = box_of(k) + row_off;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int ci = 2 * h + cc;
          uint32_t r[CH];
          tmem_ld<CH>(lane_addr + (uint32_t)(col0 + ci * CH), r);
          const uint32_t a0 = brow + (((uint32_t)(2 * cc) ^ swz) << 4), a1 = brow + (((uint32_t)(2 * cc + 1) ^ swz) << 4);
          uint32_t rs[8];
          if (use_res) { lds128(a0, rs); lds128(a1, rs + 4); }
          tmem_ld_wait();
          if (ci == CPW - 1) {
            // this warp's part of acc1 is in registers: once all 16 warps are here the next tile's MMAs may overwrite it
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t1empty);
          }
          if (live) {
            const int n0 = col0 + ci * CH;
            float v[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]) + sbias[n0 + j];
            if (use_res) {
#pragma unroll
              for (int j = 0; j < CH / 2; ++j) { v[2 * j] += bf16_lo(rs[j]); v[2 * j + 1] += bf16_hi(rs[j]); }
            }
            uint32_t wout = 0u;
            if (d.relu) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
              if (d.D_bits != nullptr) {
#pragma unroll
                for (int j = CH - 1; j >= 0; --j) wout = push_positive_bit(wout, v[j]);
              }
            } else if (d.D_bits != nullptr) {
#pragma unroll
              for (int j = 0; j < CH; ++j) wout |= (v[j] > 0.f ? 1u : 0u) << j;
            }
            if (mbits != nullptr) {
              const uint32_t w = (bits_cur[h] >> (16 * cc)) & 0xFFFFu;
              wout &= w;
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = keep_if_bit(v[j], w, j);
            }
            bits_out[h] |= wout << (16 * cc);
            uint32_t pk[CH / 2];
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            sts128(a0, pk);
            sts128(a1, pk + 4);
          } else {
            const uint32_t z[4] = {0u, 0u, 0u, 0u};       // pad rows / rows past the end: zero in D and zero in the stage-2 operand
            sts128(a0, z);
            sts128(a1, z);
          }
        }
        fence_async_smem();                 // the box is read by the async proxy (TMA store, tcgen05.mma)
        __syncwarp();
        if (lane == 0) mbar_arrive(ready(k));
      }
      if (live && d.D_bits != nullptr) {
        d.D_bits[bits_index(dest, col0 / 32, d.ld_db)] = bits_out[0];
        d.D_bits[bits_index(dest, col0 / 32 + 1, d.ld_db)] = bits_out[1];
      }
      if (it > 0) stage2(it - 1, dest_prev, wmask2_prev);
      dest_prev = dest;
      wmask2_prev = wmask2;
    }
    if (it > 0) stage2(it - 1, dest_prev, wmask2_prev);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_expect_tx(w2full, L::W2_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem_base + L::W2_OFF + kb * (N2 * 128), &tmW2, kb * BK, 0, w2full);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          const long m0 = d.m_begin + (long)tile * BM;
          int tap = 0, cb = 0;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + L::B_STAGE_BYTES);
            tma_load_2d(smem_base + L::A_OFF + stage * A_STAGE_BYTES, &tmA, cb * BK, (int)(m0 + d.row_off[tap]), full_bar(stage));
            tma_load_2d(smem_base + L::B_OFF + stage * L::B_STAGE_BYTES, &tmB, kb * BK, 0, full_bar(stage));
            if (++cb == kb_per_tap) { cb = 0; ++tap; }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (m0 < d.a2_rows) {
            for (int kb = 0; kb < num_kb2; ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1);
              mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + L::B_STAGE_BYTES);
              tma_load_2d(smem_base + L::A_OFF + stage * A_STAGE_BYTES, &tmA2, kb * BK, (int)m0, full_bar(stage));
              tma_load_2d(smem_base + L::B_OFF + stage * L::B_STAGE_BYTES, &tmB, (num_kb + kb) * BK, 0, full_bar(stage));
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (one thread) =====================
      if (lane == 0) {
        constexpr uint32_t idesc1 = make_idesc(BM, BN);
        constexpr uint32_t idesc2 = make_idesc(BM, N2);
        int stage = 0;
        uint32_t phase = 0;
        auto issue_stage1 = [&](int tile, int it) {
          const long m0 = d.m_begin + (long)tile * BM;
          const int kb_total = num_kb + (m0 < d.a2_rows ? num_kb2 : 0);
          mbar_wait(t1empty, (uint32_t)((it & 1) ^ 1));       // all 16 epilogue warps hold tile it - 1's accumulator in registers
          tcgen05_fence_after();
#pragma unroll 1
          for (int kb = 0; kb < kb_total; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint64_t adesc = make_smem_desc(smem_base + L::A_OFF + stage * A_STAGE_BYTES);
            const uint64_t bdesc = make_smem_desc(smem_base + L::B_OFF + stage * L::B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc1, (kb | k) != 0 ? 1u : 0u);
            umma_commit(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(t1full);
        };
        // stage 2 of the four boxes of half h of tile `it`: box k = part * 2 + h holds output columns [64 part + 32 h, +32)
        auto issue_stage2_half = [&](int it, int h) {
          const int b = it & 1;
          const uint32_t tmem_d2 = tmem_base + ACC2_COL + (uint32_t)(b * N2);
          if (h == 0) mbar_wait(t2empty(b), (uint32_t)(((it >> 1) & 1) ^ 1));    // epilogue 2 of tile it - 2 has drained this buffer
#pragma unroll 1
          for (int p = 0; p < 4; ++p) {
            const int k = p * 2 + h;
            mbar_wait(ready(k), (uint32_t)(it & 1));
            tcgen05_fence_after();
            const uint64_t adesc = make_smem_desc_sw64(box_of(k));
            const uint64_t bdesc = make_smem_desc(smem_base + L::W2_OFF + p * (N2 * 128)) + (uint64_t)(4 * h);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16(tmem_d2, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc2, (h | p | ks) != 0 ? 1u : 0u);
            umma_commit(bfree(k));                   // box k may be refilled once these MMAs have read it
          }
          if (h == 1) umma_commit(t2full(b));
        };
        mbar_wait(w2full, 0);
        int it = 0;
        if ((int)blockIdx.x < num_tiles) issue_stage1(blockIdx.x, 0);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
          issue_stage2_half(it, 0);
          if (tile + (int)gridDim.x < num_tiles) issue_stage1(tile + gridDim.x, it + 1);
          issue_stage2_half(it, 1);
        }
      }
    } else if (lane == 0) {
      // ===================== DMA threads (warps 2, 3): two parts each =====================
      const int part0 = (warp - 2) * 2;
      auto refill = [&](int k, int tile) {
        if (tile >= num_tiles) return;
        if (has_res) {
          mbar_expect_tx(rfull(k), EPI_BOX_BYTES);
          tma_load_2d(box_of(k), &tmR, (k >> 1) * 64 + (k & 1) * 32, (int)(d.m_begin + (long)tile * BM), rfull(k));
        } else {
          mbar_arrive(rfull(k));
        }
      };
#pragma unroll 1
      for (int s = 0; s < 4; ++s) refill((part0 + (s & 1)) * 2 + (s >> 1), blockIdx.x);
      int it = 0;
#pragma unroll 1
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int m0 = (int)(d.m_begin + (long)tile * BM);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          const int k = (part0 + (s & 1)) * 2 + (s >> 1);          // half 0 of both parts, then half 1 of both parts
          mbar_wait(ready(k), (uint32_t)(it & 1));
          tma_store_2d(&tmD, box_of(k), (k >> 1) * 64 + (k & 1) * 32, m0);
          bulk_commit();
          bulk_wait_read0();                                        // the store has read the box ...
          mbar_wait(bfree(k), (uint32_t)(it & 1));                  // ... and so have the stage-2 MMAs
          refill(k, tile + (int)gridDim.x);
        }
      }
      bulk_wait0();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ===============================================================================================================
// Patch-tile variant for the small-channel multi-tap convolutions (Cin = 64, Cout <= 64: conv1 forward / input
// gradient, the 3x3 convs of layer1).  With flat 128-row tiles every tap re-loads its own 16 KB operand tile, and these
// layers run at the L2->SM bandwidth ceiling (~42 B/clk/SM) instead of the tensor or HBM roofline.  Here
//   * an output tile is a 16 x 8 pixel patch; for each horizontal tap dx ONE box of (16 + ny - 1) x 8 pixel rows is
//     loaded (3-D TMA map [C, P, lines]; out-of-range columns / lines are zero-filled = conv padding) and serves all
//     ny vertical taps: vertical tap yi is the same box read through a matrix descriptor that starts yi * 8 rows
//     (= yi * 1024 B, a whole swizzle atom) further down;
//   * the whole weight matrix (<= 72 KB) is loaded once per CTA and stays resident in shared memory.
// Operand traffic per tile drops from ntaps * (16 KB + BN * 128 B) to nx * (18..19 KB).
// ===============================================================================================================
constexpr int PATCH_SLAB_H = 19;                               // 16 + up to 3 halo lines
constexpr int PATCH_SLAB_BYTES = PATCH_SLAB_H * 8 * 128;       // 19456 (a multiple of 1024)
#ifndef RGIE_PATCH_SLABS
#define RGIE_PATCH_SLABS 6
#endif
constexpr int PATCH_SLABS = RGIE_PATCH_SLABS;
constexpr int PATCH_W_BYTES = 72 * 1024;                       // resident weights: ntaps * BN * 128 B <= 72 KB
struct PatchSmem {
  static constexpr int A_OFF = 0;
  static constexpr int W_OFF = PATCH_SLABS * PATCH_SLAB_BYTES;
  static constexpr int BAR_OFF = W_OFF + PATCH_W_BYTES;        // afull[S], aempty[S], tfull[2], tempty[2], wfull
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * PATCH_SLABS + 10) * 8;  // afull, aempty, tfull[4], tempty[4], wfull (+1: keeps BIAS_OFF 16-byte aligned)
  static constexpr int BIAS_OFF = TMEM_PTR_OFF + 16;
  static constexpr int TOTAL = BIAS_OFF + 64 * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};
static_assert(PatchSmem::DYN_BYTES <= 232448, "patch kernel shared memory plan exceeds 227 KB");

struct PatchTaps {
  int ny, nx, dy0, dx0;          // taps (dy0 + yi, dx0 + xi)
  int tap[4][4];                 // index of tap (yi, xi) in the weight matrix (column block), -1 = absent
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

template <int BN, int NY, int NX>
__global__ void __launch_bounds__(num_threads(8), 1)
gemm_patch_kernel(const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB, const GemmDesc d,
                  const PatchTaps tp, const PatchMap pm, const int num_tiles) {
  using L = PatchSmem;
  constexpr int NACC = 4;                         // accumulator buffers: two tiles in the epilogue, two in the tensor pipe
  constexpr uint32_t TMEM_COLS = tmem_cols(BN, NACC);
  constexpr int W_TAP_BYTES = BN * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (PATCH_SLABS + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PATCH_SLABS + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PATCH_SLABS + NACC + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * PATCH_SLABS + 2 * NACC);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t slab_bytes = (uint32_t)((16 + NY - 1) * 8 * 128);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < PATCH_SLABS; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);     // the 4 warps of the group that owns the tile
    }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (d.bias != nullptr) {
    float* sb = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    for (int i = threadIdx.x; i < d.Cout; i += num_threads(8)) sb[i] = d.bias[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer: resident weights once, then one box per (tile, horizontal tap) =====================
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, (uint32_t)(d.ntaps * W_TAP_BYTES));
      for (int t = 0; t < d.ntaps; ++t)
        tma_load_2d(smem_base + L::W_OFF + t * W_TAP_BYTES, &tmB, t * BK, 0, wfull_bar);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ht = (int)pm.fd_wt.div((uint32_t)tile), wt = tile - ht * pm.WT;
#pragma unroll
        for (int xi = 0; xi < NX; ++xi) {
          mbar_wait(aempty_bar(slot), phase ^ 1);
          mbar_expect_tx(afull_bar(slot), slab_bytes);
          tma_load_3d(smem_base + L::A_OFF + slot * PATCH_SLAB_BYTES, &tmA3, 0, wt * 8 + tp.dx0 + xi, ht * 16 + tp.dy0,
                      afull_bar(slot));
          if (++slot == PATCH_SLABS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      mbar_wait(wfull_bar, 0);
      const uint64_t wdesc0 = make_smem_desc(smem_base + L::W_OFF);
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it % NACC;
        const uint32_t acc_phase = (it / NACC) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        // the single issuing thread is the critical resource of these narrow-N tiles (an MMA is 45-48 cycles of tensor
        // time): the tap loops are fully unrolled, descriptors differ by compile-time constants, the tap table sits in
        // the constant bank
#pragma unroll
        for (int xi = 0; xi < NX; ++xi) {
          mbar_wait(afull_bar(slot), phase);
          tcgen05_fence_after();
          const uint64_t adesc0 = make_smem_desc(smem_base + L::A_OFF + slot * PATCH_SLAB_BYTES);
#pragma unroll
          for (int yi = 0; yi < NY; ++yi) {
            // vertical tap yi: the same box, yi pixel lines (= yi swizzle atoms of 8 rows x 128 B = 64 descriptor units) down
            const uint64_t adesc = adesc0 + (uint64_t)(yi * 64);
            const uint64_t bdesc = wdesc0 + (uint64_t)(tp.tap[yi][xi] * (W_TAP_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (xi | yi | k) != 0 ? 1u : 0u);
          }
          umma_commit(aempty_bar(slot));
          if (++slot == PATCH_SLABS) { slot = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    epilogue_role<BN, 8, true, NACC>(d, reinterpret_cast<const float*>(smem + L::BIAS_OFF), tmem_base, tfull_bar(0),
                                     tempty_bar(0), warp, lane, num_tiles, 1, FastDiv{1, 0, 0}, pm);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ===============================================================================================================
// conv1 + max-pool in one launch.  The stem's 64-channel 224 x 224 activation is the largest tensor of the network
// (6.4 MB per crop in bf16): conv1 wrote it (2.05 GB per 320 crops) and the pooling kernel read it back.  Here the pooled
// tensor and the argmax bytes are produced straight from the accumulators:
//   * a tile is still a 16 x 8 conv patch (one 19-line TMA box, 16 MMAs).  Vertically it starts at conv line 14 ty - 1 of
//     ITS image, so its 16 lines hold 7 pooling-window rows completely (bands overlap by 2 lines: 1.14x the conv work).
//     Horizontally the tiles of a band do NOT overlap (columns 8k .. 8k+7): a CTA walks the W / 8 tiles of a band left to
//     right, and the one column a window needs from the tile on its left (window 4k = columns 8k-1, 8k, 8k+1) travels
//     through a 2 KB carry slot in shared memory.  (A first version with tiles overlapping both ways -- 1.5x the conv
//     work -- was correct and SLOWER than conv1 + maxpool_fwd_kernel, 2.03 vs 1.64 ms per 320 crops: the stem's main loop
//     costs ~1 650 cycles per tile, not the 16 x 72 of its MMAs.)
//   * an epilogue warp group turns its accumulator tile into post-ReLU bf16 rows in a 16 KB shared-memory tile
//     ([128 pixels][8 chunks of 16 B], chunk index XOR (pixel & 7): conflict-free both ways), releases the accumulator,
//     copies column 7 into carry slot (tile & 3), and 128 threads scan the 7 x 4 windows x 8 channel chunks with the
//     packed-pair compare / select of maxpool_fwd_kernel (first maximum in (dy, dx) scan order wins, strict >): pooled
//     values and argmax bytes are bit-identical to conv1 followed by maxpool_fwd_kernel.
//   * the two groups take alternate tiles; `carry_bar[g]` (one arrival per tile of group g) tells the other group that
//     tile s's column is in its slot.  A slot is rewritten four tiles later, after the writer has seen the carry signal of
//     tile s - 1, which its owner raises only after it has finished pooling tile s - 3, the last reader of that slot.
// Conv positions outside the image (line / column -1 or 224, the pool's padding) are computed and ignored.
// Shared memory: the plan of the patch kernel; the tiles and the carry slots live in the 40 KB of the weight area that
// conv1's four 8 KB tap matrices leave free.
// ===============================================================================================================
struct PoolSmem {
  static constexpr int TILE_OFF = PatchSmem::W_OFF + 4 * 64 * BK * 2;       // behind conv1's 32 KB of weights
  static constexpr int CARRY_OFF = TILE_OFF + 2 * 16384;                     // 4 slots x 2 KB
  static constexpr int END = CARRY_OFF + 4 * 2048;
};
static_assert(PoolSmem::END <= PatchSmem::W_OFF + PATCH_W_BYTES, "conv1 + pool: tiles and carry slots must fit the free weight area");

// 7 x 4 pooling windows x 8 channel chunks of one tile, 128 threads.  tile_s: [128 pixels][8 chunks ^ (pixel & 7)] bf16 of the
// 16 x 8 patch; left_s: column 7 of the tile on the left ([16 lines][8 chunks ^ (line & 7)]).  CHECK: conv positions outside
// the image are skipped (border tiles); the scan order (dy, dx) and the strict > keep the first maximum, as maxpool_fwd_kernel.
template <bool CHECK>
__device__ __forceinline__ void pool_scan(const uint8_t* tile_s, const uint8_t* left_s, const int tid, const StemPoolParams& sp,
                                          __nv_bfloat16* P1, const int n, const int ty, const int k, const int y0) {
  for (int item = tid; item < 28 * 8; item += 128) {
    const int ck = item & 7, pw = item >> 3;
    const int pr = pw >> 2, pc = pw & 3;
    const int i = 7 * ty + pr, j = 4 * k + pc;
    if (CHECK && i >= sp.Hp) continue;
    uint4 u[9];
    bool ok[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ly = 2 * pr + dy;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int lx = 2 * pc - 1 + dx;
        const int t9 = dy * 3 + dx;
        ok[t9] = !CHECK || (y0 + ly >= 0 && y0 + ly < sp.H0 && 8 * k + lx >= 0 && 8 * k + lx < sp.H0);
        const uint8_t* src = lx < 0 ? left_s + ly * 128 + ((ck ^ (ly & 7)) << 4)
                                    : tile_s + (ly * 8 + lx) * 128 + ((ck ^ ((ly * 8 + lx) & 7)) << 4);
        if (ok[t9]) u[t9] = *reinterpret_cast<const uint4*>(src);
      }
    }
    uint32_t best[4], code[4];
    bool first = true;
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) {
      if (CHECK && !ok[t9]) continue;
      const uint32_t w[4] = {u[t9].x, u[t9].y, u[t9].z, u[t9].w};
      const uint32_t tc = (uint32_t)t9 * 0x00010001u;
      if (first) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { best[e] = w[e]; code[e] = tc; }
        first = false;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&w[e]), *reinterpret_cast<const __nv_bfloat162*>(&best[e]));
          best[e] = (w[e] & m) | (best[e] & ~m);
          code[e] = (tc & m) | (code[e] & ~m);
        }
      }
    }
    *reinterpret_cast<uint4*>(P1 + geom_row(sp.g1, 0, n, i, j) * 64 + ck * 8) = make_uint4(best[0], best[1], best[2], best[3]);
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      code[e] |= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&best[e]), zero2) & 0x00100010u;
    uint2 pkc;
    pkc.x = __byte_perm(code[0], code[1], 0x6420);
    pkc.y = __byte_perm(code[2], code[3], 0x6420);
    *reinterpret_cast<uint2*>(sp.arg + (((long)n * sp.Hp + i) * sp.Hp + j) * 64 + ck * 8) = pkc;
  }
}

__global__ void __launch_bounds__(num_threads(8), 1)
gemm_conv1_pool_kernel(const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB, const GemmDesc d,
                       const PatchTaps tp, const StemPoolParams sp, const int num_bands) {
  using L = PatchSmem;
  constexpr int BN = 64, NY = 4;
  constexpr int NACC = 4;
  constexpr uint32_t TMEM_COLS = tmem_cols(BN, NACC);
  constexpr int W_TAP_BYTES = BN * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (PATCH_SLABS + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PATCH_SLABS + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PATCH_SLABS + NACC + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * PATCH_SLABS + 2 * NACC);
  // carry_bar[0]: the spare barrier slot of the patch plan; carry_bar[1]: the unused upper half of the TMEM-pointer slot
  auto carry_bar = [&](int g) { return g == 0 ? bar_base + 8u * (2 * PATCH_SLABS + 2 * NACC + 1) : smem_base + L::TMEM_PTR_OFF + 8u; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t slab_bytes = (uint32_t)((16 + NY - 1) * 8 * 128);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < PATCH_SLABS; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(carry_bar(0), 1);
    mbar_init(carry_bar(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    float* sb = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    for (int i = threadIdx.x; i < 64; i += num_threads(8)) sb[i] = d.bias != nullptr ? d.bias[i] : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // band -> (image n, band row ty); tile k of the band is the 16 x 8 patch at conv pixel (14 ty - 1, 8 k) of image n

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, (uint32_t)(d.ntaps * W_TAP_BYTES));
      for (int t = 0; t < d.ntaps; ++t)
        tma_load_2d(smem_base + L::W_OFF + t * W_TAP_BYTES, &tmB, t * BK, 0, wfull_bar);
      int slot = 0;
      uint32_t phase = 0;
      for (int band = blockIdx.x; band < num_bands; band += gridDim.x) {
        const int n = (int)sp.fd_img.div((uint32_t)band), ty = band - n * sp.TY;
        // source-grid line of conv line y of image n = n * lines_per_img + pad_t + y; the box starts dy0 lines above
        const int line = n * sp.lines_per_img + d.src.pad_t + (14 * ty - 1) + tp.dy0;
        for (int k = 0; k < sp.TX; ++k) {
          mbar_wait(aempty_bar(slot), phase ^ 1);
          mbar_expect_tx(afull_bar(slot), slab_bytes);
          tma_load_3d(smem_base + L::A_OFF + slot * PATCH_SLAB_BYTES, &tmA3, 0, d.src.pad_l + 8 * k + tp.dx0, line, afull_bar(slot));
          if (++slot == PATCH_SLABS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      mbar_wait(wfull_bar, 0);
      const uint64_t wdesc0 = make_smem_desc(smem_base + L::W_OFF);
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int band = blockIdx.x; band < num_bands; band += gridDim.x) {
        for (int k = 0; k < sp.TX; ++k, ++it) {
          const int acc = it % NACC;
          const uint32_t acc_phase = (it / NACC) & 1;
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
          mbar_wait(afull_bar(slot), phase);
          tcgen05_fence_after();
          const uint64_t adesc0 = make_smem_desc(smem_base + L::A_OFF + slot * PATCH_SLAB_BYTES);
#pragma unroll
          for (int yi = 0; yi < NY; ++yi) {
            const uint64_t adesc = adesc0 + (uint64_t)(yi * 64);
            const uint64_t bdesc = wdesc0 + (uint64_t)(tp.tap[yi][0] * (W_TAP_BYTES >> 4));
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma_bf16(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (yi | kk) != 0 ? 1u : 0u);
          }
          umma_commit(aempty_bar(slot));
          if (++slot == PATCH_SLABS) { slot = 0; phase ^= 1; }
          umma_commit(tfull_bar(acc));
        }
      }
    }
  } else {
    // ===================== epilogue + pooling: two groups of four warps take alternate tiles =====================
    const float* sbias = reinterpret_cast<const float*>(smem + L::BIAS_OFF);
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                            // accumulator row = patch pixel (row >> 3, row & 7)
    const int tid = ((warp - 2) & 3) * 32 + lane;             // 0..127 within the group (any order: used for the window scan)
    uint8_t* tile_s = smem + PoolSmem::TILE_OFF + grp * 16384;
    uint8_t* carry_s = smem + PoolSmem::CARRY_OFF;
    __nv_bfloat16* P1 = reinterpret_cast<__nv_bfloat16*>(sp.P1);
    int s = 0;                                                // index of the tile in this CTA's sequence
    int t = 0;                                                // index among this group's tiles
    for (int band = blockIdx.x; band < num_bands; band += gridDim.x) {
      const int n = (int)sp.fd_img.div((uint32_t)band), ty = band - n * sp.TY;
      const int y0 = 14 * ty - 1;
      for (int k = 0; k < sp.TX; ++k, ++s) {
        if ((s & 1) != grp) continue;
        const int acc = s % NACC;
        const uint32_t acc_phase = (s / NACC) & 1;
        mbar_wait(tfull_bar(acc), acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
        uint32_t r0[32], r1[32];
        tmem_ld<32>(taddr, r0);
        tmem_ld<32>(taddr + 32u, r1);
        tmem_ld_wait();
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          pk[j] = pack_bf16x2(fmaxf(__uint_as_float(r0[2 * j]) + sbias[2 * j], 0.f),
                              fmaxf(__uint_as_float(r0[2 * j + 1]) + sbias[2 * j + 1], 0.f));
          pk[16 + j] = pack_bf16x2(fmaxf(__uint_as_float(r1[2 * j]) + sbias[32 + 2 * j], 0.f),
                                   fmaxf(__uint_as_float(r1[2 * j + 1]) + sbias[32 + 2 * j + 1], 0.f));
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(tile_s + row * 128 + ((c ^ (row & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        // the carry signal of tile s - 1 (other group): its column is in slot (s - 1) & 3, and slot s & 3 is free again
        if (s >= 1) mbar_wait(carry_bar(grp ^ 1), (uint32_t)((grp == 1 ? t : t - 1) & 1));
        if ((row & 7) == 7) {
          uint8_t* cs = carry_s + (s & 3) * 2048 + (row >> 3) * 128;
          const int sw = (row >> 3) & 7;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(cs + ((c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        named_bar_sync(1 + grp, 128);
        if (tid == 0) mbar_arrive(carry_bar(grp));
        const uint8_t* left_s = carry_s + ((s - 1) & 3) * 2048;   // column 7 of the tile on the left (k > 0 only)
        // interior tiles (every conv position of the patch and of the carried column lies inside the image) take the scan
        // without bounds checks: nine independent 16-byte loads, then the compare / select chain
        const bool interior = k > 0 && y0 >= 0 && y0 + 15 < sp.H0 && 7 * ty + 6 < sp.Hp;
        if (interior) pool_scan<false>(tile_s, left_s, tid, sp, P1, n, ty, k, y0);
        else pool_scan<true>(tile_s, left_s, tid, sp, P1, n, ty, k, y0);
        named_bar_sync(1 + grp, 128);                         // the tile buffer is rewritten by this group's next tile
        ++t;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ===============================================================================================================
// CTA-pair form of the patch kernel (cta_group::2).  ncu on gemm_patch_kernel<64,3,3> (profiles/README.md, r2_d): the TC
// pipe is 77 % busy while the tensor datapath is 36 % active -- an N = 64, K = 16 instruction holds the pipe ~72 cycles
// for 32 cycles of math, so these layers are bound by the NUMBER of tcgen05.mma instructions.  One cta_group::2
// instruction (M = 256: the two SMs of a TPC, each with its own 16 x 8 pixel patch) does the work of two at the cost of
// one (tools/microbench/umma_rate_2cta.cu: 45.6 cycles per instruction for N <= 64).  Each CTA loads its own boxes and keeps
// HALF of the weight rows resident; barriers as in gemm_sm100_2cta_kernel: afull / wfull / tempty live in the leader CTA
// (the peer's TMA loads and epilogue warps signal them remotely), aempty / tfull are signalled in both CTAs by the
// leader's multicast tcgen05.commit.  The tile sequence of a CTA is unchanged (CTA b: tiles b, b + grid, ...), so CTAs
// 2c and 2c + 1 walk horizontally adjacent patches in lockstep; the plan rounds the tile count up to an even number (a
// phantom tile lies beyond the last line: its boxes are zero-filled and none of its rows is a row).
// ===============================================================================================================
__device__ __forceinline__ void tma_load_3d_2cta(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(leader_bar) : "memory");
}

template <int BN, int NY, int NX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(num_threads(8), 1)
gemm_patch_2cta_kernel(const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmBh, const GemmDesc d,
                       const PatchTaps tp, const PatchMap pm, const int num_tiles) {
  using L = PatchSmem;
  constexpr int NACC = 4;
  constexpr uint32_t TMEM_COLS = tmem_cols(BN, NACC);
  constexpr int W_TAP_HALF = (BN / 2) * BK * 2;             // this CTA's half of one tap's weight rows
  static_assert(BN % 16 == 0 && (BN / 2) % 8 == 0, "half weight tiles must be whole swizzle atoms");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (PATCH_SLABS + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PATCH_SLABS + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PATCH_SLABS + NACC + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * PATCH_SLABS + 2 * NACC);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  constexpr uint32_t slab_bytes = (uint32_t)((16 + NY - 1) * 8 * 128);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBh) : "memory");
    for (int s = 0; s < PATCH_SLABS; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);     // the 4 warps of the owning group in BOTH CTAs
    }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (d.bias != nullptr) {
    float* sb = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    for (int i = threadIdx.x; i < d.Cout; i += num_threads(8)) sb[i] = d.bias[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised and its TMEM is allocated before anything crosses over
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): this CTA's half of the weights once, then its own boxes =====================
    if (lane == 0) {
      const uint32_t lwfull = mapa_rank(wfull_bar, 0);
      if (rank == 0) mbar_expect_tx(wfull_bar, (uint32_t)(2 * d.ntaps * W_TAP_HALF));
      for (int t = 0; t < d.ntaps; ++t)
        tma_load_2d_2cta(smem_base + L::W_OFF + t * W_TAP_HALF, &tmBh, t * BK, (int)rank * (BN / 2), lwfull);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ht = (int)pm.fd_wt.div((uint32_t)tile), wt = tile - ht * pm.WT;
#pragma unroll
        for (int xi = 0; xi < NX; ++xi) {
          mbar_wait(aempty_bar(slot), phase ^ 1);
          const uint32_t lfull = mapa_rank(afull_bar(slot), 0);
          if (rank == 0) mbar_expect_tx(afull_bar(slot), 2 * slab_bytes);
          tma_load_3d_2cta(smem_base + L::A_OFF + slot * PATCH_SLAB_BYTES, &tmA3, 0, wt * 8 + tp.dx0 + xi, ht * 16 + tp.dy0, lfull);
          if (++slot == PATCH_SLABS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN);
      mbar_wait(wfull_bar, 0);
      tcgen05_fence_after();
      const uint64_t wdesc0 = make_smem_desc(smem_base + L::W_OFF);
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it % NACC;
        const uint32_t acc_phase = (it / NACC) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
#pragma unroll
        for (int xi = 0; xi < NX; ++xi) {
          mbar_wait(afull_bar(slot), phase);
          tcgen05_fence_after();
          const uint64_t adesc0 = make_smem_desc(smem_base + L::A_OFF + slot * PATCH_SLAB_BYTES);
#pragma unroll
          for (int yi = 0; yi < NY; ++yi) {
            const uint64_t adesc = adesc0 + (uint64_t)(yi * 64);
            const uint64_t bdesc = wdesc0 + (uint64_t)(tp.tap[yi][xi] * (W_TAP_HALF >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2cta(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (xi | yi | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(aempty_bar(slot));
          if (++slot == PATCH_SLABS) { slot = 0; phase ^= 1; }
        }
        umma_commit_2cta(tfull_bar(acc));
      }
    }
  } else {
    epilogue_role<BN, 8, true, NACC, true>(d, reinterpret_cast<const float*>(smem + L::BIAS_OFF), tmem_base, tfull_bar(0),
                                                  tempty_bar(0), warp, lane, num_tiles, 1, FastDiv{1, 0, 0}, pm);
  }

  // ===================== teardown: nobody leaves while the peer may still signal into this CTA =====================
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ===============================================================================================================
// conv1 input gradient, "horizontal taps as N":  the 7x7/2 stem's input gradient in the 2x2 phase-split form is a 4x4-tap
// convolution of the 64-channel gradient with only 12 outputs per pixel.  With N = 12(16) every one of the 64 MMAs of a
// tile re-reads its 128 x 16 operand slice from shared memory for 16 columns of output -- the patch kernel above is
// bound by those reads.  Here the four HORIZONTAL taps become the N dimension instead:
//     D[r, (j, q)] = sum_yi sum_co  dC1[r + (yi + dy0) lines, co] * W[(j, q), yi, co]          N = 4 x 12 = 48, K = 4 x 64
// over a source patch of 8 lines x 16 pixels (one 3-D TMA box of 11 lines serves the 4 vertical taps), 16 MMAs per
// tile instead of 64, and the epilogue finishes the horizontal part in registers:
//     dZ[(l, w), q] = sum_j D[(l, w + j + dx0), (j, q)]                                       warp shuffles inside a 16-lane line
// A tile therefore yields 8 x 13 output pixels (source columns 1..13 of its 16); tiles step by 13 columns.
// ===============================================================================================================
constexpr int HS_NJ = 4, HS_NQ = 12, HS_N = HS_NJ * HS_NQ, HS_NY = 4, HS_OUTW = 16 - (HS_NJ - 1);
constexpr int HS_SLAB_BYTES = (8 + HS_NY - 1) * 16 * 128;      // 22528
#ifndef RGIE_HS_SLABS
#define RGIE_HS_SLABS 6
#endif
constexpr int HS_SLABS = RGIE_HS_SLABS;
constexpr int HS_W_TAP_BYTES = HS_N * BK * 2;                   // 6144
struct HsSmem {
  static constexpr int A_OFF = 0;
  static constexpr int W_OFF = HS_SLABS * HS_SLAB_BYTES;
  static constexpr int BAR_OFF = W_OFF + HS_NY * HS_W_TAP_BYTES;   // afull[S], aempty[S], tfull[4], tempty[4], wfull
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * HS_SLABS + 10) * 8;
  static constexpr int TOTAL = TMEM_PTR_OFF + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};
struct HsParams {
  int WT;            // column tiles per line group
  FastDiv fd_wt;
  int dy0, dx0;      // first vertical / horizontal tap offset (lines / pixels)
  int col0;          // source column of tile 0, lane 0 (pad_l + dx0 so that output lane -dx0 is the first valid pixel)
  long lines;        // pixel lines of the source tensor
};

__global__ void __launch_bounds__(num_threads(8), 1)
conv_hshare_kernel(const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB, const GemmDesc d,
                   const HsParams hp, const int num_tiles) {
  using L = HsSmem;
  constexpr int NACC = 4;
  constexpr int ACC_COLS = 64;                    // accumulator stride in TMEM columns (48 used)
  constexpr uint32_t TMEM_COLS = NACC * ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (HS_SLABS + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * HS_SLABS + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * HS_SLABS + NACC + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * HS_SLABS + 2 * NACC);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < HS_SLABS; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, (uint32_t)(HS_NY * HS_W_TAP_BYTES));
      for (int t = 0; t < HS_NY; ++t) tma_load_2d(smem_base + L::W_OFF + t * HS_W_TAP_BYTES, &tmB, t * BK, 0, wfull_bar);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int g = (int)hp.fd_wt.div((uint32_t)tile), k = tile - g * hp.WT;
        mbar_wait(aempty_bar(slot), phase ^ 1);
        mbar_expect_tx(afull_bar(slot), HS_SLAB_BYTES);
        tma_load_3d(smem_base + L::A_OFF + slot * HS_SLAB_BYTES, &tmA3, 0, hp.col0 + HS_OUTW * k, 8 * g + hp.dy0, afull_bar(slot));
        if (++slot == HS_SLABS) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, HS_N);
      mbar_wait(wfull_bar, 0);
      const uint64_t wdesc0 = make_smem_desc(smem_base + L::W_OFF);
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it % NACC;
        const uint32_t acc_phase = (it / NACC) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
        mbar_wait(afull_bar(slot), phase);
        tcgen05_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_base + L::A_OFF + slot * HS_SLAB_BYTES);
#pragma unroll
        for (int yi = 0; yi < HS_NY; ++yi) {
          // vertical tap yi: the same box, yi lines of 16 pixels (= 2 swizzle atoms = 128 descriptor units) further down
          const uint64_t adesc = adesc0 + (uint64_t)(yi * 128);
          const uint64_t bdesc = wdesc0 + (uint64_t)(yi * (HS_W_TAP_BYTES >> 4));
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (yi | kk) != 0 ? 1u : 0u);
        }
        umma_commit(aempty_bar(slot));
        if (++slot == HS_SLABS) { slot = 0; phase ^= 1; }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ===================== epilogue: two groups of 4 warps take alternate tiles =====================
    const int q4 = warp & 3;                       // TMEM lane quarter: rows 32 q4 .. 32 q4 + 31 = source lines 2 q4, 2 q4 + 1
    const int part = (warp - 2) >> 2;
    const int row = q4 * 32 + lane;
    const int l = row >> 4, w = row & 15;
    const bool w_ok = w >= -hp.dx0 && w < -hp.dx0 + HS_OUTW;
    float* dz = reinterpret_cast<float*>(d.D);
    for (int it = part, tile = blockIdx.x + part * gridDim.x; tile < num_tiles; it += 2, tile += 2 * gridDim.x) {
      const int g = (int)hp.fd_wt.div((uint32_t)tile), k = tile - g * hp.WT;
      const int acc = it % NACC;
      const uint32_t acc_phase = (it / NACC) & 1;
      const long line = 8L * g + l;
      const int col = hp.col0 + HS_OUTW * k + w;
      long dest = -1;
      if (w_ok && line < hp.lines && col < d.src.P) dest = map_row(d.src, d.dst_kind, d.dst, line * d.src.P + col);
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * ACC_COLS);
      uint32_t r0[32], r1[16];
      tmem_ld<32>(taddr, r0);
      tmem_ld<16>(taddr + 32u, r1);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));   // the accumulator is in registers: release it before the shuffles
      float o[HS_NQ];
#pragma unroll
      for (int qq = 0; qq < HS_NQ; ++qq) o[qq] = 0.f;
#pragma unroll
      for (int j = 0; j < HS_NJ; ++j) {
#pragma unroll
        for (int qq = 0; qq < HS_NQ; ++qq) {
          const int n = j * HS_NQ + qq;
          const float v = __uint_as_float(n < 32 ? r0[n] : r1[n - 32]);
          // output pixel w needs D[(l, w + j + dx0), (j, q)]: pull it from the lane that holds that source pixel
          o[qq] += __shfl_sync(0xffffffffu, v, lane + j + hp.dx0);
        }
      }
      if (dest >= 0) {
        float4* op = reinterpret_cast<float4*>(dz + dest * d.ldd);
        op[0] = make_float4(o[0], o[1], o[2], o[3]);
        op[1] = make_float4(o[4], o[5], o[6], o[7]);
        op[2] = make_float4(o[8], o[9], o[10], o[11]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ===============================================================================================================
// 3x3 convolution 64 -> 64 (layer1's conv2 and its input gradient), "horizontal taps as N".
// The patch kernels above issue 36 tcgen05.mma of N = 64 per 128 pixels, and an N = 64 instruction reads 4 KB of A and 2 KB
// of B from shared memory for 32 cycles of math -- these layers are bound by shared-memory operand bandwidth (measured
// ~3 650 cycles per tile against 1 152 of tensor work).  With the three HORIZONTAL taps as the N dimension
//     D[u, (xi, co)] = sum_yi sum_ci  X[u + (yi + dy0) lines, ci] * W[co, (yi, xi), ci]         N = 3 x 64 = 192, K = 3 x 64
// a tile (source patch of 8 lines x 16 pixels, ONE 3-D TMA box of 10 lines) needs 12 instructions of N = 192 (4 KB + 6 KB
// per 96 cycles of math), and the epilogue finishes the horizontal sum in registers:
//     out[(l, w), co] = sum_xi D[(l, w + xi + dx0), (xi, co)]                                  warp shuffles inside a 16-lane line
// then bias / ReLU / sign bits / mask bits exactly as epilogue_role.  A tile yields 8 x 14 output pixels (lanes 1..14 of its
// 16 columns); tiles step by 14 columns.  Shared-memory operand traffic per 128 source pixels: 140 KB instead of 273 KB.
// The three partial products are summed in a different order than the 36-instruction accumulation (fp32, same terms).
// ===============================================================================================================
constexpr int H3_NJ = 3, H3_NY = 3, H3_CO = 64, H3_N = H3_NJ * H3_CO, H3_OUTW = 16 - (H3_NJ - 1);
constexpr int H3_SLAB_BYTES = (8 + H3_NY - 1) * 16 * 128;      // 20480
constexpr int H3_SLABS = 6;
constexpr int H3_W_TAP_BYTES = H3_N * BK * 2;                   // 24576
struct H3Smem {
  static constexpr int A_OFF = 0;
  static constexpr int W_OFF = H3_SLABS * H3_SLAB_BYTES;
  static constexpr int BAR_OFF = W_OFF + H3_NY * H3_W_TAP_BYTES;   // afull[S], aempty[S], tfull[2], tempty[2], wfull
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * H3_SLABS + 6) * 8;
  static constexpr int BIAS_OFF = TMEM_PTR_OFF + 16;
  static constexpr int TOTAL = BIAS_OFF + H3_CO * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};
static_assert(H3Smem::DYN_BYTES <= 232448, "conv3 hshare shared memory plan exceeds 227 KB");

__global__ void __launch_bounds__(num_threads(8), 1)
conv3_hshare_kernel(const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB, const GemmDesc d,
                    const HsParams hp, const int num_tiles) {
  using L = H3Smem;
  constexpr int NACC = 2;
  constexpr int ACC_COLS = 256;                   // accumulator stride in TMEM columns (192 used)
  constexpr uint32_t TMEM_COLS = NACC * ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (H3_SLABS + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * H3_SLABS + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * H3_SLABS + NACC + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * H3_SLABS + 2 * NACC);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < H3_SLABS; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    float* sb = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    for (int i = threadIdx.x; i < H3_CO; i += num_threads(8)) sb[i] = d.bias != nullptr ? d.bias[i] : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, (uint32_t)(H3_NY * H3_W_TAP_BYTES));
      for (int t = 0; t < H3_NY; ++t) tma_load_2d(smem_base + L::W_OFF + t * H3_W_TAP_BYTES, &tmB, t * BK, 0, wfull_bar);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int g = (int)hp.fd_wt.div((uint32_t)tile), k = tile - g * hp.WT;
        mbar_wait(aempty_bar(slot), phase ^ 1);
        mbar_expect_tx(afull_bar(slot), H3_SLAB_BYTES);
        tma_load_3d(smem_base + L::A_OFF + slot * H3_SLAB_BYTES, &tmA3, 0, hp.col0 + H3_OUTW * k, 8 * g + hp.dy0, afull_bar(slot));
        if (++slot == H3_SLABS) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, H3_N);
      mbar_wait(wfull_bar, 0);
      const uint64_t wdesc0 = make_smem_desc(smem_base + L::W_OFF);
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it % NACC;
        const uint32_t acc_phase = (it / NACC) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
        mbar_wait(afull_bar(slot), phase);
        tcgen05_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_base + L::A_OFF + slot * H3_SLAB_BYTES);
#pragma unroll
        for (int yi = 0; yi < H3_NY; ++yi) {
          // vertical tap yi: the same box, yi lines of 16 pixels (= 2 swizzle atoms = 128 descriptor units) further down
          const uint64_t adesc = adesc0 + (uint64_t)(yi * 128);
          const uint64_t bdesc = wdesc0 + (uint64_t)(yi * (H3_W_TAP_BYTES >> 4));
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (yi | kk) != 0 ? 1u : 0u);
        }
        umma_commit(aempty_bar(slot));
        if (++slot == H3_SLABS) { slot = 0; phase ^= 1; }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ===================== epilogue: two groups of 4 warps take alternate tiles =====================
    const float* sbias = reinterpret_cast<const float*>(smem + L::BIAS_OFF);
    const int q4 = warp & 3;                       // TMEM lane quarter: rows 32 q4 .. 32 q4 + 31 = source lines 2 q4, 2 q4 + 1
    const int part = (warp - 2) >> 2;
    const int row = q4 * 32 + lane;
    const int l = row >> 4, w = row & 15;
    const bool w_ok = w >= -hp.dx0 && w < -hp.dx0 + H3_OUTW;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.D);
    for (int it = part, tile = blockIdx.x + part * gridDim.x; tile < num_tiles; it += 2, tile += 2 * gridDim.x) {
      const int g = (int)hp.fd_wt.div((uint32_t)tile), k = tile - g * hp.WT;
      const int acc = it % NACC;
      const uint32_t acc_phase = (it / NACC) & 1;
      const long line = 8L * g + l;
      const int col = hp.col0 + H3_OUTW * k + w;
      long m = -1, dest = -1;
      if (w_ok && line < hp.lines && col < d.src.P) { m = line * d.src.P + col; dest = map_row(d.src, d.dst_kind, d.dst, m); }
      uint32_t mb0 = 0xFFFFFFFFu, mb1 = 0xFFFFFFFFu;
      if (dest >= 0 && d.mask_bits != nullptr) {
        mb0 = __ldg(d.mask_bits + bits_index(m, 0, d.ld_mb));
        mb1 = __ldg(d.mask_bits + bits_index(m, 1, d.ld_mb));
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * ACC_COLS);
      float o[H3_CO];
#pragma unroll
      for (int c = 0; c < H3_CO; ++c) o[c] = 0.f;
#pragma unroll
      for (int xi = 0; xi < H3_NJ; ++xi) {
        uint32_t r0[32], r1[32];
        tmem_ld<32>(taddr + (uint32_t)(xi * H3_CO), r0);
        tmem_ld<32>(taddr + (uint32_t)(xi * H3_CO + 32), r1);
        tmem_ld_wait();
        if (xi == H3_NJ - 1) {                       // the accumulator is in registers: release it before the last shuffles
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        // output pixel w needs D[(l, w + xi - 1), (xi, co)]: pull it from the lane that holds that source pixel (dx0 = -1 is
        // checked by the plan builder, so the centre tap needs no shuffle)
        if (xi == 1) {
#pragma unroll
          for (int c = 0; c < 32; ++c) { o[c] += __uint_as_float(r0[c]); o[32 + c] += __uint_as_float(r1[c]); }
        } else {
          const int src_lane = lane + xi - 1;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            o[c] += __shfl_sync(0xffffffffu, __uint_as_float(r0[c]), src_lane);
            o[32 + c] += __shfl_sync(0xffffffffu, __uint_as_float(r1[c]), src_lane);
          }
        }
      }
      if (dest >= 0) {
        uint32_t bo0 = 0u, bo1 = 0u;
#pragma unroll
        for (int c = 0; c < H3_CO; ++c) {
          float v = o[c] + sbias[c];
          if (d.relu) v = fmaxf(v, 0.f);
          const uint32_t keep = ((c < 32 ? mb0 : mb1) >> (c & 31)) & 1u;
          v = keep ? v : 0.f;
          if (v > 0.f) { if (c < 32) bo0 |= 1u << c; else bo1 |= 1u << (c - 32); }
          o[c] = v;
        }
        uint32_t pk[H3_CO / 2];
#pragma unroll
        for (int c = 0; c < H3_CO / 2; ++c) pk[c] = pack_bf16x2(o[2 * c], o[2 * c + 1]);
        __nv_bfloat16* op = out + dest * d.ldd;
#pragma unroll
        for (int c = 0; c < H3_CO / 16; ++c) stg256(op + 16 * c, pk + 8 * c);
        if (d.D_bits != nullptr) {
          d.D_bits[bits_index(dest, 0, d.ld_db)] = bo0;
          d.D_bits[bits_index(dest, 1, d.ld_db)] = bo1;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Wh3[(xi * 64 + n), yi * 64 + c] = Wt[n, tap(yi, xi) * 64 + c]
__global__ void repack_h3_weights_kernel(const __nv_bfloat16* __restrict__ wt, __nv_bfloat16* __restrict__ wh, int ntaps,
                                         int t00, int t01, int t02, int t10, int t11, int t12, int t20, int t21, int t22) {
  const int taps[3][3] = {{t00, t01, t02}, {t10, t11, t12}, {t20, t21, t22}};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H3_N * H3_NY * 64; i += gridDim.x * blockDim.x) {
    const int rowi = i / (H3_NY * 64), kcol = i - rowi * (H3_NY * 64);
    const int xi = rowi / 64, n = rowi - xi * 64, yi = kcol / 64, c = kcol - yi * 64;
    wh[i] = wt[(long)n * ntaps * 64 + taps[yi][xi] * 64 + c];
  }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace

// General 2-D map (exported for gemm_tc32.cu): `dtype` 0 = bf16, 1 = fp32; row_elems = elements between rows (0 = dense)
int make_tensor_map_2d(CUtensorMap* map, const void* base, int dtype, uint64_t inner, uint64_t rows, uint32_t box_inner,
                       uint32_t box_rows, int swizzle_bytes, uint64_t row_elems) {
  PFN_encodeTiled enc = get_encode_fn();
  RGIE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  const uint64_t esz = dtype == 1 ? 4 : 2;
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {(row_elems ? row_elems : inner) * esz};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = enc(map, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

namespace {

int make_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows,
                CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, uint64_t row_elems = 0) {
  PFN_encodeTiled enc = get_encode_fn();
  RGIE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {(row_elems ? row_elems : inner) * 2};     // row_elems < inner: overlapped rows (GemmDesc::a_ld)
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

// A as [C = 64, P, lines] with a box of 8 pixels x box_lines lines (one 128-byte row per pixel, SWIZZLE_128B)
// pix_elems: elements between consecutive pixels in memory (64 = dense; 16 = overlapped 4-pixel windows, GemmDesc::a_ld)
int make_map_3d(CUtensorMap* map, const void* base, uint64_t P, uint64_t lines, uint32_t box_lines, uint64_t pix_elems = 64) {
  PFN_encodeTiled enc = get_encode_fn();
  RGIE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[3] = {64, P, lines};
  cuuint64_t gstride[2] = {pix_elems * 2, P * pix_elems * 2};
  cuuint32_t box[3] = {64, 8, box_lines};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (3-D) failed with CUresult " + std::to_string((int)r));
  return 0;
}

template <int BN, int NY, int NX>
int run_patch(const GemmPlanSm100& p, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(gemm_patch_kernel<BN, NY, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PatchSmem::DYN_BYTES));
    attr_once.done();
  }
  PatchTaps tp;
  tp.ny = p.patch_ny; tp.nx = p.patch_nx; tp.dy0 = p.patch_dy0; tp.dx0 = p.patch_dx0;
  for (int y = 0; y < 4; ++y)
    for (int x = 0; x < 4; ++x) tp.tap[y][x] = p.patch_tap[y][x];
  PatchMap pm;
  pm.enabled = 1; pm.P = p.d.src.P; pm.WT = p.patch_wt; pm.fd_wt = make_fastdiv((uint32_t)p.patch_wt);
  gemm_patch_kernel<BN, NY, NX><<<p.grid, num_threads(8), PatchSmem::DYN_BYTES, st>>>(p.tmA, p.tmB, p.d, tp, pm, p.num_m_tiles);
  RGIE_LAUNCH_OK();
  return 0;
}

template <int BN, int NY, int NX>
int run_patch_2cta(const GemmPlanSm100& p, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(gemm_patch_2cta_kernel<BN, NY, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PatchSmem::DYN_BYTES));
    attr_once.done();
  }
  PatchTaps tp;
  tp.ny = p.patch_ny; tp.nx = p.patch_nx; tp.dy0 = p.patch_dy0; tp.dx0 = p.patch_dx0;
  for (int y = 0; y < 4; ++y)
    for (int x = 0; x < 4; ++x) tp.tap[y][x] = p.patch_tap[y][x];
  PatchMap pm;
  pm.enabled = 1; pm.P = p.d.src.P; pm.WT = p.patch_wt; pm.fd_wt = make_fastdiv((uint32_t)p.patch_wt);
  gemm_patch_2cta_kernel<BN, NY, NX><<<p.grid, num_threads(8), PatchSmem::DYN_BYTES, st>>>(p.tmA, p.tmB, p.d, tp, pm, p.num_m_tiles);
  RGIE_LAUNCH_OK();
  return 0;
}

int run_b2b(const GemmPlanSm100& p, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(gemm_b2b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemB2b::DYN_BYTES));
    attr_once.done();
  }
  B2bParams q;
  q.bias2 = p.d2.bias; q.D2 = p.d2.D; q.ldd2 = p.d2.ldd; q.relu2 = p.d2.relu;
  q.mask_bits2 = p.d2.mask_bits; q.ld_mb2 = p.d2.ld_mb; q.D2_bits = p.d2.D_bits; q.ld_db2 = p.d2.ld_db;
  gemm_b2b_kernel<<<p.grid, 640, SmemB2b::DYN_BYTES, st>>>(p.tmA, p.tmA2, p.tmB, p.tmW2, p.tmD, p.tmR, p.d, q, p.num_m_tiles);
  RGIE_LAUNCH_OK();
  return 0;
}

template <int BN, int STAGES>
int run_2cta(const GemmPlanSm100& p, cudaStream_t st) {
  using L = Smem2<BN, STAGES>;
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(gemm_sm100_2cta_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::DYN_BYTES));
    attr_once.done();
  }
  gemm_sm100_2cta_kernel<BN, STAGES><<<p.grid, 640, L::DYN_BYTES, st>>>(
      p.tmA, p.tmA2, p.tmB, p.d, p.num_m_tiles, p.num_n_tiles, make_fastdiv((uint32_t)p.num_n_tiles));
  RGIE_LAUNCH_OK();
  return 0;
}

template <int BN, int STAGES, int EPI, int NEW>
int run_impl(const GemmPlanSm100& p, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES, EPI>;
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(gemm_sm100_kernel<BN, STAGES, EPI, NEW>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
    attr_once.done();
  }
  gemm_sm100_kernel<BN, STAGES, EPI, NEW><<<p.grid, num_threads(NEW), L::DYN_BYTES, st>>>(
      p.tmA, p.tmA2, p.tmB, p.tmD, p.tmR, p.d, p.num_m_tiles, p.num_n_tiles, make_fastdiv((uint32_t)p.num_n_tiles));
  RGIE_LAUNCH_OK();
  return 0;
}

}  // namespace

int gemm_sm100_num_sms() {
  static int cache[64] = {0};                  // per device ordinal (a benign race: every writer stores the same value)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

int build_gemm_sm100(const GemmDesc& d, GemmPlanSm100* p) {
  RGIE_CHECK(d.Cin % BK == 0, "gemm_sm100: Cin must be a multiple of 64");
  RGIE_CHECK(d.ntaps >= 1 && d.ntaps <= kMaxTaps, "gemm_sm100: ntaps out of range");
  RGIE_CHECK(d.Cout % 16 == 0 && d.Cout <= MAX_BIAS, "gemm_sm100: Cout must be a multiple of 16 and <= 2048");
  RGIE_CHECK(d.a_rows < (1L << 31), "gemm_sm100: too many A rows for a TMA coordinate");
  int bn = d.Cout >= 256 ? 256 : (d.Cout >= 128 ? 128 : (d.Cout >= 64 ? 64 : 16));
  RGIE_CHECK(d.Cout % bn == 0 && d.n_pad % bn == 0, "gemm_sm100: Cout/n_pad must be a multiple of the N tile");
  RGIE_CHECK(d.A2 == nullptr || (d.Cin2 % BK == 0 && d.a2_rows < (1L << 31)), "gemm_sm100: second operand: Cin2 % 64, rows");
  RGIE_CHECK((d.mask_bits == nullptr && d.D_bits == nullptr) || bn >= 64, "gemm_sm100: bit masks need Cout >= 64");
  RGIE_CHECK(d.ldd % 16 == 0 && (d.res == nullptr || d.ld_res % 16 == 0) && (d.mask == nullptr || d.ld_mask % 16 == 0),
             "gemm_sm100: leading dimensions must be multiples of 16 elements (32-byte accesses)");
  p->d = d;
  p->bn = bn;
  p->patch = 0;
  p->patch_2cta = 0;
  p->special = 0;
  p->b2b = 0;
  p->pool = 0;
  // ---- patch-tile variant: Cin = 64, one N tile, single-plane source whose rows are whole pixel lines, taps on a
  //      (dy, dx) grid of at most 4 x 4 with |dx| well below the pitch
  //      Measured on B200 (320 crops): 3x3 layer1 0.52 -> 0.41 ms, conv1 forward 1.08 -> 0.76 ms, conv1 input gradient
  //      3.6 -> 2.5 ms.  (A first version with a generic tap loop in the issuing thread was SLOWER than the flat kernel:
  //      for narrow N tiles the single MMA-issuing thread is the critical resource.)  RGIE_GEMM_PATCH=0 disables it.
  static const int env_patch = getenv("RGIE_GEMM_PATCH") ? atoi(getenv("RGIE_GEMM_PATCH")) : 1;
  if (env_patch && d.Cin == 64 && bn <= 64 && d.Cout == bn && d.ntaps >= 3 && d.A2 == nullptr && d.src.planes == 1 &&
      d.m_begin == 0 && d.m_end == d.a_rows && d.a_rows == d.src.rows() && d.src.P >= 16 && d.mask == nullptr) {
    const int P = d.src.P;
    int dy[kMaxTaps], dx[kMaxTaps], dymin = 1 << 30, dymax = -(1 << 30), dxmin = 1 << 30, dxmax = -(1 << 30);
    for (int t = 0; t < d.ntaps; ++t) {
      const long off = d.row_off[t];
      long y = (off >= 0 ? off + P / 2 : off - P / 2) / P;     // nearest line
      dy[t] = (int)y; dx[t] = (int)(off - y * P);
      dymin = dy[t] < dymin ? dy[t] : dymin; dymax = dy[t] > dymax ? dy[t] : dymax;
      dxmin = dx[t] < dxmin ? dx[t] : dxmin; dxmax = dx[t] > dxmax ? dx[t] : dxmax;
    }
    const int ny = dymax - dymin + 1, nx = dxmax - dxmin + 1;
    if (ny <= 4 && nx <= 4 && dxmin >= -4 && dxmax <= 4 && (long)d.ntaps * bn * 128 <= PATCH_W_BYTES) {
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) p->patch_tap[y][x] = -1;
      bool ok = true;
      for (int t = 0; t < d.ntaps; ++t) {
        int& slot = p->patch_tap[dy[t] - dymin][dx[t] - dxmin];
        if (slot >= 0) ok = false;
        slot = t;
      }
      int variant = 0;
      if (ok && d.ntaps == ny * nx) {
        if (bn == 64 && ny == 3 && nx == 3) variant = 1;
        else if (bn == 64 && ny == 4 && nx == 1) variant = 2;
        else if (bn == 16 && ny == 4 && nx == 4) variant = 3;
      }
      if (variant) {
        p->patch = variant;
        p->patch_ny = ny; p->patch_nx = nx; p->patch_dy0 = dymin; p->patch_dx0 = dxmin;
        const long lines = d.a_rows / P;
        p->patch_wt = (P + 7) / 8;
        p->num_m_tiles = (int)(((lines + 15) / 16) * p->patch_wt);
        p->num_n_tiles = 1;
        int sms = gemm_sm100_num_sms();
        p->grid = p->num_m_tiles < sms ? p->num_m_tiles : sms;
        p->epi = 0;
        // CTA pairs (gemm_patch_2cta_kernel) for the 3x3 variant (36 MMAs per tile).  Measured (B200, 320 crops, same box):
        // layer1 3x3 forward 0.468 -> 0.414 ms, input gradient 0.476 -> 0.434 ms.  conv1 forward (16 MMAs and ONE box per
        // tile) LOSES with pairs, 0.88 -> 1.21 ms: the cross-SM barrier round trip per tile is no longer hidden.
        // RGIE_PATCH_2CTA=0 keeps one CTA per tile, =2 also pairs conv1.
        static const int env_p2 = getenv("RGIE_PATCH_2CTA") ? atoi(getenv("RGIE_PATCH_2CTA")) : 1;
        p->patch_2cta = (env_p2 && bn == 64 && (variant == 1 || (variant == 2 && env_p2 >= 2)) && p->num_m_tiles >= 2 && sms >= 2) ? 1 : 0;
        if (p->patch_2cta) {
          p->num_m_tiles = (p->num_m_tiles + 1) & ~1;          // a phantom last tile keeps the pairs whole
          const int even_sms = sms & ~1;
          p->grid = p->num_m_tiles < even_sms ? p->num_m_tiles : even_sms;
        }
        p->tmA2 = p->tmA; p->tmD = p->tmA; p->tmR = p->tmA;
        int rc = make_map_3d(&p->tmA, d.A, (uint64_t)P, (uint64_t)lines, (uint32_t)(16 + ny - 1),
                             (uint64_t)(d.a_ld ? d.a_ld : 64));
        if (rc) return rc;
        p->tmA2 = p->tmA; p->tmD = p->tmA; p->tmR = p->tmA;
        return make_map_2d(&p->tmB, d.Wt, (uint64_t)d.ntaps * d.Cin, (uint64_t)d.n_pad, BK,
                           (uint32_t)(p->patch_2cta ? bn / 2 : bn));   // CTA pairs: each CTA loads half of the weight rows
      }
    }
  }
  long M = d.m_end - d.m_begin;
  p->num_m_tiles = ceil_div(M, BM);
  p->num_n_tiles = d.Cout / bn;
  long tiles = (long)p->num_m_tiles * p->num_n_tiles;
  int sms = gemm_sm100_num_sms();
  p->grid = (int)(tiles < sms ? tiles : sms);
  if (p->grid < 1) p->grid = 1;
  // Epilogue variant (measured per shape class on B200, profiles/README.md).  256-wide bf16 tiles take the 16-warp lean role
  // (setmaxnreg: producer warpgroup 32 registers, epilogue warpgroups 112): -16.  Where the destination rows are the source
  // rows (DST_SAME) and the launch is HBM-bound (contraction < 768) the output leaves through shared-memory boxes and TMA
  // stores: -17 (ncu: the L1TEX LSU data pipe was 73-82 % busy with the 32-byte row-per-thread stores; 0.61 -> 0.53,
  // 0.89 -> 0.77, 0.50 -> 0.41, 0.29 -> 0.23 ms on the layer1-3 expansions).  -18: the same launches with the residual
  // operand ALSO moved by TMA (epilogue_lean_dma_role: in-place half boxes served by two DMA threads); measured (320 crops,
  // same box): conv1 input gradient + skip of layer1 0.875 -> 0.773 ms, of layer2 0.455 -> 0.389 ms, conv3 + skip 0.768 ->
  // 0.746 / 0.412 -> 0.389 ms, GEMM family 30.7 -> 30.0 ms per micro-batch, step 74.7 -> 73.6 ms.  Everything else takes the
  // classic 8-warp row-per-thread role: 0.
  // Switches for A/B runs: RGIE_GEMM_EPI = 0 (classic everywhere) / -16 (lean without the TMA-store path);
  // RGIE_LEAN_DMA = 0 (keep -17) / 1 (default: -18 for the ops that have a residual) / 2 (-18 for every -17 op).
  // Variants that were measured and LOST are no longer in the file (numbers in profiles/README.md): 8-warp TMA-store
  // epilogues, two M tiles per weight load inside one CTA, CTA pairs for 128-wide tiles.
  static const int env_epi = getenv("RGIE_GEMM_EPI") ? atoi(getenv("RGIE_GEMM_EPI")) : -1;
  const int ktot = d.ntaps * d.Cin + (d.A2 ? d.Cin2 : 0);
  const bool lean_ok = bn == 256 && !d.d_fp32 && d.mask == nullptr && env_epi != 0;
  p->epi = lean_ok ? -16 : 0;
  if (lean_ok && env_epi != -16 && d.dst_kind == DST_SAME && ktot < 768) p->epi = -17;
  static const int env_dma = getenv("RGIE_LEAN_DMA") ? atoi(getenv("RGIE_LEAN_DMA")) : 1;
  if (p->epi == -17 && env_dma > 0 && (d.res != nullptr || env_dma >= 2) && d.ldd % 32 == 0 && (d.res == nullptr || d.ld_res % 32 == 0))
    p->epi = -18;
  // CTA pairs (cta_group::2, M = 256 across the two SMs of a TPC) for the tensor-bound 256-wide tiles: contraction >= 768.
  // Measured (320 crops, same box, clock drift removed): K >= 1024 layers -6 % (1.36-1.47 PFLOP/s); K <= 640 layers are
  // epilogue-bound and LOSE 7-28 % (the leader's MMA waits for the epilogues of BOTH CTAs across the TPC), so they stay on
  // the single-CTA kernel.  RGIE_GEMM_2CTA = 0 switches pairs off, = K sets the contraction threshold.
  static const int env_2cta = getenv("RGIE_GEMM_2CTA") ? atoi(getenv("RGIE_GEMM_2CTA")) : 768;
  if (p->epi == -16 && env_2cta > 0 && ktot >= env_2cta && p->num_m_tiles >= 2) {
    p->epi = -32;
    const long pairs = (long)((p->num_m_tiles + 1) / 2) * p->num_n_tiles;
    const long ctas = 2 * pairs < (long)(sms & ~1) ? 2 * pairs : (long)(sms & ~1);
    p->grid = (int)ctas;
  }
  p->tmD = p->tmA; p->tmR = p->tmA;
  int rc = 0;
  if (p->epi == -17) {
    rc = make_map_2d(&p->tmD, d.D, (uint64_t)d.ldd, (uint64_t)d.m_end, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (p->epi == -18) {
    rc = make_map_2d(&p->tmD, d.D, (uint64_t)d.ldd, (uint64_t)d.m_end, 32, BM, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    if (d.res != nullptr) {
      // rows at or beyond res_rows are out of range for the map: zero fill = "no residual" there
      const long rr = d.res_rows < d.m_end ? d.res_rows : d.m_end;
      rc = make_map_2d(&p->tmR, d.res, (uint64_t)d.ld_res, (uint64_t)rr, 32, BM, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
  }
  rc = make_map_2d(&p->tmA, d.A, (uint64_t)d.Cin, (uint64_t)d.a_rows, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B, (uint64_t)d.a_ld);
  if (rc) return rc;
  if (d.A2 != nullptr) rc = make_map_2d(&p->tmA2, d.A2, (uint64_t)d.Cin2, (uint64_t)d.a2_rows, BK, BM);
  else p->tmA2 = p->tmA;
  if (rc) return rc;
  return make_map_2d(&p->tmB, d.Wt, (uint64_t)d.ntaps * d.Cin + (d.A2 ? d.Cin2 : 0), (uint64_t)d.n_pad, BK,
                     (uint32_t)(p->epi == -32 ? bn / 2 : bn));   // CTA pairs: each CTA loads half of the weight rows
}

// Plan for conv_hshare_kernel.  d: the 16-tap input-gradient descriptor of the flat formulation (geometry, source,
// destination); Wh: [48, 4 * 64] bf16, row j * 12 + q, column yi * 64 + co; taps (dy0 + yi, dx0 + j).
// ---- back-to-back fusion: d1 = a 256-wide single-tap op whose output rows are its own rows (DST_SAME), d2 = the single-tap
//      256 -> 64 op that reads exactly that output over the same rows
bool gemm_b2b_eligible(const GemmDesc& d1, const GemmDesc& d2) {
  auto same_geom = [](const Geom& a, const Geom& b) {
    return a.planes == b.planes && a.n_img == b.n_img && a.H == b.H && a.W == b.W && a.pad_t == b.pad_t && a.pad_l == b.pad_l &&
           a.P == b.P && a.S == b.S;
  };
  if (d1.Cout != 256 || d1.n_pad != 256 || d1.d_fp32 || d1.mask != nullptr || d1.ntaps != 1 || d1.row_off[0] != 0) return false;
  if (d1.dst_kind != DST_SAME || d1.ldd != 256 || d1.Cin % BK != 0 || (d1.A2 != nullptr && d1.Cin2 % BK != 0)) return false;
  if (d2.A != d1.D || d2.Cin != 256 || d2.ntaps != 1 || d2.row_off[0] != 0 || d2.A2 != nullptr || d2.res != nullptr) return false;
  if (d2.mask != nullptr || d2.d_fp32 || d2.Cout != B2B_N2 || d2.n_pad != B2B_N2 || d2.dst_kind != DST_SAME) return false;
  if (d2.m_begin != d1.m_begin || d2.m_end != d1.m_end || d2.ldd % 16 != 0) return false;
  if (!same_geom(d1.src, d2.src) || d1.m_end <= d1.m_begin) return false;
  if (d1.res != nullptr && d1.ld_res % 16 != 0) return false;
  return true;
}

int build_gemm_b2b_sm100(const GemmDesc& d1, const GemmDesc& d2, GemmPlanSm100* p) {
  RGIE_CHECK(gemm_b2b_eligible(d1, d2), "gemm_b2b: the two ops cannot be fused");
  RGIE_CHECK(d1.a_rows < (1L << 31) && (d1.A2 == nullptr || d1.a2_rows < (1L << 31)), "gemm_b2b: too many rows for a TMA coordinate");
  p->d = d1; p->d2 = d2;
  p->bn = 256; p->patch = 0; p->special = 0; p->epi = 0; p->b2b = 1; p->pool = 0;
  p->num_m_tiles = ceil_div(d1.m_end - d1.m_begin, (long)BM);
  p->num_n_tiles = 1;
  const int sms = gemm_sm100_num_sms();
  p->grid = p->num_m_tiles < sms ? p->num_m_tiles : sms;
  p->tmD = p->tmA; p->tmR = p->tmA;
  int rc = make_map_2d(&p->tmA, d1.A, (uint64_t)d1.Cin, (uint64_t)d1.a_rows, BK, BM);
  if (rc) return rc;
  p->tmA2 = p->tmA;
  if (d1.A2 != nullptr) rc = make_map_2d(&p->tmA2, d1.A2, (uint64_t)d1.Cin2, (uint64_t)d1.a2_rows, BK, BM);
  if (rc) return rc;
  rc = make_map_2d(&p->tmB, d1.Wt, (uint64_t)d1.ntaps * d1.Cin + (d1.A2 ? d1.Cin2 : 0), (uint64_t)d1.n_pad, BK, 256);
  if (rc) return rc;
  rc = make_map_2d(&p->tmD, d1.D, (uint64_t)d1.ldd, (uint64_t)d1.m_end, 32, BM, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  if (d1.res != nullptr) {
    const long rr = d1.res_rows < d1.m_end ? d1.res_rows : d1.m_end;
    rc = make_map_2d(&p->tmR, d1.res, (uint64_t)d1.ld_res, (uint64_t)rr, 32, BM, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  return make_map_2d(&p->tmW2, d2.Wt, 256, (uint64_t)B2B_N2, BK, (uint32_t)B2B_N2);
}

int build_conv_hshare_sm100(const GemmDesc& d, const void* Wh, int dy0, int dx0, GemmPlanSm100* p) {
  RGIE_CHECK(d.Cin == 64 && d.src.planes == 1 && d.d_fp32 && d.ldd >= HS_NQ && d.ldd % 4 == 0 && d.a_rows == d.src.rows(),
             "conv_hshare: unsupported descriptor");
  RGIE_CHECK(dx0 <= 0 && dx0 + HS_NJ - 1 >= 0 && d.src.pad_l + dx0 >= 0, "conv_hshare: tap range");
  p->d = d;
  p->bn = HS_N;
  p->patch = 0; p->epi = 0; p->b2b = 0; p->pool = 0;
  p->special = 1;
  const int P = d.src.P;
  const long lines = d.a_rows / P;
  p->hs_wt = ceil_div(d.src.W, HS_OUTW);
  p->hs_dy0 = dy0; p->hs_dx0 = dx0; p->hs_col0 = d.src.pad_l + dx0; p->hs_lines = lines;
  p->num_m_tiles = (int)(((lines + 7) / 8) * p->hs_wt);
  p->num_n_tiles = 1;
  const int sms = gemm_sm100_num_sms();
  p->grid = p->num_m_tiles < sms ? p->num_m_tiles : sms;
  // A as [C = 64, P, lines] with a box of 16 pixels x 11 lines
  PFN_encodeTiled enc = get_encode_fn();
  RGIE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[3] = {64, (cuuint64_t)P, (cuuint64_t)lines};
  cuuint64_t gstride[2] = {64 * 2, (cuuint64_t)P * 64 * 2};
  cuuint32_t box[3] = {64, 16, 8 + HS_NY - 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&p->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d.A), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (hshare) failed with CUresult " + std::to_string((int)r));
  p->tmA2 = p->tmA; p->tmD = p->tmA; p->tmR = p->tmA;
  return make_map_2d(&p->tmB, Wh, (uint64_t)HS_NY * 64, (uint64_t)HS_N, BK, (uint32_t)HS_N);
}

static int run_conv_hshare(const GemmPlanSm100& p, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(conv_hshare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HsSmem::DYN_BYTES));
    attr_once.done();
  }
  HsParams hp;
  hp.WT = p.hs_wt; hp.fd_wt = make_fastdiv((uint32_t)p.hs_wt); hp.dy0 = p.hs_dy0; hp.dx0 = p.hs_dx0; hp.col0 = p.hs_col0;
  hp.lines = p.hs_lines;
  conv_hshare_kernel<<<p.grid, num_threads(8), HsSmem::DYN_BYTES, st>>>(p.tmA, p.tmB, p.d, hp, p.num_m_tiles);
  RGIE_LAUNCH_OK();
  return 0;
}

// Plan for conv3_hshare_kernel from the descriptor of a 3x3 64 -> 64 op (any 3 x 3 tap grid: forward or input gradient).
// `wh_buf`: device buffer of 192 * 192 bf16 owned by the caller; it receives the re-ordered weights.
int build_conv3_hshare_sm100(const GemmDesc& d, void* wh_buf, GemmPlanSm100* p) {
  if (int rc = build_gemm_sm100(d, p)) return rc;
  RGIE_CHECK(p->patch == 1 && p->patch_ny == 3 && p->patch_nx == 3, "conv3_hshare: not a 3x3 64 -> 64 patch op");
  RGIE_CHECK(!d.d_fp32 && d.mask == nullptr && d.res == nullptr && d.A2 == nullptr && d.ldd % 16 == 0 && d.Cout == 64,
             "conv3_hshare: unsupported epilogue operands");
  RGIE_CHECK(p->patch_dx0 == -1 && d.src.pad_l + p->patch_dx0 >= 0, "conv3_hshare: the horizontal taps must be -1, 0, +1");
  const int (*t)[4] = p->patch_tap;
  repack_h3_weights_kernel<<<64, 256>>>(reinterpret_cast<const __nv_bfloat16*>(d.Wt), reinterpret_cast<__nv_bfloat16*>(wh_buf), d.ntaps,
                                        t[0][0], t[0][1], t[0][2], t[1][0], t[1][1], t[1][2], t[2][0], t[2][1], t[2][2]);
  RGIE_LAUNCH_OK();
  RGIE_CUDA_OK(cudaDeviceSynchronize());
  const int dy0 = p->patch_dy0, dx0 = p->patch_dx0;
  p->patch = 0; p->patch_2cta = 0; p->epi = 0; p->b2b = 0; p->pool = 0;
  p->special = 2;
  p->bn = H3_N;
  const int P = d.src.P;
  const long lines = d.a_rows / P;
  p->hs_wt = ceil_div(d.src.W, H3_OUTW);
  p->hs_dy0 = dy0; p->hs_dx0 = dx0; p->hs_col0 = d.src.pad_l + dx0; p->hs_lines = lines;
  p->num_m_tiles = (int)(((lines + 7) / 8) * p->hs_wt);
  p->num_n_tiles = 1;
  const int sms = gemm_sm100_num_sms();
  p->grid = p->num_m_tiles < sms ? p->num_m_tiles : sms;
  PFN_encodeTiled enc = get_encode_fn();
  RGIE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[3] = {64, (cuuint64_t)P, (cuuint64_t)lines};
  cuuint64_t gstride[2] = {64 * 2, (cuuint64_t)P * 64 * 2};
  cuuint32_t box[3] = {64, 16, 8 + H3_NY - 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&p->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d.A), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (conv3 hshare) failed with CUresult " + std::to_string((int)r));
  p->tmA2 = p->tmA; p->tmD = p->tmA; p->tmR = p->tmA;
  return make_map_2d(&p->tmB, wh_buf, (uint64_t)H3_NY * 64, (uint64_t)H3_N, BK, (uint32_t)H3_N);
}

static int run_conv3_hshare(const GemmPlanSm100& p, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(conv3_hshare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H3Smem::DYN_BYTES));
    attr_once.done();
  }
  HsParams hp;
  hp.WT = p.hs_wt; hp.fd_wt = make_fastdiv((uint32_t)p.hs_wt); hp.dy0 = p.hs_dy0; hp.dx0 = p.hs_dx0; hp.col0 = p.hs_col0;
  hp.lines = p.hs_lines;
  conv3_hshare_kernel<<<p.grid, num_threads(8), H3Smem::DYN_BYTES, st>>>(p.tmA, p.tmB, p.d, hp, p.num_m_tiles);
  RGIE_LAUNCH_OK();
  return 0;
}

int run_conv1_pool(const GemmPlanSm100& p, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(gemm_conv1_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PatchSmem::DYN_BYTES));
    attr_once.done();
  }
  PatchTaps tp;
  tp.ny = p.patch_ny; tp.nx = p.patch_nx; tp.dy0 = p.patch_dy0; tp.dx0 = p.patch_dx0;
  for (int y = 0; y < 4; ++y)
    for (int x = 0; x < 4; ++x) tp.tap[y][x] = p.patch_tap[y][x];
  const int bands = p.sp.n_img * p.sp.TY;
  const int sms = gemm_sm100_num_sms();
  gemm_conv1_pool_kernel<<<bands < sms ? bands : sms, num_threads(8), PatchSmem::DYN_BYTES, st>>>(p.tmA, p.tmB, p.d, tp, p.sp, bands);
  RGIE_LAUNCH_OK();
  return 0;
}

int build_conv1_pool_sm100(const GemmDesc& conv1, void* P1, const Geom& g1, uint8_t* arg, int Hp, GemmPlanSm100* p) {
  if (int rc = build_gemm_sm100(conv1, p)) return rc;
  RGIE_CHECK(p->patch == 2 && !p->patch_2cta && p->patch_ny == 4 && p->patch_nx == 1, "conv1 + pool: conv1 must be the 4-tap patch variant");
  RGIE_CHECK(conv1.relu && conv1.Cout == 64 && conv1.ntaps == 4 && conv1.src.planes == 1 && conv1.src.H == conv1.src.W &&
             conv1.src.H == 2 * Hp && conv1.src.W % 8 == 0, "conv1 + pool: unexpected stem geometry");
  RGIE_CHECK(g1.planes == 1 && g1.n_img == conv1.src.n_img && g1.H == Hp && g1.W == Hp, "conv1 + pool: pooled geometry");
  StemPoolParams& sp = p->sp;
  sp.n_img = conv1.src.n_img; sp.H0 = conv1.src.H; sp.Hp = Hp;
  sp.lines_per_img = conv1.src.S / conv1.src.P;
  sp.TY = ceil_div(Hp, 7); sp.TX = conv1.src.W / 8;     // tiles per band: 8 conv columns = 4 pooled columns each
  RGIE_CHECK((long)sp.n_img * sp.TY < (1L << 31), "conv1 + pool: too many bands");
  sp.fd_img = make_fastdiv((uint32_t)sp.TY);           // band -> image
  sp.fd_tx = make_fastdiv((uint32_t)sp.TX);
  sp.P1 = P1; sp.g1 = g1; sp.arg = arg;
  p->pool = 1;
  return 0;
}

int run_gemm_sm100(const GemmPlanSm100& p, cudaStream_t st) {
  if (p.pool == 1) return run_conv1_pool(p, st);
  if (p.b2b == 1) return run_b2b(p, st);
  if (p.d.m_end <= p.d.m_begin) return 0;
  if (p.special == 1) return run_conv_hshare(p, st);
  if (p.special == 2) return run_conv3_hshare(p, st);
  if (p.patch == 1) return p.patch_2cta ? run_patch_2cta<64, 3, 3>(p, st) : run_patch<64, 3, 3>(p, st);     // 3x3, 64 -> 64 (layer1)
  if (p.patch == 2) return p.patch_2cta ? run_patch_2cta<64, 4, 1>(p, st) : run_patch<64, 4, 1>(p, st);     // conv1 forward: 4 vertical taps
  if (p.patch == 3) return run_patch<16, 4, 4>(p, st);     // conv1 input gradient: 4 x 4 taps, 64 -> 16
  switch (p.bn) {
    case 256:
      switch (p.epi) {
        case -16: return run_impl<256, 4, 0, 16>(p, st);
        case -17: return run_impl<256, 3, 8, 16>(p, st);
        case -18: return run_impl<256, 3, 9, 16>(p, st);
        case -32: return run_2cta<256, 6>(p, st);
        default: return run_impl<256, 4, 0, 8>(p, st);
      }
    case 128: return run_impl<128, 6, 0, 8>(p, st);
    case 64: return run_impl<64, 8, 0, 8>(p, st);
    case 16: return run_impl<16, 8, 0, 8>(p, st);
  }
  return fail("gemm_sm100: unsupported N tile");
}

int launch_gemm_sm100(const GemmDesc& d, cudaStream_t st) {
  GemmPlanSm100 p;
  int rc = build_gemm_sm100(d, &p);
  if (rc) return rc;
  return run_gemm_sm100(p, st);
}

}  // namespace rgie

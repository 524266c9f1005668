// fp32-ACCURATE row-shifted GEMM on the tcgen05 tensor cores (sm_100a): the parity mode of the regressor and of the MiDU
// head without leaving the tensor pipe.
//
//   D[map(m), n] = epi( sum_t sum_c A[m + row_off[t], c] * W[n, t*Cin + c] + sum_c A2[m, c] * W[n, ntaps*Cin + c] )     A, W, D: fp32
//
// Every fp32 operand is split EXACTLY into three bf16 pieces, x = hi + mid + lo (8 + 8 + 8 significand bits), and the
// product is formed from the six partial products whose weight is above 2^-24 of the full product,
//   x*w ~ hi*whi + hi*wmid + mid*whi + mid*wmid + hi*wlo + lo*whi,
// each a bf16 x bf16 tcgen05.mma (kind::f16) accumulating in fp32 in tensor memory: the result carries fp32-level error
// (the dropped terms mid*lo, lo*mid, lo*lo are below 2^-24 relative) at 6 tensor instructions per fp32 product block.
// Weights are split once on the host (three K-major bf16 planes); activations stay fp32 in HBM and are split ON CHIP:
//   warp 0      TMA producer: per k-block of 64 the fp32 A tile as two 128 x 32 boxes (raw, SWIZZLE_128B) + three weight
//               plane boxes (bf16, SWIZZLE_128B)
//   warps 2..5  converters: thread r splits row r of the raw tile into the three bf16 operand tiles (SWIZZLE_128B K-major)
//   warp 1      one thread issues the 6 x 4 tcgen05.mma of the k-block (M = 128, N = 64, K = 16)
//   warps 6..9  epilogue: tcgen05.ld -> + bias, + residual (fp32), ReLU, ReLU mask (fp32 activation) -> fp32 rows
// Two smem stages (raw 32 KB + operand tiles 48 KB + weights 24 KB each).  The tensor core adds into its fp32 accumulator
// with truncation (measured: the error grows linearly with the number of accumulating instructions, 3.9e-5 of 7 at
// K = 1152 against 6e-6 for an FMA chain), so a tile has TWO accumulators: hi*whi -- the only product of full magnitude --
// goes to the first (K/16 instructions), the five small products (<= 2^-8 of it) to the second, and the epilogue adds the
// two in fp32.  Two tiles in flight: 2 x 2 x 64 TMEM columns.
// It replaces gemm_simt_kernel<float> as the fp32 mode's GEMM (gemm_simt stays as the cross-check in tests/).
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "gemm_sm100.cuh"
#include "gemm_tc32.cuh"

namespace rgie {
namespace {

constexpr int TC_BN = 64;
constexpr int TC_STAGES = 2;
constexpr int TC_RAW_BYTES = BM * BK * 4;             // 128 rows x 64 fp32 (two boxes of 128 x 32)
constexpr int TC_A_PLANE = BM * BK * 2;               // one bf16 operand tile
constexpr int TC_W_PLANE = TC_BN * BK * 2;            // one bf16 weight tile
struct TcSmem {
  static constexpr int RAW_OFF = 0;
  static constexpr int A_OFF = RAW_OFF + TC_STAGES * TC_RAW_BYTES;             // [stage][3 planes]
  static constexpr int W_OFF = A_OFF + TC_STAGES * 3 * TC_A_PLANE;             // [stage][3 planes]
  static constexpr int BAR_OFF = W_OFF + TC_STAGES * 3 * TC_W_PLANE;           // full[S], conv[S], empty[S], tfull[2], tempty[2]
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (3 * TC_STAGES + 4) * 8;
  static constexpr int BIAS_OFF = (TMEM_PTR_OFF + 16 + 15) & ~15;
  static constexpr int TOTAL = BIAS_OFF + MAX_BIAS * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static_assert(DYN_BYTES <= 232448, "shared memory plan exceeds 227 KB");
};
constexpr int TC_THREADS = 320;

// the six (A piece, W piece) products, smallest first
__device__ __constant__ int kPairA[6] = {2, 0, 1, 1, 0, 0};
__device__ __constant__ int kPairW[6] = {0, 2, 1, 0, 1, 0};

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmW, const GemmDesc d, const int num_m_tiles, const int num_n_tiles,
                 const FastDiv fd_nt, const int w_rows, const int n_pad) {
  using L = TcSmem;
  constexpr int BN = TC_BN, STAGES = TC_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto conv_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR_OFF);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int kb_per_tap = d.Cin / BK;
  const int num_kb = d.ntaps * kb_per_tap;
  const int num_kb2 = d.A2 != nullptr ? d.Cin2 / BK : 0;
  const uint32_t w_bytes = (uint32_t)(w_rows * BK * 2);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 4);       // one arrive per converter warp
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);     // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L::TMEM_PTR_OFF),
                 "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (d.bias != nullptr) {
    float* sb = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    for (int i = threadIdx.x; i < d.Cout; i += TC_THREADS) sb[i] = d.bias[i];
  }
  if (w_rows < BN) {
    // fewer weight rows than the N tile (conv1 input gradient, N = 16): the unused rows of the weight tiles must hold
    // finite numbers (their output columns are never stored, but NaN bit patterns would be multiplied all the same)
    for (int i = threadIdx.x; i < STAGES * 3 * TC_W_PLANE / 4; i += TC_THREADS)
      reinterpret_cast<uint32_t*>(smem + L::W_OFF)[i] = 0u;
    fence_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto load_block = [&](const CUtensorMap* map, int col, int row, int wcol, int nt) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), (uint32_t)TC_RAW_BYTES + 3u * w_bytes);
        const uint32_t raw = smem_base + L::RAW_OFF + stage * TC_RAW_BYTES;
        tma_load_2d(raw, map, col, row, full_bar(stage));
        tma_load_2d(raw + TC_RAW_BYTES / 2, map, col + 32, row, full_bar(stage));
#pragma unroll
        for (int p = 0; p < 3; ++p)
          tma_load_2d(smem_base + L::W_OFF + (stage * 3 + p) * TC_W_PLANE, &tmW, wcol, p * n_pad + nt * BN, full_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = (int)fd_nt.div((uint32_t)tile), nt = tile - mt * num_n_tiles;
        const long m0 = d.m_begin + (long)mt * BM;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          load_block(&tmA, cb * BK, (int)(m0 + d.row_off[tap]), kb * BK, nt);
          if (++cb == kb_per_tap) { cb = 0; ++tap; }
        }
        if (m0 < d.a2_rows)
          for (int kb = 0; kb < num_kb2; ++kb) load_block(&tmA2, kb * BK, (int)m0, (num_kb + kb) * BK, nt);
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
        const uint32_t tmem_big = tmem_base + (uint32_t)(acc * 2 * BN);        // hi * whi
        const uint32_t tmem_small = tmem_big + (uint32_t)BN;                   // the five small products
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(full_bar(stage), phase);          // the weight planes have landed
          mbar_wait(conv_bar(stage), phase);          // the three operand tiles are written
          tcgen05_fence_after();
#pragma unroll 1
          for (int pr = 0; pr < 6; ++pr) {
            const uint64_t adesc = make_smem_desc(smem_base + L::A_OFF + (stage * 3 + kPairA[pr]) * TC_A_PLANE);
            const uint64_t bdesc = make_smem_desc(smem_base + L::W_OFF + (stage * 3 + kPairW[pr]) * TC_W_PLANE);
            const uint32_t tmem_d = pr == 5 ? tmem_big : tmem_small;
            const int first = pr == 5 ? 0 : pr;        // pair 5 opens the big accumulator, pair 0 the small one
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | first | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else if (warp < 6) {
    // ===================== converters: fp32 row -> three bf16 operand rows =====================
    const int row = (warp - 2) * 32 + lane;
    const uint32_t swz = (uint32_t)(row & 7);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long m0 = d.m_begin + (long)((int)fd_nt.div((uint32_t)tile)) * BM;
      const int kb_total = num_kb + (m0 < d.a2_rows ? num_kb2 : 0);
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(full_bar(stage), phase);
        const uint32_t raw = smem_base + L::RAW_OFF + stage * TC_RAW_BYTES + (uint32_t)row * 128u;
        const uint32_t a0 = smem_base + L::A_OFF + (stage * 3) * TC_A_PLANE + (uint32_t)row * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                 // 8 output pieces of 8 bf16 = 16 input pieces of 4 fp32
          uint32_t x[8];
          const int box = j >> 2, pj = (2 * j) & 7;    // raw box (32 fp32 = 8 pieces per row) and first piece inside it
          lds128(raw + (uint32_t)(box * (TC_RAW_BYTES / 2)) + (((uint32_t)pj ^ swz) << 4), x);
          lds128(raw + (uint32_t)(box * (TC_RAW_BYTES / 2)) + (((uint32_t)(pj + 1) ^ swz) << 4), x + 4);
          uint32_t hi[4], mid[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float f0 = __uint_as_float(x[2 * e]), f1 = __uint_as_float(x[2 * e + 1]);
            const __nv_bfloat16 h0 = __float2bfloat16_rn(f0), h1 = __float2bfloat16_rn(f1);
            const float r0 = f0 - __bfloat162float(h0), r1 = f1 - __bfloat162float(h1);            // exact
            const __nv_bfloat16 m0b = __float2bfloat16_rn(r0), m1b = __float2bfloat16_rn(r1);
            const float s0 = r0 - __bfloat162float(m0b), s1 = r1 - __bfloat162float(m1b);          // exact
            const __nv_bfloat16 l0 = __float2bfloat16_rn(s0), l1 = __float2bfloat16_rn(s1);
            hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            mid[e] = (uint32_t)__bfloat16_as_ushort(m0b) | ((uint32_t)__bfloat16_as_ushort(m1b) << 16);
            lo[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          const uint32_t off = (((uint32_t)j ^ swz) << 4);
          sts128(a0 + off, hi);
          sts128(a0 + TC_A_PLANE + off, mid);
          sts128(a0 + 2 * TC_A_PLANE + off, lo);
        }
        fence_async_smem();             // operand tiles are read by tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 6..9: TMEM lane quarter = warp & 3) =====================
    constexpr int CH = 16;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float* sbias = reinterpret_cast<const float*>(smem + L::BIAS_OFF);
    const float* res = reinterpret_cast<const float*>(d.res);
    const float* mask = reinterpret_cast<const float*>(d.mask);
    float* D = reinterpret_cast<float*>(d.D);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = (int)fd_nt.div((uint32_t)tile), nt = tile - mt * num_n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long m = d.m_begin + (long)mt * BM + row;
      long dest = -1;
      if (m < d.m_end) dest = map_row(d.src, d.dst_kind, d.dst, m);
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN);
#pragma unroll
      for (int ci = 0; ci < BN / CH; ++ci) {
        uint32_t r[CH], r2[CH];
        tmem_ld<CH>(taddr + (uint32_t)(ci * CH), r);
        tmem_ld<CH>(taddr + (uint32_t)(BN + ci * CH), r2);
        tmem_ld_wait();
        const int n0 = nt * BN + ci * CH;
        if (dest >= 0 && n0 < d.Cout) {
          float v[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
          if (d.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] += sbias[n0 + j];
          }
          if (res != nullptr && m < d.res_rows) {
            const float4* rp = reinterpret_cast<const float4*>(res + m * d.ld_res + n0);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 t = rp[j];
              v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
          }
          if (d.relu) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (mask != nullptr) {
            const float4* mp = reinterpret_cast<const float4*>(mask + m * d.ld_mask + n0);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 t = mp[j];
              v[4 * j] = t.x > 0.f ? v[4 * j] : 0.f; v[4 * j + 1] = t.y > 0.f ? v[4 * j + 1] : 0.f;
              v[4 * j + 2] = t.z > 0.f ? v[4 * j + 2] : 0.f; v[4 * j + 3] = t.w > 0.f ? v[4 * j + 3] : 0.f;
            }
          }
          float4* op = reinterpret_cast<float4*>(D + dest * d.ldd + n0);
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace

// fp32 [n_pad, K] -> three bf16 planes [3][n_pad][K]: w = hi + mid + lo exactly (round-to-nearest at every step)
void split_weights_bf16x3(const float* w, size_t n, __nv_bfloat16* planes) {
  for (size_t i = 0; i < n; ++i) {
    const float x = w[i];
    const __nv_bfloat16 h = __float2bfloat16(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16(r1);
    const float r2 = r1 - __bfloat162float(m);
    planes[i] = h;
    planes[n + i] = m;
    planes[2 * n + i] = __float2bfloat16(r2);
  }
}

int build_gemm_tc32(const GemmDesc& d, GemmPlanTc32* p) {
  RGIE_CHECK(d.Cin % BK == 0, "gemm_tc32: Cin must be a multiple of 64");
  RGIE_CHECK(d.A2 == nullptr || d.Cin2 % BK == 0, "gemm_tc32: Cin2 must be a multiple of 64");
  RGIE_CHECK(d.ntaps >= 1 && d.ntaps <= kMaxTaps, "gemm_tc32: ntaps out of range");
  RGIE_CHECK(d.Cout % 16 == 0 && d.Cout <= MAX_BIAS, "gemm_tc32: Cout must be a multiple of 16 and <= 2048");
  RGIE_CHECK(d.n_pad == d.Cout, "gemm_tc32: the weight planes hold exactly Cout rows");
  RGIE_CHECK(d.ldd % 4 == 0 && (d.res == nullptr || d.ld_res % 4 == 0) && (d.mask == nullptr || d.ld_mask % 4 == 0),
             "gemm_tc32: leading dimensions must be multiples of 4 floats (16-byte accesses)");
  RGIE_CHECK(d.mask_bits == nullptr && d.D_bits == nullptr, "gemm_tc32: bit masks belong to the bf16 modes");
  RGIE_CHECK(d.a_rows < (1L << 31) && d.m_end < (1L << 31), "gemm_tc32: too many rows for a TMA coordinate");
  p->d = d;
  p->num_m_tiles = ceil_div(d.m_end - d.m_begin, (long)BM);
  p->num_n_tiles = ceil_div(d.Cout, TC_BN);
  p->w_rows = d.Cout < TC_BN ? d.Cout : TC_BN;
  const long tiles = (long)p->num_m_tiles * p->num_n_tiles;
  const int sms = gemm_sm100_num_sms();
  p->grid = (int)(tiles < sms ? (tiles < 1 ? 1 : tiles) : sms);
  const uint64_t ktot = (uint64_t)d.ntaps * d.Cin + (d.A2 ? d.Cin2 : 0);
  int rc = make_tensor_map_2d(&p->tmA, d.A, 1, (uint64_t)d.Cin, (uint64_t)d.a_rows, 32, BM, 128, (uint64_t)d.a_ld);
  if (rc) return rc;
  p->tmA2 = p->tmA;
  if (d.A2 != nullptr) rc = make_tensor_map_2d(&p->tmA2, d.A2, 1, (uint64_t)d.Cin2, (uint64_t)d.a2_rows, 32, BM, 128, 0);
  if (rc) return rc;
  return make_tensor_map_2d(&p->tmW, d.Wt, 0, ktot, (uint64_t)3 * d.n_pad, BK, (uint32_t)p->w_rows, 128, 0);
}

int run_gemm_tc32(const GemmPlanTc32& p, cudaStream_t st) {
  if (p.d.m_end <= p.d.m_begin) return 0;
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(gemm_tc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::DYN_BYTES));
    attr_once.done();
  }
  gemm_tc32_kernel<<<p.grid, TC_THREADS, TcSmem::DYN_BYTES, st>>>(p.tmA, p.tmA2, p.tmW, p.d, p.num_m_tiles, p.num_n_tiles,
                                                                  make_fastdiv((uint32_t)p.num_n_tiles), p.w_rows, p.d.n_pad);
  RGIE_LAUNCH_OK();
  return 0;
}

}  // namespace rgie

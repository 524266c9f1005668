// tcgen05 / TMEM / TMA backend of the row-shifted GEMM (common.cuh: GemmDesc) -- sm_100a only.
//
//   D[map(m), n] = epi( sum_t sum_c A[m + row_off[t], c] * Wt[n, t*Cin + c] )          bf16 x bf16 -> fp32 (TMEM)
//
// One persistent CTA per SM, 320 threads:
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d, 128B swizzle, zero fill for out-of-range rows = conv padding)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (cta_group::1, M=128, N=BN, K=16 per instruction)
//   warps 2..9  : epilogue       (tcgen05.ld 32x32b -> warp-private swizzled smem transpose -> bias / residual / ReLU /
//                                 ReLU-mask in a row-coalesced distribution -> row-remapped, fully coalesced stores;
//                                 residual/mask operands prefetched one chunk ahead)
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) between TMA and MMA; 2 TMEM accumulator buffers
// (tmem_full/tmem_empty) between MMA and epilogue so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "gemm_sm100.cuh"

namespace rgie {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 B = one swizzle-128B row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 320;          // TMA warp + MMA warp + 8 epilogue warps
constexpr int MAX_BIAS = 2048;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B operand tile (rows of 128 B, 8-row groups 1024 B apart): the sm_100 shared-memory matrix
// descriptor -- start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024>>4 |
// version=1 [46,48) | layout=2 (SWIZZLE_128B) [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  desc |= (uint64_t)1 << 16;
  desc |= (uint64_t)(1024 >> 4) << 32;
  desc |= (uint64_t)1 << 46;
  desc |= (uint64_t)2 << 61;
  return desc;
}

// instruction descriptor: c=f32 [4,6)=1 | a=bf16 [7,10)=1 | b=bf16 [10,13)=1 | K-major A,B (bits 15,16 = 0) |
// N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

__host__ __device__ constexpr int tmem_cols(int bn) { return 2 * bn <= 32 ? 32 : (2 * bn <= 64 ? 64 : (2 * bn <= 128 ? 128 : (2 * bn <= 256 ? 256 : 512))); }

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_STAGE_BYTES;
  static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE_BYTES;   // full[S], empty[S], tfull[2], tempty[2]
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * STAGES + 4) * 8;
  static constexpr int STAGE_OFF = (TMEM_PTR_OFF + 16 + 127) / 128 * 128;   // 8 warps x [32 rows x CH fp32] staging
  static constexpr int TOTAL = STAGE_OFF + 8 * 32 * (BN >= 32 ? 32 : 16) * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;   // + slack for manual 1024 B alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmDesc d, const int num_m_tiles, const int num_n_tiles) {
  using L = SmemLayout<BN, STAGES>;
  constexpr int CH = BN >= 32 ? 32 : 16;            // epilogue column chunk
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
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int kb_per_tap = d.Cin / BK;
  const int num_kb = d.ntaps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);     // one arrive per epilogue warp
    }
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
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / num_n_tiles, nt = tile - mt * num_n_tiles;
        const long m0 = d.m_begin + (long)mt * BM;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + L::B_STAGE_BYTES);
          tma_load_2d(smem_base + L::A_OFF + stage * A_STAGE_BYTES, &tmA, cb * BK, (int)(m0 + d.row_off[tap]),
                      full_bar(stage));
          tma_load_2d(smem_base + L::B_OFF + stage * L::B_STAGE_BYTES, &tmB, kb * BK, nt * BN, full_bar(stage));
          if (++cb == kb_per_tap) { cb = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc(smem_base + L::A_OFF + stage * A_STAGE_BYTES);
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
  } else {
    // ===================== epilogue (8 warps) =====================
    // warps 2..9: TMEM lane quarter q = warp & 3 (hardware restriction on tcgen05.ld), column half = (warp - 2) >> 2.
    // tcgen05.ld hands every thread one accumulator ROW; storing it that way makes each warp-wide 16 B store hit 32
    // different rows (32 half-filled sectors per request).  So every 32-column chunk is transposed through a
    // warp-private, XOR-swizzled fp32 staging tile in shared memory, after which LPR = 4 lanes own one row and each
    // lane owns 8 consecutive columns: residual / mask loads and output stores then move 8 rows x 64 contiguous bytes
    // per warp instruction (full sectors), and the bias / residual / ReLU / mask math is done in that distribution.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int NCH = BN / CH;                     // chunks per tile
    constexpr int CPW = NCH >= 2 ? NCH / 2 : 1;      // chunks per warp
    constexpr int LPR = CH / 8;                      // lanes per row in the coalesced distribution (4 or 2)
    constexpr int RPP = 32 / LPR;                    // rows per pass (8 or 16)
    constexpr int NQ = CH / 4;                       // 16-byte fp32 quads per staged row (8 or 4)
    const int c_begin = NCH >= 2 ? half * CPW : 0;
    const bool active = NCH >= 2 || half == 0;
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(d.res);
    const __nv_bfloat16* mask = reinterpret_cast<const __nv_bfloat16*>(d.mask);
    float4* stage4 = reinterpret_cast<float4*>(smem + L::STAGE_OFF + (warp - 2) * (32 * CH * 4));
    const int lr = lane / LPR;                       // row within a pass
    const int lc = lane % LPR;                       // 8-column group within the chunk
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / num_n_tiles, nt = tile - mt * num_n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long m_base = d.m_begin + (long)mt * BM + q * 32;
      // rows this lane owns in the coalesced distribution: R_k = lr + k*RPP
      long mrow[LPR], dest[LPR];
#pragma unroll
      for (int k = 0; k < LPR; ++k) {
        mrow[k] = m_base + lr + k * RPP;
        dest[k] = (active && mrow[k] < d.m_end) ? map_row(d.src, d.dst_kind, d.dst, mrow[k]) : -1;
      }
      uint4 rcur[LPR], kcur[LPR], rnxt[LPR], knxt[LPR];
      auto fetch = [&](int c, uint4* rr, uint4* kk) {
        const int n = nt * BN + c * CH + lc * 8;
#pragma unroll
        for (int k = 0; k < LPR; ++k) {
          if (dest[k] >= 0) {
            if (res != nullptr && mrow[k] < d.res_rows) rr[k] = __ldg(reinterpret_cast<const uint4*>(res + mrow[k] * d.ld_res + n));
            if (mask != nullptr) kk[k] = __ldg(reinterpret_cast<const uint4*>(mask + mrow[k] * d.ld_mask + n));
          }
        }
      };
      fetch(c_begin, rcur, kcur);
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if (active) {
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int c = c_begin + ci;
          if (ci + 1 < CPW) fetch(c + 1, rnxt, knxt);
          uint32_t r[CH];
          tmem_ld<CH>(taddr + (uint32_t)(c * CH), r);
          tmem_ld_wait();
          // stage: thread = row `lane`, quad j at swizzled position j ^ (lane & (NQ-1))
#pragma unroll
          for (int j = 0; j < NQ; ++j)
            stage4[lane * NQ + (j ^ (lane & (NQ - 1)))] =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                            __uint_as_float(r[4 * j + 3]));
          __syncwarp();
          const int n0 = nt * BN + c * CH + lc * 8;
          float bias8[8];
          if (d.bias != nullptr) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(d.bias + n0));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(d.bias + n0 + 4));
            bias8[0] = b0.x; bias8[1] = b0.y; bias8[2] = b0.z; bias8[3] = b0.w;
            bias8[4] = b1.x; bias8[5] = b1.y; bias8[6] = b1.z; bias8[7] = b1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) bias8[e] = 0.f;
          }
#pragma unroll
          for (int k = 0; k < LPR; ++k) {
            const int R = lr + k * RPP;
            const float4 a0 = stage4[R * NQ + ((2 * lc) ^ (R & (NQ - 1)))];
            const float4 a1 = stage4[R * NQ + ((2 * lc + 1) ^ (R & (NQ - 1)))];
            if (dest[k] < 0) continue;
            float v[8] = {a0.x + bias8[0], a0.y + bias8[1], a0.z + bias8[2], a0.w + bias8[3],
                          a1.x + bias8[4], a1.y + bias8[5], a1.z + bias8[6], a1.w + bias8[7]};
            if (res != nullptr && mrow[k] < d.res_rows) {
              const uint4 u = rcur[k];
              v[0] += bf16_lo(u.x); v[1] += bf16_hi(u.x); v[2] += bf16_lo(u.y); v[3] += bf16_hi(u.y);
              v[4] += bf16_lo(u.z); v[5] += bf16_hi(u.z); v[6] += bf16_lo(u.w); v[7] += bf16_hi(u.w);
            }
            if (d.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            if (mask != nullptr) {
              const uint4 u = kcur[k];
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};     // bf16 > 0  <=>  sign clear and magnitude non-zero
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t lo = w[e] & 0xFFFFu, hi = w[e] >> 16;
                if (!(lo != 0 && lo < 0x8000u)) v[2 * e] = 0.f;
                if (!(hi != 0 && hi < 0x8000u)) v[2 * e + 1] = 0.f;
              }
            }
            if (d.d_fp32) {
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(d.D) + dest[k] * d.ldd + n0);
              o[0] = make_float4(v[0], v[1], v[2], v[3]);
              o[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.D) + dest[k] * d.ldd + n0) =
                  make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                             pack_bf16x2(v[6], v[7]));
            }
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < LPR; ++k) { rcur[k] = rnxt[k]; kcur[k] = knxt[k]; }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
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

int make_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  RGIE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

template <int BN, int STAGES>
int run_impl(const GemmPlanSm100& p, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    RGIE_CUDA_OK(cudaFuncSetAttribute(gemm_sm100_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::DYN_BYTES));
    attr_set = true;
  }
  gemm_sm100_kernel<BN, STAGES><<<p.grid, NUM_THREADS, L::DYN_BYTES, st>>>(p.tmA, p.tmB, p.d, p.num_m_tiles,
                                                                           p.num_n_tiles);
  RGIE_LAUNCH_OK();
  return 0;
}

}  // namespace

int gemm_sm100_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int build_gemm_sm100(const GemmDesc& d, GemmPlanSm100* p) {
  RGIE_CHECK(d.Cin % BK == 0, "gemm_sm100: Cin must be a multiple of 64");
  RGIE_CHECK(d.ntaps >= 1 && d.ntaps <= kMaxTaps, "gemm_sm100: ntaps out of range");
  RGIE_CHECK(d.Cout % 16 == 0 && d.Cout <= MAX_BIAS, "gemm_sm100: Cout must be a multiple of 16 and <= 2048");
  RGIE_CHECK(d.a_rows < (1L << 31), "gemm_sm100: too many A rows for a TMA coordinate");
  int bn = d.Cout >= 256 ? 256 : (d.Cout >= 128 ? 128 : (d.Cout >= 64 ? 64 : 16));
  RGIE_CHECK(d.Cout % bn == 0 && d.n_pad % bn == 0, "gemm_sm100: Cout/n_pad must be a multiple of the N tile");
  if (!d.d_fp32) RGIE_CHECK(d.ldd % 8 == 0, "gemm_sm100: ldd must be a multiple of 8 for 16B stores");
  p->d = d;
  p->bn = bn;
  long M = d.m_end - d.m_begin;
  p->num_m_tiles = ceil_div(M, BM);
  p->num_n_tiles = d.Cout / bn;
  long tiles = (long)p->num_m_tiles * p->num_n_tiles;
  int sms = gemm_sm100_num_sms();
  p->grid = (int)(tiles < sms ? tiles : sms);
  if (p->grid < 1) p->grid = 1;
  int rc = make_map_2d(&p->tmA, d.A, (uint64_t)d.Cin, (uint64_t)d.a_rows, BK, BM);
  if (rc) return rc;
  return make_map_2d(&p->tmB, d.Wt, (uint64_t)d.ntaps * d.Cin, (uint64_t)d.n_pad, BK, (uint32_t)bn);
}

int run_gemm_sm100(const GemmPlanSm100& p, cudaStream_t st) {
  if (p.d.m_end <= p.d.m_begin) return 0;
  switch (p.bn) {
    case 256: return run_impl<256, 4>(p, st);
    case 128: return run_impl<128, 6>(p, st);
    case 64: return run_impl<64, 8>(p, st);
    case 16: return run_impl<16, 8>(p, st);
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

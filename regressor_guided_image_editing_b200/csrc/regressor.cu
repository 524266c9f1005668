// Valence/arousal regressor: torchvision resnet50 (eval, BatchNorm folded) on 10 random crops per image, forward and
// input-gradient backward.  Reference: src/baselines/models/EmotionPredictionModel.py:10-54 (load_model_eval),
// utilities/ReplicateAndCrop.py:30-45, torchvision.models.resnet50.
//
// Every convolution is lowered to the row-shifted GEMM of common.cuh (3x3 = nine row-shifted operand loads over a
// zero-padded flat pixel layout; stride-2 convs read a 2x2 phase-split layout so they are row shifts too; conv1 7x7/2
// becomes 4 vertical taps over a space-to-depth + 4-horizontal-tap packed input with 64 channels).  Backward keeps only
// the post-ReLU activations (for the ReLU masks) and the max-pool argmax; weight gradients are never computed.
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "gemm_tc32.cuh"
#include "rgie.h"

namespace rgie {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// small kernels around the GEMMs (templated on the activation type T: float = parity mode, bf16 = throughput mode)
// ---------------------------------------------------------------------------------------------------------------

// Input transform applied while the crops are packed.  mode 0: identity; 1: (v - 0.5) / 0.5 (ReplicateAndCrop's
// Normalize(.5,.5)); 2: the EmoNet ten-crop pipeline (EmoNet.py:63-88): t = clamp(v * ps + pb, 0, 1) ("denorm"), * 255,
// / 255, then ImageNet (t - mean_c) / std_c.  A crop whose `left` offset has bit 30 set is mirrored horizontally
// (EmoNet's flipped crops 5..9).
struct InXform {
  int mode;
  float ps, pb, mean[3], std[3];
};
constexpr int kFlipBit = 1 << 30;

// crop + normalise + 2x2 space-to-depth + 4 horizontal taps -> ZZ[n, i, j, bi*16 + (pr*2+pc)*3 + c]
template <typename T>
__global__ void __launch_bounds__(256) pack_crops_kernel(const float* __restrict__ img, const int* __restrict__ offsets,
                                                        const int* __restrict__ step_ptr, long off_step_stride,
                                                        T* __restrict__ zz, Geom g, int reps, int Hr, int Wr,
                                                        InXform xf, int n0) {
  // one thread block = one output line (n, i): no per-element division; a thread = one (pixel j, horizontal tap bi)
  const int* offs = offsets + (step_ptr ? (long)(*step_ptr) * off_step_stride : 0);
  const int n = blockIdx.x / g.H, i = blockIdx.x - n * g.H;
  const int b = (n0 + n) / reps;
  const int top = offs[2 * (n0 + n)], left_raw = offs[2 * (n0 + n) + 1];
  const bool flip = (left_raw & kFlipBit) != 0;
  const int left = left_raw & (kFlipBit - 1);
  const int cw = 2 * g.W;                          // crop width
  const float* src = img + (long)b * 3 * Hr * Wr + (long)(top + 2 * i) * Wr + left;
  const long plane = (long)Hr * Wr;
  T* dst_row = zz + geom_row(g, 0, n, i, 0) * 64;
  for (int t = threadIdx.x; t < g.W * 4; t += blockDim.x) {
    const int bi = t & 3, j = t >> 2;
    const int jj = j + bi - 2;
    T vals[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) vals[k] = from_f<T>(0.f);
    if (jj >= 0 && jj < g.W) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const float* q = src + c * plane + pr * Wr;               // (crop offsets are arbitrary: no vector alignment)
          float v0 = flip ? q[cw - 1 - 2 * jj] : q[2 * jj], v1 = flip ? q[cw - 2 - 2 * jj] : q[2 * jj + 1];
          if (xf.mode == 1) { v0 = (v0 - 0.5f) / 0.5f; v1 = (v1 - 0.5f) / 0.5f; }
          else if (xf.mode == 2) {
            v0 = fminf(fmaxf(__fadd_rn(__fmul_rn(v0, xf.ps), xf.pb), 0.f), 1.f) * 255.0f / 255.0f;
            v1 = fminf(fmaxf(__fadd_rn(__fmul_rn(v1, xf.ps), xf.pb), 0.f), 1.f) * 255.0f / 255.0f;
            v0 = (v0 - xf.mean[c]) / xf.std[c];
            v1 = (v1 - xf.mean[c]) / xf.std[c];
          }
          vals[(pr * 2 + 0) * 3 + c] = from_f<T>(v0);
          vals[(pr * 2 + 1) * 3 + c] = from_f<T>(v1);
        }
    }
    T* dst = dst_row + (long)j * 64 + bi * 16;
    if (sizeof(T) == 2) {
      const uint4* sv = reinterpret_cast<const uint4*>(vals);
      reinterpret_cast<uint4*>(dst)[0] = sv[0];
      reinterpret_cast<uint4*>(dst)[1] = sv[1];
    } else {
      const float4* sv = reinterpret_cast<const float4*>(vals);
#pragma unroll
      for (int k = 0; k < 4; ++k) reinterpret_cast<float4*>(dst)[k] = sv[k];
    }
  }
}

// The same operand WITHOUT the 4x horizontal-tap replication: Z16[n, i, jm, (pr*2+pc)*3 + c] (12 of 16 channels used), one
// 16-channel pixel per 2x2 input quad, memory pixel jm = j + 2 of a line of P = W + 4 pixels (two zero pixels left, two
// right: they are the conv's horizontal padding).  conv1 reads row m as the 64 elements starting at pixel m: the window over
// memory pixels m .. m+3 = image columns j-2 .. j+1, i.e. exactly the four horizontal taps -- the rows of the GEMM operand
// OVERLAP in memory (GemmDesc::a_ld = 16 < Cin = 64; the TMA tensor map simply has a 32-byte row stride).  The operand
// shrinks from 128 to 32 bytes per conv1 output pixel (6.5 -> 1.7 MB per crop), for the pack's writes and conv1's reads.
template <typename T>
__global__ void __launch_bounds__(256) pack_crops16_kernel(const float* __restrict__ img, const int* __restrict__ offsets,
                                                          const int* __restrict__ step_ptr, long off_step_stride,
                                                          T* __restrict__ zz, Geom g, int reps, int Hr, int Wr, InXform xf,
                                                          int n0) {
  // one thread block = one output line (n, i); a thread = one 16-channel pixel; n0: index of this launch's first crop in the batch
  const int* offs = offsets + (step_ptr ? (long)(*step_ptr) * off_step_stride : 0);
  const int n = blockIdx.x / g.H, i = blockIdx.x - n * g.H;
  const int b = (n0 + n) / reps;
  const int top = offs[2 * (n0 + n)], left_raw = offs[2 * (n0 + n) + 1];
  const bool flip = (left_raw & kFlipBit) != 0;
  const int left = left_raw & (kFlipBit - 1);
  const int cw = 2 * g.W;
  const float* src = img + (long)b * 3 * Hr * Wr + (long)(top + 2 * i) * Wr + left;
  const long plane = (long)Hr * Wr;
  T* dst_row = zz + (geom_row(g, 0, n, i, 0) + 2) * 16;
  for (int j = threadIdx.x; j < g.W; j += blockDim.x) {
    T vals[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) vals[k] = from_f<T>(0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        const float* q = src + c * plane + pr * Wr;
        float v0 = flip ? q[cw - 1 - 2 * j] : q[2 * j], v1 = flip ? q[cw - 2 - 2 * j] : q[2 * j + 1];
        if (xf.mode == 1) { v0 = (v0 - 0.5f) / 0.5f; v1 = (v1 - 0.5f) / 0.5f; }
        else if (xf.mode == 2) {
          v0 = fminf(fmaxf(__fadd_rn(__fmul_rn(v0, xf.ps), xf.pb), 0.f), 1.f) * 255.0f / 255.0f;
          v1 = fminf(fmaxf(__fadd_rn(__fmul_rn(v1, xf.ps), xf.pb), 0.f), 1.f) * 255.0f / 255.0f;
          v0 = (v0 - xf.mean[c]) / xf.std[c];
          v1 = (v1 - xf.mean[c]) / xf.std[c];
        }
        vals[(pr * 2 + 0) * 3 + c] = from_f<T>(v0);
        vals[(pr * 2 + 1) * 3 + c] = from_f<T>(v1);
      }
    T* dst = dst_row + (long)j * 16;
    if (sizeof(T) == 2) {
      const uint4* sv = reinterpret_cast<const uint4*>(vals);
      reinterpret_cast<uint4*>(dst)[0] = sv[0];
      reinterpret_cast<uint4*>(dst)[1] = sv[1];
    } else {
      const float4* sv = reinterpret_cast<const float4*>(vals);
#pragma unroll
      for (int k = 0; k < 4; ++k) reinterpret_cast<float4*>(dst)[k] = sv[k];
    }
  }
}

// 16-byte vector of activations: 8 bf16 or 4 fp32
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };
template <typename T>
__device__ __forceinline__ void vec_load(const T* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  if (sizeof(T) == 2) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(w[e] << 16); v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u); }
  } else {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
}
template <typename T>
__device__ __forceinline__ void vec_store(T* p, const float* v) {
  uint4 u;
  if (sizeof(T) == 2) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]),
                   c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  } else {
    u.x = __float_as_uint(v[0]); u.y = __float_as_uint(v[1]); u.z = __float_as_uint(v[2]); u.w = __float_as_uint(v[3]);
  }
  *reinterpret_cast<uint4*>(p) = u;
}

// 3x3 stride-2 pad-1 max pool over plain NHWC [N,Hi,Hi,C] -> padded layout g (Ho = Hi/2).  Per element one byte:
// argmax code 0..8 (first max) | 0x10 when the maximum is > 0 (the pool input is the stem's post-ReLU activation, so
// the backward pass never has to re-read it for the ReLU mask).
// grid = (N * Ho) output rows; a thread = one output pixel x 16 bytes of channels (no 64-bit div/mod in the loop).
template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                         uint8_t* __restrict__ arg, Geom g, int Hi, int C) {
  constexpr int V = Vec16<T>::N;
  const int cvn = C / V;
  const int n = blockIdx.x / g.H, i = blockIdx.x - n * g.H;
  const T* in_n = in + (long)n * Hi * Hi * C;
  if constexpr (sizeof(T) == 2) {
    // bf16: the kernel was bound by instruction issue (ncu: issue slots 68 % busy, DRAM 48 %), so the 9-tap scan runs on
    // PACKED pairs: v > best as a 0xFFFF-per-half mask (__hgt2_mask), then best and the tap code are bitwise selects --
    // 4 instructions per pair and tap instead of ~8.  Same result as the scalar scan: first maximum wins (strict >), the
    // comparison of two bf16 values is exact.
    for (int t = threadIdx.x; t < g.W * cvn; t += blockDim.x) {
      const int j = t / cvn, c = (t - j * cvn) * V;
      uint32_t best[4], code[4];            // 4 pairs of bf16 / of 16-bit tap codes
      bool first = true;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int y = 2 * i - 1 + dy;
        if (y < 0 || y >= Hi) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int x = 2 * j - 1 + dx;
          if (x < 0 || x >= Hi) continue;
          const uint4 u = *reinterpret_cast<const uint4*>(in_n + ((long)y * Hi + x) * C + c);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          const uint32_t tc = (uint32_t)(dy * 3 + dx) * 0x00010001u;
          if (first) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { best[e] = w[e]; code[e] = tc; }
            first = false;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&w[e]),
                                             *reinterpret_cast<const __nv_bfloat162*>(&best[e]));
              best[e] = (w[e] & m) | (best[e] & ~m);
              code[e] = (tc & m) | (code[e] & ~m);
            }
          }
        }
      }
      *reinterpret_cast<uint4*>(out + geom_row(g, 0, n, i, j) * C + c) = make_uint4(best[0], best[1], best[2], best[3]);
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        code[e] |= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&best[e]), zero2) & 0x00100010u;   // "maximum > 0" flag
      uint2 pk;
      pk.x = __byte_perm(code[0], code[1], 0x6420);     // low byte of each 16-bit code: elements 0..3
      pk.y = __byte_perm(code[2], code[3], 0x6420);     // elements 4..7
      *reinterpret_cast<uint2*>(arg + (((long)n * g.H + i) * g.W + j) * C + c) = pk;
    }
  } else {
  for (int t = threadIdx.x; t < g.W * cvn; t += blockDim.x) {
    const int j = t / cvn, c = (t - j * cvn) * V;
    float best[V];
    int code[V];
    bool first = true;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = 2 * i - 1 + dy;
      if (y < 0 || y >= Hi) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int x = 2 * j - 1 + dx;
        if (x < 0 || x >= Hi) continue;
        float v[V];
        vec_load<T>(in_n + ((long)y * Hi + x) * C + c, v);
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (first || v[e] > best[e]) { best[e] = v[e]; code[e] = dy * 3 + dx; }
        first = false;
      }
    }
    vec_store<T>(out + geom_row(g, 0, n, i, j) * C + c, best);
#pragma unroll
    for (int e = 0; e < V; ++e) code[e] |= best[e] > 0.f ? 0x10 : 0;
    uint8_t* a = arg + (((long)n * g.H + i) * g.W + j) * C + c;
    if (V == 8) {
      uint2 pk;
      pk.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
      pk.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
      *reinterpret_cast<uint2*>(a) = pk;
    } else {
      *reinterpret_cast<uint32_t*>(a) = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
    }
  }
  }
}
// backward of the pool fused with the stem ReLU mask: dC1[n,y,x,c] (layout gd) = sum over the windows whose argmax is
// (y,x) AND whose maximum is positive of dP.  grid = (N * Ho) rows of 2x2 input quads; a thread = one quad (a,b) x 16 bytes
// of channels: pixel (2a+ry, 2b+rx) can only be the argmax of windows (a..a+1, b..b+1), so four window reads (argmax byte
// + gradient) produce the four outputs -- the stem activation itself is not read.
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ dP, const uint8_t* __restrict__ arg,
                                                         T* __restrict__ dC1, Geom gp, Geom gd, int C) {
  constexpr int V = Vec16<T>::N;
  const int cvn = C / V;
  const int n = blockIdx.x / gp.H, a = blockIdx.x - n * gp.H;
  for (int t = threadIdx.x; t < gp.W * cvn; t += blockDim.x) {
    const int b = t / cvn, c = (t - b * cvn) * V;
    float s[2][2][V];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < V; ++e) s[q >> 1][q & 1][e] = 0.f;
#pragma unroll
    for (int wi = 0; wi < 2; ++wi) {
      const int i = a + wi;
      if (i >= gp.H) continue;
#pragma unroll
      for (int wj = 0; wj < 2; ++wj) {
        const int j = b + wj;
        if (j >= gp.W) continue;
        const uint8_t* ap = arg + (((long)n * gp.H + i) * gp.W + j) * C + c;
        uint32_t codes[2];
        if (V == 8) { const uint2 pk = *reinterpret_cast<const uint2*>(ap); codes[0] = pk.x; codes[1] = pk.y; }
        else { codes[0] = *reinterpret_cast<const uint32_t*>(ap); codes[1] = 0; }
        float gr[V];
        vec_load<T>(dP + geom_row(gp, 0, n, i, j) * C + c, gr);
        // window (i,j) covers input rows 2i-1..2i+1: quad row ry maps to dy = ry+1 (wi = 0) or ry-1 (wi = 1, ry = 1 only)
#pragma unroll
        for (int ry = wi; ry < 2; ++ry)
#pragma unroll
          for (int rx = wj; rx < 2; ++rx) {
            const uint32_t want = 0x10u | (uint32_t)((wi ? 0 : ry + 1) * 3 + (wj ? 0 : rx + 1));
#pragma unroll
            for (int e = 0; e < V; ++e)
              if (((codes[e >> 2] >> (8 * (e & 3))) & 0xFF) == want) s[ry][rx][e] += gr[e];
          }
      }
    }
#pragma unroll
    for (int ry = 0; ry < 2; ++ry)
#pragma unroll
      for (int rx = 0; rx < 2; ++rx) vec_store<T>(dC1 + geom_row(gd, 0, n, 2 * a + ry, 2 * b + rx) * C + c, s[ry][rx]);
  }
}

// global average pool over the valid window of layout g: grid (N, C / (32*V)), 256 threads = 32 channel vectors x 8 pixel
// groups; 16-byte channel vectors, coalesced over channels
template <typename T>
__global__ void __launch_bounds__(256) avgpool_kernel(const T* __restrict__ x, Geom g, int C, float* __restrict__ feat) {
  constexpr int V = Vec16<T>::N;
  __shared__ float red[8][32 * V];
  const int n = blockIdx.x;
  const int cv = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + cv) * V;
  float s[V];
#pragma unroll
  for (int e = 0; e < V; ++e) s[e] = 0.f;
  const int npx = g.H * g.W;
  for (int p = pg; p < npx; p += 8) {
    const int i = p / g.W, j = p - i * g.W;
    float v[V];
    vec_load<T>(x + geom_row(g, 0, n, i, j) * C + c, v);
#pragma unroll
    for (int e = 0; e < V; ++e) s[e] += v[e];
  }
#pragma unroll
  for (int e = 0; e < V; ++e) red[pg][cv * V + e] = s[e];
  __syncthreads();
  if (pg == 0) {
    const float inv = 1.0f / (float)npx;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += red[q][cv * V + e];
      feat[(long)n * C + c + e] = t * inv;
    }
  }
}
// logits[n,k] = b[k] + sum_c feat[n,c] * w[k,c]
__global__ void __launch_bounds__(256) fc_kernel(const float* __restrict__ feat, int C, const float* __restrict__ wfc,
                                                const float* __restrict__ bfc, int K, float* __restrict__ logits) {
  __shared__ float red[8][8];
  const int n = blockIdx.x;
  float part[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) part[k] = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float f = feat[(long)n * C + c];
    for (int k = 0; k < K; ++k) part[k] = fmaf(f, wfc[(long)k * C + c], part[k]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < K; ++k) {
    float v = part[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float t = bfc[threadIdx.x];
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    logits[(long)n * K + threadIdx.x] = t;
  }
}
// dfeat[n,c] = (sum_k dlogits[n,k] * w[k,c]) / (H*W)
__global__ void __launch_bounds__(256) fc_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ wfc, int K,
                                                    int C, float inv, float* __restrict__ dfeat) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float d = 0.f;
    for (int k = 0; k < K; ++k) d = fmaf(dlogits[(long)n * K + k], wfc[(long)k * C + c], d);
    dfeat[(long)n * C + c] = d * inv;
  }
}
// d(layer4 out)[n,i,j,c] = (out > 0) ? dfeat[n,c] : 0    (one thread = one pixel x 16 bytes of channels)
template <typename T>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dfeat, const T* __restrict__ x,
                                                         T* __restrict__ dx, Geom g, int C, long total) {
  constexpr int V = Vec16<T>::N;
  const int cvn = C / V;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cvn) * V;
    long q = idx / cvn;
    const int j = (int)(q % g.W); q /= g.W;
    const int i = (int)(q % g.H);
    const int n = (int)(q / g.H);
    const long r = geom_row(g, 0, n, i, j) * C + c;
    float v[V], o[V];
    vec_load<T>(x + r, v);
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = v[e] > 0.f ? dfeat[(long)n * C + c + e] : 0.f;
    vec_store<T>(dx + r, o);
  }
}

// dimg[b,c,Y,X] = nscale * sum_r dZ[(b*reps+r), (Y-top)>>1, (X-left)>>1, ((Y-top)&1)*2+((X-left)&1))*3 + c]
__global__ void __launch_bounds__(256) crop_grad_gather_kernel(const float* __restrict__ dz, const int* __restrict__ offsets,
                                                              const int* __restrict__ step_ptr, long off_step_stride,
                                                              float* __restrict__ dimg, const float* __restrict__ img, int B,
                                                              int reps, int Hr, int Wr, int crop, InXform xf) {
  // one thread block = one image line (b, Y); threads along X (no per-element division)
  const int* offs = offsets + (step_ptr ? (long)(*step_ptr) * off_step_stride : 0);
  const int Ho = crop / 2;
  const int b = blockIdx.x / Hr, Y = blockIdx.x - b * Hr;
  const long plane = (long)Hr * Wr;
  for (int X = threadIdx.x; X < Wr; X += blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < reps; ++r) {
      const int n = b * reps + r;
      const int left_raw = offs[2 * n + 1];
      const int y = Y - offs[2 * n];
      int x = X - (left_raw & (kFlipBit - 1));
      if (y < 0 || y >= crop || x < 0 || x >= crop) continue;
      if (left_raw & kFlipBit) x = crop - 1 - x;
      const float* p = dz + (((long)n * Ho + (y >> 1)) * Ho + (x >> 1)) * 16 + ((y & 1) * 2 + (x & 1)) * 3;
      s0 += p[0]; s1 += p[1]; s2 += p[2];
    }
    float* o = dimg + (long)b * 3 * plane + (long)Y * Wr + X;
    float k0 = 1.f, k1 = 1.f, k2 = 1.f;
    if (xf.mode == 1) { k0 = k1 = k2 = 2.0f; }
    else if (xf.mode == 2) {
      const float* ip = img + (long)b * 3 * plane + (long)Y * Wr + X;
      const float t0 = __fadd_rn(__fmul_rn(ip[0], xf.ps), xf.pb), t1 = __fadd_rn(__fmul_rn(ip[plane], xf.ps), xf.pb),
                  t2 = __fadd_rn(__fmul_rn(ip[2 * plane], xf.ps), xf.pb);
      k0 = (t0 >= 0.f && t0 <= 1.f) ? xf.ps / xf.std[0] : 0.f;
      k1 = (t1 >= 0.f && t1 <= 1.f) ? xf.ps / xf.std[1] : 0.f;
      k2 = (t2 >= 0.f && t2 <= 1.f) ? xf.ps / xf.std[2] : 0.f;
    }
    o[0] = s0 * k0; o[plane] = s1 * k1; o[2 * plane] = s2 * k2;
  }
}

// activation (any layout) -> NCHW fp32, for parity taps
template <typename T>
__global__ void tap_kernel(const T* __restrict__ x, Geom g, int C, int full_h, int full_w, float* __restrict__ out, long total) {
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % full_w);
    long q = idx / full_w;
    const int i = (int)(q % full_h); q /= full_h;
    const int c = (int)(q % C);
    const int n = (int)(q / C);
    long row;
    if (g.planes == 4) row = geom_row(g, (i & 1) * 2 + (j & 1), n, i >> 1, j >> 1);
    else row = geom_row(g, 0, n, i, j);
    out[idx] = to_f<T>(x[row * C + c]);
  }
}

int grid_for(long total) {
  long g = (total + 255) / 256;
  const long cap = 148L * 32;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace
}  // namespace rgie

using namespace rgie;

// =================================================================================================================
// plan
// =================================================================================================================
struct ConvW {        // host fp32 folded weights
  const float* w; const float* b; int co, ci, k;
};

struct Block {
  int stage, ci, cm, co;
  bool ds, last, stride2;
  ConvW c1, c2, c3, dsw;
  // device weights (T) / biases (fp32)
  void *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *wds = nullptr;           // forward
  void *w1t = nullptr, *w2t[4] = {nullptr, nullptr, nullptr, nullptr}, *w3t = nullptr, *wdst = nullptr;   // dgrad
  int w2t_taps[4] = {0, 0, 0, 0};
  long w2t_off[4][kMaxTaps];
  float *b1 = nullptr, *b2 = nullptr, *b3 = nullptr, *bds = nullptr;
  // activations (+ their sign bits, 1 bit per element, bf16 modes only: the ReLU masks of the backward pass)
  void *h1 = nullptr, *h2 = nullptr, *out = nullptr;
  uint32_t *h1_bits = nullptr, *h2_bits = nullptr, *out_bits = nullptr;
  Geom gx, gs, gout;    // layout of X / of the stage / of OUT
  const void* x = nullptr;
  const uint32_t* x_bits = nullptr;
};

struct GemmOp {
  GemmDesc d;
  GemmPlanSm100 plan;
  GemmPlanTc32 tplan;   // fp32 mode: the bf16x3 tensor-core plan (gemm_tc32.cu)
  int fused_next = 0;   // this launch also computes the NEXT op of the list (gemm_b2b_kernel)
  int absorbed = 0;     // computed by the previous launch: skipped at run time
};

struct RgieRegressor {
  int precision = 0, dtype = 0, esz = 4;     // dtype 0 fp32 / 1 bf16
  int N = 0, crop = 0, K = 0;
  int H0 = 0;
  int zz16 = 0;          // conv1 operand as 16-channel pixels read through overlapped rows (pack_crops16_kernel)
  int tc32 = 0;          // fp32 mode on the tensor cores: weights are three bf16 planes, GEMMs run gemm_tc32_kernel
  // conv1 + max-pool in ONE launch (gemm_conv1_pool_kernel, bf16 tcgen05 mode): the 64-channel 224 x 224 stem activation is
  // never written to HBM.  `conv1_plain` keeps the plain conv1 plan (rgie_regressor_tap("stem") runs it on demand).
  int stem_pool = 0;
  GemmPlanSm100 conv1_plain;
  void* wh_dev = nullptr;     // conv_hshare weights (bf16 mode)
  int Hs[5] = {0, 0, 0, 0, 0};
  Geom gZZ, gDY, gS[5], gPh[5];
  std::vector<Block> blocks;
  std::vector<void*> allocs;
  long ws_bytes = 0;
  // stem
  void *wc1 = nullptr, *wc1t = nullptr; float* bc1 = nullptr;
  void *zz = nullptr, *c1 = nullptr, *p1 = nullptr; uint8_t* arg = nullptr;
  float *wfc = nullptr, *bfc = nullptr, *feat = nullptr, *dfeat = nullptr;
  // gradients
  void* dOut[5][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  void *dH2[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *dH1[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  void* dC1 = nullptr; float* dZ = nullptr;
  std::vector<GemmOp> fwd_ops, bwd_ops;
  // call state
  const int* offsets = nullptr; const int* step_ptr = nullptr; long off_stride = 0;
  int B = 0, reps = 0, Hr = 0, Wr = 0, normalize = 1;
  const float* img = nullptr;               // resized input of the last forward (read again by backward in transform mode 2)
  InXform xf2 = {2, 1.f, 0.f, {0.485f, 0.456f, 0.406f}, {0.229f, 0.224f, 0.225f}};   // mode-2 transform (EmoNet defaults)
  void* final_dout = nullptr;   // where the tail backward writes d(layer4 out)
  // optional per-GEMM timing (cudaEvent pairs), enabled by rgie_regressor_set_profiling
  bool profiling = false;
  std::vector<cudaEvent_t> ev;  // 2 per op: fwd ops then bwd ops
};

namespace {

int dev_alloc(RgieRegressor* R, void** p, size_t bytes, bool zero) {
  if (bytes == 0) bytes = 16;
  RGIE_CUDA_OK(cudaMalloc(p, bytes));
  if (zero) RGIE_CUDA_OK(cudaMemset(*p, 0, bytes));
  R->allocs.push_back(*p);
  R->ws_bytes += (long)bytes;
  return 0;
}

// upload a host fp32 matrix as T (fp32 or bf16)
int upload(RgieRegressor* R, const std::vector<float>& h, void** dptr) {
  if (R->tc32) {
    // fp32-accurate tensor-core mode: w = hi + mid + lo as three bf16 planes [3][rows][K]
    std::vector<__nv_bfloat16> planes(3 * h.size());
    split_weights_bf16x3(h.data(), h.size(), planes.data());
    if (int rc = dev_alloc(R, dptr, planes.size() * 2, false)) return rc;
    RGIE_CUDA_OK(cudaMemcpy(*dptr, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice));
    return 0;
  }
  if (int rc = dev_alloc(R, dptr, h.size() * R->esz, false)) return rc;
  if (R->dtype == 0) {
    RGIE_CUDA_OK(cudaMemcpy(*dptr, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  } else {
    std::vector<__nv_bfloat16> hb(h.size());
    for (size_t i = 0; i < h.size(); ++i) hb[i] = __float2bfloat16(h[i]);
    RGIE_CUDA_OK(cudaMemcpy(*dptr, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  }
  return 0;
}
int upload_f32(RgieRegressor* R, const float* h, size_t n, float** dptr) {
  if (int rc = dev_alloc(R, (void**)dptr, n * 4, false)) return rc;
  RGIE_CUDA_OK(cudaMemcpy(*dptr, h, n * 4, cudaMemcpyHostToDevice));
  return 0;
}

// phase decomposition of a stride-2 pad-1 3x3 tap offset d = r-1 in {-1,0,1}: plane parity and plane-row shift
inline void s2_phase(int d, int& par, int& shift) { par = d & 1; shift = (d - par) / 2; }

GemmDesc base_desc() {
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  return d;
}

int add_op(RgieRegressor* R, std::vector<GemmOp>& ops, const GemmDesc& d) {
  GemmOp op;
  op.d = d;
  if (R->precision == RGIE_PREC_BF16) {
    if (int rc = build_gemm_sm100(d, &op.plan)) return rc;
  } else if (R->tc32) {
    if (int rc = build_gemm_tc32(d, &op.tplan)) return rc;
  }
  ops.push_back(op);
  return 0;
}

// Back-to-back fusion pass over an op list (bf16 tcgen05 mode): a 256-wide 1x1 op followed by the 256 -> 64 1x1 op that reads
// its output (layer1: conv3 + skip -> next block's conv1; conv1 input gradient + skip gradient -> previous block's conv3
// input gradient) becomes ONE launch, and the 256-channel tensor is not re-read from HBM (-8.2 GB of 127.9 GB per 320 crops).
// MEASURED (B200, 320 crops, same box; gemm_b2b_kernel with the DMA-thread epilogue and SWIZZLE_64B half boxes as stage-2
// operands): per pair 1.01 -> 0.82, 1.16 -> 0.90, 1.19 -> 0.91, 1.19 -> 0.93 ms; step 74.5 / 74.9 -> 72.9 / 73.1 ms.
// ON by default; RGIE_GEMM_B2B=0 runs the two launches separately.  (Its two predecessors were slower or neutral, see the
// kernel's header comment.)
int fuse_b2b(RgieRegressor* R, std::vector<GemmOp>& ops) {
  static const int env_b2b = getenv("RGIE_GEMM_B2B") ? atoi(getenv("RGIE_GEMM_B2B")) : 1;
  if (R->precision != RGIE_PREC_BF16 || !env_b2b) return 0;
  for (size_t i = 0; i + 1 < ops.size(); ++i) {
    if (ops[i].absorbed || ops[i].plan.special || ops[i].plan.patch) continue;
    if (!gemm_b2b_eligible(ops[i].d, ops[i + 1].d)) continue;
    if (int rc = build_gemm_b2b_sm100(ops[i].d, ops[i + 1].d, &ops[i].plan)) return rc;
    ops[i].fused_next = 1;
    ops[i + 1].absorbed = 1;
    ++i;
  }
  return 0;
}

int run_op_raw(RgieRegressor* R, const GemmOp& op, cudaStream_t st) {
  if (op.absorbed) return 0;
  if (R->precision == RGIE_PREC_BF16) return run_gemm_sm100(op.plan, st);
  if (R->tc32) return run_gemm_tc32(op.tplan, st);
  return launch_gemm_simt(op.d, R->dtype, st);
}
// idx: position in [fwd_ops..., bwd_ops...]
int run_op(RgieRegressor* R, const GemmOp& op, cudaStream_t st, size_t idx) {
  if (!R->profiling) return run_op_raw(R, op, st);
  RGIE_CUDA_OK(cudaEventRecord(R->ev[2 * idx], st));
  int rc = run_op_raw(R, op, st);
  RGIE_CUDA_OK(cudaEventRecord(R->ev[2 * idx + 1], st));
  return rc;
}
// algorithmic flops of one op: 2 * (valid output rows) * Cout_useful * K
double op_flops(const GemmDesc& d, double useful_frac) {
  const Geom& g = d.src;
  double rows_per_plane = (double)g.n_img * g.H * g.W;
  double planes = (double)(d.m_end - d.m_begin) / (double)g.plane_rows();
  double f = 2.0 * rows_per_plane * planes * (double)d.Cout * (double)d.ntaps * (double)d.Cin * useful_frac;
  if (d.A2 != nullptr) f += 2.0 * rows_per_plane * (double)d.Cout * (double)d.Cin2;   // second operand: one plane of rows
  return f;
}

// algorithmic HBM bytes of one op: every distinct operand element read once, every output element written once
// (valid pixels only; the zero padding is never written and costs no algorithmic traffic)
double op_bytes(const GemmDesc& d, int esz) {
  const Geom& g = d.src;
  const double valid_frac = (double)g.H * g.W / (double)g.S;
  const double out_rows = (double)(d.m_end - d.m_begin) * valid_frac;
  double b = (double)d.a_rows * valid_frac * (d.a_ld ? d.a_ld : d.Cin) * esz;      // A (overlapped rows: every element once)
  if (d.A2 != nullptr) b += (double)d.a2_rows * valid_frac * d.Cin2 * esz;        // second operand
  b += (double)d.n_pad * (d.ntaps * d.Cin + (d.A2 ? d.Cin2 : 0)) * esz;            // weights
  b += out_rows * d.Cout * (d.d_fp32 ? 4 : esz);                                   // D
  if (d.res != nullptr) b += (double)(d.res_rows < d.m_end - d.m_begin ? d.res_rows : d.m_end - d.m_begin) * valid_frac * d.Cout * esz;
  if (d.mask != nullptr) b += out_rows * d.Cout * esz;
  if (d.mask_bits != nullptr) b += out_rows * d.Cout / 8.0;
  if (d.D_bits != nullptr) b += out_rows * d.Cout / 8.0;
  return b;
}

// set the ReLU mask of a backward op: sign bits in the bf16 modes, the activation itself in fp32 parity mode
void set_mask(const RgieRegressor* R, GemmDesc& d, const void* act, const uint32_t* bits, int channels) {
  if (R->dtype == 1) { d.mask = nullptr; d.mask_bits = bits; d.ld_mb = channels / 32; }
  else { d.mask = act; d.ld_mask = channels; }
}
void set_out_bits(const RgieRegressor* R, GemmDesc& d, uint32_t* bits, int channels) {
  if (R->dtype == 1) { d.D_bits = bits; d.ld_db = channels / 32; }
}

}  // namespace

extern "C" {

void rgie_regressor_destroy(RgieRegressor* R) {
  if (!R) return;
  for (cudaEvent_t e : R->ev) cudaEventDestroy(e);
  for (void* p : R->allocs) cudaFree(p);
  delete R;
}

long rgie_regressor_workspace_bytes(const RgieRegressor* R) { return R ? R->ws_bytes : 0; }

int rgie_regressor_create(const float* const* h_tensors, int n_tensors, int num_classes, int crop_size, int max_crops,
                          int precision, RgieRegressor** out) {
  RGIE_CHECK(out != nullptr && h_tensors != nullptr, "rgie_regressor_create: null argument");
  RGIE_CHECK(n_tensors == 2 + 2 * (3 * 16 + 4) + 2, "rgie_regressor_create: expected 108 tensors");
  RGIE_CHECK(crop_size % 32 == 0 && crop_size >= 64, "rgie_regressor_create: crop_size must be a multiple of 32");
  RGIE_CHECK(num_classes >= 1 && num_classes <= 8, "rgie_regressor_create: num_classes must be in 1..8");
  RGIE_CHECK(max_crops >= 1, "rgie_regressor_create: max_crops");
  RGIE_CHECK(precision >= 0 && precision <= 3, "rgie_regressor_create: precision");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("rgie_regressor_create: no CUDA device (there is no CPU fallback)");

  RgieRegressor* R = new RgieRegressor();
  struct Guard { RgieRegressor* r; bool ok = false; ~Guard() { if (!ok) rgie_regressor_destroy(r); } } guard{R};
  R->precision = precision;
  R->dtype = (precision == RGIE_PREC_FP32 || precision == RGIE_PREC_FP32_SIMT) ? 0 : 1;
  // fp32 mode: GEMMs on the tensor cores with the exact bf16x3 split (RGIE_FP32_SIMT=1 forces the CUDA-core kernel)
  static const int env_simt = getenv("RGIE_FP32_SIMT") ? atoi(getenv("RGIE_FP32_SIMT")) : 0;
  R->tc32 = (precision == RGIE_PREC_FP32 && !env_simt) ? 1 : 0;
  R->esz = R->dtype == 0 ? 4 : 2;
  const int N = R->N = max_crops;
  R->crop = crop_size;
  R->K = num_classes;
  const int H0 = R->H0 = crop_size / 2;
  for (int s = 1; s <= 4; ++s) R->Hs[s] = H0 >> s;
  // RGIE_ZZ16=0 keeps the replicated 64-channel conv1 operand (pack_crops_kernel).  Measured (B200, 640 crops per step,
  // same box): GEMM family 65.1 -> 63.1 ms per step, step 73.6 -> 72.7 ms; tools/microbench/tma_overlap.cu shows that the
  // driver accepts a row stride below the row extent and that the TMA unit delivers the overlapped windows.
  static const int env_zz16 = getenv("RGIE_ZZ16") ? atoi(getenv("RGIE_ZZ16")) : 1;
  R->zz16 = env_zz16;
  R->gZZ = R->zz16 ? make_geom(1, N, H0, H0, 2, 1, 0, 4) : make_geom(1, N, H0, H0, 2, 1, 0, 0);
  R->gDY = make_geom(1, N, H0, H0, 1, 2, 1, 2);
  for (int s = 1; s <= 4; ++s) {
    // one zero pad line above and one zero pad column left of every image: the pad line of the NEXT image (or the TMA
    // out-of-range fill after the last one) is the bottom padding, the pad column of the next line the right padding
    R->gS[s] = make_geom(1, N, R->Hs[s], R->Hs[s], 1, 0, 1, 0);
    R->gPh[s] = make_geom(4, N, R->Hs[s], R->Hs[s], 1, 0, 1, 0);
  }
  const int esz = R->esz;

  // ---- parse tensors
  int ti = 0;
  ConvW conv1{h_tensors[0], h_tensors[1], 64, 3, 7};
  ti = 2;
  const int nblk[5] = {0, 3, 4, 6, 3};
  int cin = 64;
  for (int s = 1; s <= 4; ++s) {
    const int cm = 64 << (s - 1), co = 4 * cm;
    for (int b = 0; b < nblk[s]; ++b) {
      Block k;
      k.stage = s; k.ci = cin; k.cm = cm; k.co = co;
      k.ds = (b == 0); k.last = (b == nblk[s] - 1); k.stride2 = (b == 0 && s > 1);
      k.c1 = ConvW{h_tensors[ti], h_tensors[ti + 1], cm, cin, 1};
      k.c2 = ConvW{h_tensors[ti + 2], h_tensors[ti + 3], cm, cm, 3};
      k.c3 = ConvW{h_tensors[ti + 4], h_tensors[ti + 5], co, cm, 1};
      ti += 6;
      if (k.ds) { k.dsw = ConvW{h_tensors[ti], h_tensors[ti + 1], co, cin, 1}; ti += 2; }
      R->blocks.push_back(k);
      cin = co;
    }
  }
  const float* h_wfc = h_tensors[ti];
  const float* h_bfc = h_tensors[ti + 1];
  for (int i = 0; i < n_tensors; ++i) RGIE_CHECK(h_tensors[i] != nullptr, "rgie_regressor_create: null tensor");

  // ---- stem weights: conv1 as 4 vertical taps over the packed input (64 channels = 4 horizontal taps x 16)
  {
    std::vector<float> w((size_t)64 * 256, 0.f), wt((size_t)16 * 1024, 0.f);
    for (int k = 0; k < 64; ++k)
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 7; ++r)
          for (int s = 0; s < 7; ++s) {
            const int dr = r - 3, dc = s - 3;
            const int pr = dr & 1, pc = dc & 1;
            const int a = (dr - pr) / 2, b = (dc - pc) / 2;      // in {-2,..,1}
            const int ai = a + 2, bi = b + 2, q = (pr * 2 + pc) * 3 + c;
            const float v = conv1.w[((k * 3 + c) * 7 + r) * 7 + s];
            w[(size_t)k * 256 + ai * 64 + bi * 16 + q] = v;
            wt[(size_t)q * 1024 + (ai * 4 + (3 - bi)) * 64 + k] = v;     // tap order j = 3 - bi: row offsets ascend
          }
    if (int rc = upload(R, w, &R->wc1)) return rc;
    if (int rc = upload(R, wt, &R->wc1t)) return rc;
    if (int rc = upload_f32(R, conv1.b, 64, &R->bc1)) return rc;
  }
  if (int rc = upload_f32(R, h_wfc, (size_t)num_classes * 2048, &R->wfc)) return rc;
  if (int rc = upload_f32(R, h_bfc, num_classes, &R->bfc)) return rc;

  // ---- stem buffers
  // (zz16: the last rows' windows reach 3 pixels past the last row; + one zero line keeps them inside the allocation)
  if (int rc = dev_alloc(R, &R->zz, ((size_t)R->gZZ.rows() + R->gZZ.P) * (R->zz16 ? 16 : 64) * esz, true)) return rc;
  if (int rc = dev_alloc(R, &R->c1, (size_t)N * H0 * H0 * 64 * esz, true)) return rc;
  if (int rc = dev_alloc(R, &R->p1, (size_t)R->gS[1].rows() * 64 * esz, true)) return rc;
  if (int rc = dev_alloc(R, (void**)&R->arg, (size_t)N * R->Hs[1] * R->Hs[1] * 64, true)) return rc;
  if (int rc = dev_alloc(R, (void**)&R->feat, (size_t)N * 2048 * 4, true)) return rc;
  if (int rc = dev_alloc(R, (void**)&R->dfeat, (size_t)N * 2048 * 4, true)) return rc;
  if (int rc = dev_alloc(R, &R->dC1, (size_t)R->gDY.rows() * 64 * esz, true)) return rc;
  if (int rc = dev_alloc(R, (void**)&R->dZ, (size_t)N * H0 * H0 * 16 * 4, true)) return rc;

  // ---- per-stage gradient scratch
  for (int s = 1; s <= 4; ++s) {
    const int cm = 64 << (s - 1), co = 4 * cm;
    const long rows = R->gS[s].rows();
    for (int p = 0; p < 2; ++p)
      if (int rc = dev_alloc(R, &R->dOut[s][p], (size_t)rows * co * esz, true)) return rc;
    if (int rc = dev_alloc(R, &R->dH2[s], (size_t)rows * cm * esz, true)) return rc;
    if (int rc = dev_alloc(R, &R->dH1[s], (size_t)(s > 1 ? 4 : 1) * rows * cm * esz, true)) return rc;
  }

  // ---- forward GEMM 0: conv1
  {
    GemmDesc d = base_desc();
    d.A = R->zz; d.a_rows = R->gZZ.rows(); d.Cin = 64;
    d.a_ld = R->zz16 ? 16 : 0;
    d.Wt = R->wc1; d.n_pad = 64; d.ntaps = 4;
    for (int a = 0; a < 4; ++a) d.row_off[a] = (long)(a - 2) * R->gZZ.P;
    d.m_begin = 0; d.m_end = R->gZZ.rows(); d.Cout = 64;
    d.src = R->gZZ; d.dst_kind = DST_TO_PLAIN; d.dst = R->gZZ;
    d.D = R->c1; d.ldd = 64; d.bias = R->bc1; d.relu = 1;
    if (int rc = add_op(R, R->fwd_ops, d)) return rc;
  }

  // ---- blocks: weights, buffers, forward ops
  const void* x = R->p1;
  const uint32_t* x_bits = nullptr;
  for (size_t bi = 0; bi < R->blocks.size(); ++bi) {
    Block& k = R->blocks[bi];
    const int s = k.stage;
    k.gs = R->gS[s];
    k.gx = k.stride2 ? R->gPh[s] : R->gS[s];
    k.gout = (k.last && s < 4) ? R->gPh[s + 1] : R->gS[s];
    k.x = x;
    const Geom& gs = k.gs;
    const long Mp = gs.rows();
    // weights
    {
      std::vector<float> w1((size_t)k.cm * k.ci), w1t((size_t)k.ci * k.cm);
      for (int n = 0; n < k.cm; ++n)
        for (int c = 0; c < k.ci; ++c) { w1[(size_t)n * k.ci + c] = k.c1.w[(size_t)n * k.ci + c]; w1t[(size_t)c * k.cm + n] = k.c1.w[(size_t)n * k.ci + c]; }
      if (int rc = upload(R, w1, &k.w1)) return rc;
      if (int rc = upload(R, w1t, &k.w1t)) return rc;
      std::vector<float> w3((size_t)k.co * k.cm), w3t((size_t)k.cm * k.co);
      for (int n = 0; n < k.co; ++n)
        for (int c = 0; c < k.cm; ++c) { w3[(size_t)n * k.cm + c] = k.c3.w[(size_t)n * k.cm + c]; w3t[(size_t)c * k.co + n] = k.c3.w[(size_t)n * k.cm + c]; }
      if (int rc = upload(R, w3, &k.w3)) return rc;
      if (int rc = upload(R, w3t, &k.w3t)) return rc;
      // 3x3: forward [cm, 9*cm] (t = r*3+s); dgrad: stride 1 -> one matrix [cm, 9*cm]; stride 2 -> one per input phase
      std::vector<float> w2((size_t)k.cm * 9 * k.cm);
      for (int n = 0; n < k.cm; ++n)
        for (int c = 0; c < k.cm; ++c)
          for (int t = 0; t < 9; ++t) w2[(size_t)n * 9 * k.cm + (size_t)t * k.cm + c] = k.c2.w[((size_t)n * k.cm + c) * 9 + t];
      if (int rc = upload(R, w2, &k.w2)) return rc;
      if (!k.stride2) {
        // tap slot tt = r*3 + j holds kernel tap (r, s = 2-j): the row offsets -((r-1)P + (s-1)) then ascend with j, so the
        // GEMM serves the three taps of a kernel row from one operand slab
        std::vector<float> w2t((size_t)k.cm * 9 * k.cm);
        for (int n = 0; n < k.cm; ++n)
          for (int c = 0; c < k.cm; ++c)
            for (int tt = 0; tt < 9; ++tt) {
              const int t = (tt / 3) * 3 + (2 - tt % 3);
              w2t[(size_t)c * 9 * k.cm + (size_t)tt * k.cm + n] = k.c2.w[((size_t)n * k.cm + c) * 9 + t];
            }
        if (int rc = upload(R, w2t, &k.w2t[0])) return rc;
        k.w2t_taps[0] = 9;
        for (int tt = 0; tt < 9; ++tt) {
          const int t = (tt / 3) * 3 + (2 - tt % 3);
          k.w2t_off[0][tt] = -((long)(t / 3 - 1) * gs.P + (t % 3 - 1));
        }
      } else {
        for (int ph = 0; ph < 4; ++ph) {
          const int pr = ph >> 1, pc = ph & 1;
          std::vector<int> taps;
          for (int t = 0; t < 9; ++t) {
            int par_r, sh_r, par_c, sh_c;
            s2_phase(t / 3 - 1, par_r, sh_r);
            s2_phase(t % 3 - 1, par_c, sh_c);
            if (par_r == pr && par_c == pc) {
              k.w2t_off[ph][taps.size()] = -((long)sh_r * gs.P + sh_c) - (long)ph * Mp;
              taps.push_back(t);
            }
          }
          k.w2t_taps[ph] = (int)taps.size();
          std::vector<float> w2t((size_t)k.cm * taps.size() * k.cm);
          for (int n = 0; n < k.cm; ++n)
            for (int c = 0; c < k.cm; ++c)
              for (size_t tt = 0; tt < taps.size(); ++tt)
                w2t[(size_t)c * taps.size() * k.cm + tt * k.cm + n] = k.c2.w[((size_t)n * k.cm + c) * 9 + taps[tt]];
          if (int rc = upload(R, w2t, &k.w2t[ph])) return rc;
        }
      }
      if (k.ds) {
        // the downsample branch rides in the same GEMMs as a second operand (K-concatenation):
        //   forward  out = relu([h2 | x_ds] . [W3 | Wds]^T + b3 + bds)        -> wds  = [co, cm + ci]
        //   backward dX  = mask * ([dH1 | dOut] . [W1^T | Wds^T]^T)           -> wdst = [ci, cm + co]
        const int kf = k.cm + k.ci, kb = k.cm + k.co;
        std::vector<float> wd((size_t)k.co * kf), wdt((size_t)k.ci * kb), bsum(k.co);
        for (int n = 0; n < k.co; ++n) {
          for (int c = 0; c < k.cm; ++c) wd[(size_t)n * kf + c] = k.c3.w[(size_t)n * k.cm + c];
          for (int c = 0; c < k.ci; ++c) wd[(size_t)n * kf + k.cm + c] = k.dsw.w[(size_t)n * k.ci + c];
          bsum[n] = k.c3.b[n] + k.dsw.b[n];
        }
        for (int c = 0; c < k.ci; ++c) {
          for (int n = 0; n < k.cm; ++n) wdt[(size_t)c * kb + n] = k.c1.w[(size_t)n * k.ci + c];
          for (int n = 0; n < k.co; ++n) wdt[(size_t)c * kb + k.cm + n] = k.dsw.w[(size_t)n * k.ci + c];
        }
        if (int rc = upload(R, wd, &k.wds)) return rc;
        if (int rc = upload(R, wdt, &k.wdst)) return rc;
        if (int rc = upload_f32(R, bsum.data(), k.co, &k.bds)) return rc;
      }
      if (int rc = upload_f32(R, k.c1.b, k.cm, &k.b1)) return rc;
      if (int rc = upload_f32(R, k.c2.b, k.cm, &k.b2)) return rc;
      if (int rc = upload_f32(R, k.c3.b, k.co, &k.b3)) return rc;
    }
    // activations
    if (int rc = dev_alloc(R, &k.h1, (size_t)k.gx.rows() * k.cm * esz, true)) return rc;
    if (int rc = dev_alloc(R, &k.h2, (size_t)Mp * k.cm * esz, true)) return rc;
    if (int rc = dev_alloc(R, &k.out, (size_t)k.gout.rows() * k.co * esz, true)) return rc;
    if (R->dtype == 1) {
      if (int rc = dev_alloc(R, (void**)&k.h1_bits, (size_t)bits_words(k.gx.rows(), k.cm / 32) * 4, true)) return rc;
      if (int rc = dev_alloc(R, (void**)&k.h2_bits, (size_t)bits_words(Mp, k.cm / 32) * 4, true)) return rc;
      if (int rc = dev_alloc(R, (void**)&k.out_bits, (size_t)bits_words(k.gout.rows(), k.co / 32) * 4, true)) return rc;
    }

    // c1
    {
      GemmDesc d = base_desc();
      d.A = k.x; d.a_rows = k.gx.rows(); d.Cin = k.ci; d.Wt = k.w1; d.n_pad = k.cm; d.ntaps = 1; d.row_off[0] = 0;
      d.m_begin = 0; d.m_end = k.gx.rows(); d.Cout = k.cm;
      d.src = k.gx; d.dst_kind = DST_SAME; d.dst = k.gx; d.D = k.h1; d.ldd = k.cm; d.bias = k.b1; d.relu = 1;
      set_out_bits(R, d, k.h1_bits, k.cm);
      if (int rc = add_op(R, R->fwd_ops, d)) return rc;
    }
    // c2
    {
      GemmDesc d = base_desc();
      d.A = k.h1; d.a_rows = k.gx.rows(); d.Cin = k.cm; d.Wt = k.w2; d.n_pad = k.cm; d.ntaps = 9;
      for (int t = 0; t < 9; ++t) {
        if (!k.stride2) d.row_off[t] = (long)(t / 3 - 1) * gs.P + (t % 3 - 1);
        else {
          int par_r, sh_r, par_c, sh_c;
          s2_phase(t / 3 - 1, par_r, sh_r);
          s2_phase(t % 3 - 1, par_c, sh_c);
          d.row_off[t] = (long)(par_r * 2 + par_c) * Mp + (long)sh_r * gs.P + sh_c;
        }
      }
      d.m_begin = 0; d.m_end = Mp; d.Cout = k.cm;
      d.src = gs; d.dst_kind = DST_SAME; d.dst = gs; d.D = k.h2; d.ldd = k.cm; d.bias = k.b2; d.relu = 1;
      set_out_bits(R, d, k.h2_bits, k.cm);
      if (int rc = add_op(R, R->fwd_ops, d)) return rc;
    }
    // c3 + residual (identity skip: epilogue operand; downsample branch: second GEMM operand) + relu
    {
      GemmDesc d = base_desc();
      d.A = k.h2; d.a_rows = Mp; d.Cin = k.cm; d.n_pad = k.co; d.ntaps = 1; d.row_off[0] = 0;
      d.m_begin = 0; d.m_end = Mp; d.Cout = k.co;
      d.src = gs; d.dst_kind = (k.last && s < 4) ? DST_TO_PHASE : DST_SAME; d.dst = k.gout;
      d.D = k.out; d.ldd = k.co; d.relu = 1;
      if (k.ds) {
        // 1x1 (stride-2: phase plane (0,0) of the phase-split input = rows [0, Mp)) downsample conv on the block input
        d.Wt = k.wds; d.bias = k.bds; d.A2 = k.x; d.a2_rows = Mp; d.Cin2 = k.ci;
      } else {
        d.Wt = k.w3; d.bias = k.b3; d.res = k.x; d.ld_res = k.co; d.res_rows = Mp;
      }
      set_out_bits(R, d, k.out_bits, k.co);
      if (int rc = add_op(R, R->fwd_ops, d)) return rc;
    }
    k.x_bits = x_bits;
    x_bits = k.out_bits;
    x = k.out;
  }

  // ---- backward ops (built in execution order: last block first)
  int pp[5] = {0, 0, 0, 0, 0};     // ping-pong index of the CURRENT dOut per stage
  R->final_dout = R->dOut[4][0];
  for (int bi = (int)R->blocks.size() - 1; bi >= 0; --bi) {
    Block& k = R->blocks[bi];
    const int s = k.stage;
    const Geom& gs = k.gs;
    const long Mp = gs.rows();
    void* dout = R->dOut[s][pp[s]];
    // c3 dgrad
    {
      GemmDesc d = base_desc();
      d.A = dout; d.a_rows = Mp; d.Cin = k.co; d.Wt = k.w3t; d.n_pad = k.cm; d.ntaps = 1; d.row_off[0] = 0;
      d.m_begin = 0; d.m_end = Mp; d.Cout = k.cm;
      d.src = gs; d.dst_kind = DST_SAME; d.dst = gs; d.D = R->dH2[s]; d.ldd = k.cm;
      set_mask(R, d, k.h2, k.h2_bits, k.cm);
      if (int rc = add_op(R, R->bwd_ops, d)) return rc;
    }
    // c2 dgrad
    if (!k.stride2) {
      GemmDesc d = base_desc();
      d.A = R->dH2[s]; d.a_rows = Mp; d.Cin = k.cm; d.Wt = k.w2t[0]; d.n_pad = k.cm; d.ntaps = 9;
      for (int t = 0; t < 9; ++t) d.row_off[t] = k.w2t_off[0][t];
      d.m_begin = 0; d.m_end = Mp; d.Cout = k.cm;
      d.src = gs; d.dst_kind = DST_SAME; d.dst = gs; d.D = R->dH1[s]; d.ldd = k.cm;
      set_mask(R, d, k.h1, k.h1_bits, k.cm);
      if (int rc = add_op(R, R->bwd_ops, d)) return rc;
    } else {
      for (int ph = 0; ph < 4; ++ph) {
        GemmDesc d = base_desc();
        d.A = R->dH2[s]; d.a_rows = Mp; d.Cin = k.cm; d.Wt = k.w2t[ph]; d.n_pad = k.cm; d.ntaps = k.w2t_taps[ph];
        for (int t = 0; t < d.ntaps; ++t) d.row_off[t] = k.w2t_off[ph][t];
        d.m_begin = (long)ph * Mp; d.m_end = (long)(ph + 1) * Mp; d.Cout = k.cm;
        d.src = k.gx; d.dst_kind = DST_SAME; d.dst = k.gx; d.D = R->dH1[s]; d.ldd = k.cm;
        set_mask(R, d, k.h1, k.h1_bits, k.cm);
        if (int rc = add_op(R, R->bwd_ops, d)) return rc;
      }
    }
    // c1 dgrad (+ skip gradient as epilogue operand / downsample-branch gradient as second GEMM operand,
    //           * ReLU mask of the block input)
    {
      GemmDesc d = base_desc();
      d.A = R->dH1[s]; d.a_rows = k.gx.rows(); d.Cin = k.cm; d.n_pad = k.ci; d.ntaps = 1; d.row_off[0] = 0;
      d.m_begin = 0; d.m_end = k.gx.rows(); d.Cout = k.ci;
      d.src = k.gx;
      if (k.ds) { d.Wt = k.wdst; d.A2 = dout; d.a2_rows = Mp; d.Cin2 = k.co; }
      else { d.Wt = k.w1t; d.res = dout; d.ld_res = k.ci; d.res_rows = Mp; }
      if (bi == 0) {
        // block input = max-pool output (no ReLU of its own)
        d.dst_kind = DST_SAME; d.dst = gs; d.D = R->dOut[1][pp[1] ^ 1];   // = d(pool out)
        pp[1] ^= 1;
      } else if (k.stride2) {
        set_mask(R, d, k.x, k.x_bits, k.ci);
        d.dst_kind = DST_FROM_PHASE; d.dst = R->gS[s - 1]; d.D = R->dOut[s - 1][pp[s - 1]];
      } else {
        set_mask(R, d, k.x, k.x_bits, k.ci);
        d.dst_kind = DST_SAME; d.dst = gs; d.D = R->dOut[s][pp[s] ^ 1];
        pp[s] ^= 1;
      }
      d.ldd = k.ci;
      if (int rc = add_op(R, R->bwd_ops, d)) return rc;
    }
  }
  // conv1 dgrad: dZ[m, q] = sum_{a,b} dC1[m - (a*P + b)] . Wt
  {
    GemmDesc d = base_desc();
    d.A = R->dC1; d.a_rows = R->gDY.rows(); d.Cin = 64; d.Wt = R->wc1t; d.n_pad = 16; d.ntaps = 16;
    for (int a = 0; a < 4; ++a)
      for (int j = 0; j < 4; ++j) d.row_off[a * 4 + j] = -((long)(a - 2) * R->gDY.P + ((3 - j) - 2));   // ascending in j
    d.m_begin = 0; d.m_end = R->gDY.rows(); d.Cout = 16;
    d.src = R->gDY; d.dst_kind = DST_TO_PLAIN; d.dst = R->gDY; d.D = R->dZ; d.ldd = 16; d.d_fp32 = 1;
    static const int env_hs = getenv("RGIE_CONV1_HSHARE") ? atoi(getenv("RGIE_CONV1_HSHARE")) : 1;
    if (R->precision == RGIE_PREC_BF16 && env_hs) {
      // tcgen05 path: the four horizontal taps become the N dimension (gemm_sm100.cu: conv_hshare_kernel).
      // Wh[(j*12 + q), yi*64 + co] = weight of tap (dy = yi - 1, dx = j - 1) for output q = (pr*2+pc)*3 + c; in the 16-tap
      // matrix above tap slot (a, j) has dy = 2 - a, dx = j - 1, so yi = 3 - a.
      std::vector<float> wh((size_t)48 * 256, 0.f);
      // rebuild from the folded conv1 weights exactly like wc1t (same index algebra)
      for (int k = 0; k < 64; ++k)
        for (int c = 0; c < 3; ++c)
          for (int r = 0; r < 7; ++r)
            for (int s7 = 0; s7 < 7; ++s7) {
              const int dr = r - 3, dc = s7 - 3;
              const int pr = dr & 1, pc = dc & 1;
              const int a = (dr - pr) / 2, b = (dc - pc) / 2;      // in {-2,..,1}
              const int ai = a + 2, bi = b + 2, q = (pr * 2 + pc) * 3 + c;
              const int j = 3 - bi, yi = 3 - ai;
              wh[(size_t)(j * 12 + q) * 256 + yi * 64 + k] = conv1.w[((k * 3 + c) * 7 + r) * 7 + s7];
            }
      void* wh_dev = nullptr;
      if (int rc = upload(R, wh, &wh_dev)) return rc;
      R->wh_dev = wh_dev;
      GemmOp op;
      op.d = d;
      if (int rc = build_conv_hshare_sm100(d, wh_dev, -1, -1, &op.plan)) return rc;
      R->bwd_ops.push_back(op);
    } else {
      if (int rc = add_op(R, R->bwd_ops, d)) return rc;
    }
  }
  // (An "L2-resident stem" -- pack -> conv1 -> max-pool and the mirrored backward per group of 4-16 crops through small reused
  //  buffers -- was measured on B200: 74.2-75.9 ms per step against 72.8 for the whole batch at once; the 8 GB per micro-batch
  //  that stopped travelling through HBM were paid back with interest by 40-160 small launches per pass.  It is no longer in
  //  the file; the stem activation now stays on chip through gemm_conv1_pool_kernel instead.)
  if (int rc = fuse_b2b(R, R->fwd_ops)) return rc;
  if (int rc = fuse_b2b(R, R->bwd_ops)) return rc;
  {
    // layer1's 3x3 convs and their input gradients with the horizontal taps as the N dimension (conv3_hshare_kernel: 12
    // instructions of N = 192 per tile instead of 36 of N = 64).  RGIE_CONV3_HSHARE=0 keeps the CTA-pair patch kernel.
    static const int env_h3 = getenv("RGIE_CONV3_HSHARE") ? atoi(getenv("RGIE_CONV3_HSHARE")) : 1;
    if (env_h3 && R->precision == RGIE_PREC_BF16) {
      for (std::vector<GemmOp>* ops : {&R->fwd_ops, &R->bwd_ops})
        for (GemmOp& op : *ops) {
          if (op.absorbed || op.fused_next || op.plan.patch != 1 || op.d.d_fp32 || op.d.mask != nullptr || op.d.res != nullptr) continue;
          void* wh = nullptr;
          if (int rc = dev_alloc(R, &wh, (size_t)192 * 192 * 2, false)) return rc;
          if (int rc = build_conv3_hshare_sm100(op.d, wh, &op.plan)) return rc;
        }
    }
  }
  {
    // conv1 + max-pool fused (RGIE_STEM_POOL=0 keeps conv1 -> c1 -> maxpool_fwd_kernel).  Needs the 16-channel conv1 operand,
    // the single-CTA 4-tap patch plan and a 224-type geometry (width a multiple of 8).
    static const int env_pool = getenv("RGIE_STEM_POOL") ? atoi(getenv("RGIE_STEM_POOL")) : 1;
    GemmOp& c1op = R->fwd_ops[0];
    if (env_pool && R->precision == RGIE_PREC_BF16 && R->zz16 && c1op.plan.patch == 2 && !c1op.plan.patch_2cta &&
        !c1op.fused_next && R->H0 % 8 == 0 && R->Hs[1] * 2 == R->H0) {
      R->conv1_plain = c1op.plan;
      if (int rc = build_conv1_pool_sm100(c1op.d, R->p1, R->gS[1], R->arg, R->Hs[1], &c1op.plan)) return rc;
      R->stem_pool = 1;
    }
  }
  RGIE_CUDA_OK(cudaDeviceSynchronize());
  guard.ok = true;
  *out = R;
  return 0;
}

int rgie_regressor_forward_ex(RgieRegressor* R, const float* img, int B, int Hr, int Wr, const int* offsets,
                              const int* step_ptr, long off_step_stride, int reps, int normalize, float* logits,
                              void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RGIE_CHECK(R && img && offsets && logits, "rgie_regressor_forward: null argument");
  RGIE_CHECK(B * reps == R->N, "rgie_regressor_forward: B*reps must equal the max_crops the handle was created with");
  RGIE_CHECK(Hr >= R->crop && Wr >= R->crop, "rgie_regressor_forward: image smaller than the crop");
  R->offsets = offsets; R->step_ptr = step_ptr; R->off_stride = off_step_stride;
  R->B = B; R->reps = reps; R->Hr = Hr; R->Wr = Wr; R->normalize = normalize; R->img = img;
  RGIE_CHECK(normalize >= 0 && normalize <= 2, "rgie_regressor_forward: normalize must be 0, 1 or 2");
  InXform xf = R->xf2;
  xf.mode = normalize;
  const int N = R->N, H0 = R->H0, H1 = R->Hs[1];
  const int pack_threads = (H0 * 4) % 224 == 0 ? 224 : 256;     // 4 work items per output pixel: whole iterations per line
  // stem: pack -> conv1 (+ max-pool in the same launch when fused) -> max-pool
  {
    if (R->zz16) {
      if (R->dtype == 0)
        pack_crops16_kernel<float><<<N * H0, 224, 0, st>>>(img, offsets, step_ptr, off_step_stride, (float*)R->zz, R->gZZ, reps, Hr, Wr, xf, 0);
      else
        pack_crops16_kernel<__nv_bfloat16><<<N * H0, 224, 0, st>>>(img, offsets, step_ptr, off_step_stride, (__nv_bfloat16*)R->zz, R->gZZ,
                                                                   reps, Hr, Wr, xf, 0);
    } else if (R->dtype == 0) {
      pack_crops_kernel<float><<<N * H0, pack_threads, 0, st>>>(img, offsets, step_ptr, off_step_stride, (float*)R->zz, R->gZZ, reps, Hr, Wr,
                                                                xf, 0);
    } else {
      pack_crops_kernel<__nv_bfloat16><<<N * H0, pack_threads, 0, st>>>(img, offsets, step_ptr, off_step_stride, (__nv_bfloat16*)R->zz,
                                                                        R->gZZ, reps, Hr, Wr, xf, 0);
    }
    RGIE_LAUNCH_OK();
    if (int rc = run_op(R, R->fwd_ops[0], st, 0)) return rc;
    if (!R->stem_pool) {             // otherwise fwd_ops[0] was conv1 + max-pool: p1 and the argmax bytes are written
      if (R->dtype == 0)
        maxpool_fwd_kernel<float><<<N * H1, 256, 0, st>>>((const float*)R->c1, (float*)R->p1, R->arg, R->gS[1], H0, 64);
      else
        maxpool_fwd_kernel<__nv_bfloat16><<<N * H1, 256, 0, st>>>((const __nv_bfloat16*)R->c1, (__nv_bfloat16*)R->p1, R->arg, R->gS[1], H0, 64);
      RGIE_LAUNCH_OK();
    }
  }
  for (size_t i = 1; i < R->fwd_ops.size(); ++i)
    if (int rc = run_op(R, R->fwd_ops[i], st, i)) return rc;
  const Block& last = R->blocks.back();
  if (R->dtype == 0)
    avgpool_kernel<float><<<dim3(N, 2048 / (32 * 4)), 256, 0, st>>>((const float*)last.out, R->gS[4], 2048, R->feat);
  else
    avgpool_kernel<__nv_bfloat16><<<dim3(N, 2048 / (32 * 8)), 256, 0, st>>>((const __nv_bfloat16*)last.out, R->gS[4], 2048, R->feat);
  RGIE_LAUNCH_OK();
  fc_kernel<<<N, 256, 0, st>>>(R->feat, 2048, R->wfc, R->bfc, R->K, logits);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_regressor_forward(RgieRegressor* R, const float* img, int B, int Hr, int Wr, const int* offsets, int reps,
                           int normalize, float* logits, void* stream) {
  return rgie_regressor_forward_ex(R, img, B, Hr, Wr, offsets, nullptr, 0, reps, normalize, logits, stream);
}

int rgie_regressor_backward(RgieRegressor* R, const float* dlogits, float* dimg, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RGIE_CHECK(R && dlogits && dimg, "rgie_regressor_backward: null argument");
  RGIE_CHECK(R->offsets != nullptr, "rgie_regressor_backward: call rgie_regressor_forward first");
  const int N = R->N;
  const Block& last = R->blocks.back();
  const int H4 = R->Hs[4];
  fc_bwd_kernel<<<N, 256, 0, st>>>(dlogits, R->wfc, R->K, 2048, 1.0f / (float)(H4 * H4), R->dfeat);
  RGIE_LAUNCH_OK();
  {
    const long tot4 = (long)N * H4 * H4 * (2048 / (16 / R->esz));
    if (R->dtype == 0)
      avgpool_bwd_kernel<float><<<grid_for(tot4), 256, 0, st>>>(R->dfeat, (const float*)last.out, (float*)R->final_dout, R->gS[4], 2048, tot4);
    else
      avgpool_bwd_kernel<__nv_bfloat16><<<grid_for(tot4), 256, 0, st>>>(R->dfeat, (const __nv_bfloat16*)last.out, (__nv_bfloat16*)R->final_dout, R->gS[4], 2048, tot4);
    RGIE_LAUNCH_OK();
  }
  const size_t nb = R->bwd_ops.size();
  for (size_t i = 0; i + 1 < nb; ++i)
    if (int rc = run_op(R, R->bwd_ops[i], st, R->fwd_ops.size() + i)) return rc;
  // d(pool out) is the destination of the last block op (layer1.0 c1 dgrad)
  const void* dP = R->bwd_ops[nb - 2].d.D;
  {
    const int H1 = R->Hs[1];
    if (R->dtype == 0)
      maxpool_bwd_kernel<float><<<N * H1, 256, 0, st>>>((const float*)dP, R->arg, (float*)R->dC1, R->gS[1], R->gDY, 64);
    else
      maxpool_bwd_kernel<__nv_bfloat16><<<N * H1, 256, 0, st>>>((const __nv_bfloat16*)dP, R->arg, (__nv_bfloat16*)R->dC1, R->gS[1], R->gDY, 64);
    RGIE_LAUNCH_OK();
    if (int rc = run_op(R, R->bwd_ops[nb - 1], st, R->fwd_ops.size() + nb - 1)) return rc;
  }
  InXform xf = R->xf2;
  xf.mode = R->normalize;
  crop_grad_gather_kernel<<<R->B * R->Hr, 256, 0, st>>>(R->dZ, R->offsets, R->step_ptr, R->off_stride, dimg, R->img, R->B, R->reps,
                                                        R->Hr, R->Wr, R->crop, xf);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_regressor_set_input_transform(RgieRegressor* R, float pre_scale, float pre_shift, const float* mean3,
                                       const float* std3) {
  RGIE_CHECK(R && mean3 && std3, "rgie_regressor_set_input_transform: null");
  R->xf2.mode = 2; R->xf2.ps = pre_scale; R->xf2.pb = pre_shift;
  for (int c = 0; c < 3; ++c) { R->xf2.mean[c] = mean3[c]; R->xf2.std[c] = std3[c]; }
  return 0;
}

int rgie_regressor_set_profiling(RgieRegressor* R, int on) {
  RGIE_CHECK(R != nullptr, "rgie_regressor_set_profiling: null");
  if (on && R->ev.empty()) {
    R->ev.resize(2 * (R->fwd_ops.size() + R->bwd_ops.size()));
    for (auto& e : R->ev) RGIE_CUDA_OK(cudaEventCreate(&e));
  }
  R->profiling = on != 0;
  return 0;
}

int rgie_regressor_num_ops(const RgieRegressor* R) { return R ? (int)(R->fwd_ops.size() + R->bwd_ops.size()) : 0; }

int rgie_regressor_get_profile(RgieRegressor* R, float* h_ms, double* h_flops, double* h_bytes, int* h_info, int capacity,
                               int* n_out) {
  RGIE_CHECK(R && h_ms && h_flops && h_bytes && h_info && n_out, "rgie_regressor_get_profile: null");
  RGIE_CHECK(!R->ev.empty(), "rgie_regressor_get_profile: profiling was never enabled");
  const int n = (int)(R->fwd_ops.size() + R->bwd_ops.size());
  RGIE_CHECK(capacity >= n, "rgie_regressor_get_profile: capacity");
  for (int i = 0; i < n; ++i) {
    const bool fwd = i < (int)R->fwd_ops.size();
    const GemmOp& op = fwd ? R->fwd_ops[i] : R->bwd_ops[i - R->fwd_ops.size()];
    float ms = 0.f;
    RGIE_CUDA_OK(cudaEventSynchronize(R->ev[2 * i + 1]));
    RGIE_CUDA_OK(cudaEventElapsedTime(&ms, R->ev[2 * i], R->ev[2 * i + 1]));
    h_ms[i] = ms;
    // useful fraction of the padded stem formulations: conv1 fwd 147 of 4*64 K-elements; conv1 dgrad 147*? of 16*16*64
    double useful = 1.0;
    if (fwd && i == 0) useful = 147.0 / 256.0;
    if (!fwd && i == n - 1) useful = (147.0 * 64.0) / (16.0 * 16.0 * 64.0);
    h_flops[i] = op_flops(op.d, useful);
    h_bytes[i] = op_bytes(op.d, R->esz);
    if (fwd && i == 0 && R->stem_pool) {
      // conv1 + max-pool in one launch: the packed crops are read, the pooled tensor and one argmax byte per pooled element
      // are written; the 224 x 224 x 64 conv output never reaches HBM
      const double pooled = (double)R->N * R->Hs[1] * R->Hs[1] * 64.0;
      h_bytes[i] = (double)R->gZZ.rows() * 16.0 * R->esz + pooled * (R->esz + 1.0);
    }
    if (op.absorbed) { h_flops[i] = 0.0; h_bytes[i] = 0.0; h_ms[i] = 0.f; }
    if (op.fused_next) {
      // the fused launch does both ops; the second one's A operand never travels through HBM
      const GemmOp& nx = fwd ? R->fwd_ops[i + 1] : R->bwd_ops[i + 1 - R->fwd_ops.size()];
      const double valid_frac = (double)nx.d.src.H * nx.d.src.W / (double)nx.d.src.S;
      h_flops[i] += op_flops(nx.d, 1.0);
      h_bytes[i] += op_bytes(nx.d, R->esz) - (double)nx.d.a_rows * valid_frac * nx.d.Cin * R->esz;
    }
    h_info[4 * i + 0] = fwd ? 0 : 1;
    h_info[4 * i + 1] = op.d.Cout;
    h_info[4 * i + 2] = op.d.ntaps * op.d.Cin + (op.d.A2 ? op.d.Cin2 : 0);
    h_info[4 * i + 3] = (int)((op.d.m_end - op.d.m_begin + 127) / 128);
  }
  *n_out = n;
  return 0;
}

int rgie_regressor_tap(RgieRegressor* R, const char* name, float* out, long capacity, long* n_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RGIE_CHECK(R && name && out && n_out, "rgie_regressor_tap: null argument");
  const void* buf = nullptr;
  Geom g; int C = 0, fh = 0, fw = 0;
  std::string nm(name);
  if (nm == "feat") {
    long n = (long)R->N * 2048;
    RGIE_CHECK(n <= capacity, "rgie_regressor_tap: capacity");
    RGIE_CUDA_OK(cudaMemcpyAsync(out, R->feat, n * 4, cudaMemcpyDeviceToDevice, st));
    *n_out = n;
    return 0;
  }
  if (nm == "stem") {
    // with the fused conv1 + max-pool the stem activation does not exist: produce it on demand from the packed crops of the
    // last forward (diagnostics only)
    if (R->stem_pool) { if (int rc = run_gemm_sm100(R->conv1_plain, st)) return rc; }
    buf = R->c1; g = make_geom(1, R->N, R->H0, R->H0, 0, 0, 0, 0); C = 64; fh = fw = R->H0;
  }
  else if (nm == "pool") { buf = R->p1; g = R->gS[1]; C = 64; fh = fw = R->Hs[1]; }
  else {
    // "layer{s}.{i}" [".c1" | ".c2"]
    int s = 0, i = 0; char tail[8] = {0};
    int got = sscanf(name, "layer%d.%d.%7s", &s, &i, tail);
    RGIE_CHECK(got >= 2 && s >= 1 && s <= 4, "rgie_regressor_tap: unknown name");
    const int nblk[5] = {0, 3, 4, 6, 3};
    RGIE_CHECK(i >= 0 && i < nblk[s], "rgie_regressor_tap: block index");
    int bi = i;
    for (int q = 1; q < s; ++q) bi += nblk[q];
    const Block& k = R->blocks[bi];
    if (got == 2) { buf = k.out; g = k.gout; C = k.co; }
    else if (!strcmp(tail, "c1")) { buf = k.h1; g = k.gx; C = k.cm; }
    else if (!strcmp(tail, "c2")) { buf = k.h2; g = k.gs; C = k.cm; }
    else return fail("rgie_regressor_tap: unknown suffix");
    fh = fw = g.planes == 4 ? 2 * g.H : g.H;
  }
  const long total = (long)R->N * C * fh * fw;
  RGIE_CHECK(total <= capacity, "rgie_regressor_tap: capacity");
  if (R->dtype == 0) tap_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)buf, g, C, fh, fw, out, total);
  else tap_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)buf, g, C, fh, fw, out, total);
  RGIE_LAUNCH_OK();
  *n_out = total;
  return 0;
}

int rgie_gemm_selftest_ex(int backend, const void* A, long a_rows, int Cin, const void* A2, long a2_rows, int Cin2,
                          const void* W, int n_pad, int ntaps, const long* h_row_off, long m_begin, long m_end, int Cout,
                          const float* bias, const void* res, const unsigned* mask_bits, int relu, void* D, int d_fp32,
                          unsigned* D_bits, void* stream) {
  RGIE_CHECK(ntaps >= 1 && ntaps <= kMaxTaps, "rgie_gemm_selftest: ntaps");
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  d.A = A; d.a_rows = a_rows; d.Cin = Cin; d.Wt = W; d.n_pad = n_pad; d.ntaps = ntaps;
  d.A2 = A2; d.a2_rows = a2_rows; d.Cin2 = Cin2;
  for (int t = 0; t < ntaps; ++t) d.row_off[t] = h_row_off[t];
  d.m_begin = m_begin; d.m_end = m_end; d.Cout = Cout;
  // trivial geometry: one "image" of (m_end) x 1 pixels, nothing is padding
  d.src = make_geom(1, 1, (int)m_end, 1, 0, 0, 0, 0);
  d.dst_kind = DST_SAME; d.dst = d.src;
  d.D = D; d.ldd = Cout; d.d_fp32 = d_fp32; d.bias = bias;
  d.res = res; d.ld_res = Cout; d.res_rows = m_end; d.relu = relu;
  d.mask_bits = mask_bits; d.ld_mb = Cout / 32; d.D_bits = D_bits; d.ld_db = Cout / 32;
  if (backend == 1) return launch_gemm_sm100(d, (cudaStream_t)stream);
  return launch_gemm_simt(d, 1, (cudaStream_t)stream);
}

int rgie_gemm_selftest_fp32(int backend, const float* A, long a_rows, int Cin, int a_ld, const float* A2, long a2_rows,
                            int Cin2, const float* h_W, int ntaps, const long* h_row_off, long m_begin, long m_end, int Cout,
                            const float* bias, const float* res, const float* mask, int relu, float* D, void* stream) {
  RGIE_CHECK(ntaps >= 1 && ntaps <= kMaxTaps && h_W != nullptr, "rgie_gemm_selftest_fp32: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  const size_t ktot = (size_t)ntaps * Cin + (A2 ? Cin2 : 0);
  d.A = A; d.a_rows = a_rows; d.Cin = Cin; d.a_ld = a_ld; d.n_pad = Cout; d.ntaps = ntaps;
  d.A2 = A2; d.a2_rows = a2_rows; d.Cin2 = Cin2;
  for (int t = 0; t < ntaps; ++t) d.row_off[t] = h_row_off[t];
  d.m_begin = m_begin; d.m_end = m_end; d.Cout = Cout;
  d.src = make_geom(1, 1, (int)m_end, 1, 0, 0, 0, 0);
  d.dst_kind = DST_SAME; d.dst = d.src;
  d.D = D; d.ldd = Cout; d.d_fp32 = 1; d.bias = bias;
  d.res = res; d.ld_res = Cout; d.res_rows = m_end; d.relu = relu;
  d.mask = mask; d.ld_mask = Cout;
  void* w_dev = nullptr;
  int rc = 0;
  if (backend == 2) {
    std::vector<__nv_bfloat16> planes(3 * (size_t)Cout * ktot);
    split_weights_bf16x3(h_W, (size_t)Cout * ktot, planes.data());
    RGIE_CUDA_OK(cudaMalloc(&w_dev, planes.size() * 2));
    RGIE_CUDA_OK(cudaMemcpy(w_dev, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice));
    d.Wt = w_dev;
    GemmPlanTc32 plan;
    rc = build_gemm_tc32(d, &plan);
    if (!rc) rc = run_gemm_tc32(plan, st);
  } else {
    RGIE_CUDA_OK(cudaMalloc(&w_dev, (size_t)Cout * ktot * 4));
    RGIE_CUDA_OK(cudaMemcpy(w_dev, h_W, (size_t)Cout * ktot * 4, cudaMemcpyHostToDevice));
    d.Wt = w_dev;
    rc = launch_gemm_simt(d, 0, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(w_dev);
  if (rc) return rc;
  RGIE_CUDA_OK(e);
  return 0;
}

int rgie_gemm_selftest(int backend, const void* A, long a_rows, int Cin, const void* W, int n_pad, int ntaps,
                       const long* h_row_off, long m_begin, long m_end, int Cout, const float* bias, const void* res,
                       int relu, void* D, int d_fp32, void* stream) {
  return rgie_gemm_selftest_ex(backend, A, a_rows, Cin, nullptr, 0, 0, W, n_pad, ntaps, h_row_off, m_begin, m_end, Cout, bias,
                               res, nullptr, relu, D, d_fp32, nullptr, stream);
}

}  // extern "C"

// Parametric photo-filter kernels (forward + backward) for the default 8-filter chain of the reference
// (src/baselines/image_transformations/image_transformations.py:7-66, img_trans_torch_diff.py, kornia 0.8.2 calls).
//
// Layout: images NCHW fp32 contiguous [B,3,H,W] in [0,1]; every filter is followed by clamp(0,1) exactly as
// apply_params does (:60).  Parameters are the EFFECTIVE per-image values (after the reference's own clamps, which live
// in the host mirror / the param-transform kernel): `p` points at image 0's values, `p_stride` floats between images
// (0 = shared by the batch).  Backward kernels produce d(in) and per-image d(param) through a deterministic two-stage
// reduction (warp shuffle -> block partials in `ws` -> fixed-order finalize), no atomics.
//
// All kernels are HBM-bound elementwise/stencil passes: float4-vectorised over pixels when H*W % 4 == 0.
#include <stdlib.h>
#include "common.cuh"
#include "rgie.h"

namespace rgie {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlk = 128;       // blocks per image along x (partials per image)
constexpr float kTwoPi = 6.283185307179586f;

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ bool in01(float v) { return v >= 0.f && v <= 1.f; }   // torch clamp backward: inclusive

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-reduce NP per-thread accumulators, thread 0 writes them to dst[0..NP)
template <int NP>
__device__ __forceinline__ void block_reduce_store(float* acc, float* dst) {
  __shared__ float red[kThreads / 32][NP > 0 ? NP : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float v = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NP) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

// finalize: out[b*stride + i] = sum_k partial[(b*nblk + k)*NP + i]
__global__ void finalize_partials(const float* __restrict__ partial, int nblk, int NP, float* __restrict__ out, int stride) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < NP; i += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nblk; ++k) s += partial[((long)b * nblk + k) * NP + i];
    out[(long)b * stride + i] = s;
  }
}

__global__ void copy_strided_kernel(const float* __restrict__ src, int src_stride, float* __restrict__ dst,
                                    int dst_stride, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[(long)i * dst_stride] = src[(long)i * src_stride];
}

struct LaunchShape {
  int nblk;    // blocks per image
  int chunk;   // pixels per block (multiple of 4)
};
// Forward passes: 2048 pixels per block.  Backward passes end in a block reduction of up to 24 parameter gradients
// (5 shuffles each + shared memory), which costs as much as ~25 pixels of arithmetic per thread: they take 8192 pixels per
// block (32 per thread) unless that would leave the GPU with fewer than ~4 blocks per SM.
LaunchShape shape_for(int HW, int B = 1, bool bwd = false) {
  LaunchShape s;
  s.nblk = ceil_div(HW, 2048);
  if (bwd) {
    const int coarse = ceil_div(HW, 8192);
    if ((long)coarse * B >= 592) s.nblk = coarse;
  }
  if (s.nblk > kMaxBlk) s.nblk = kMaxBlk;
  if (s.nblk < 1) s.nblk = 1;
  s.chunk = ceil_div(ceil_div(HW, s.nblk), 4) * 4;
  return s;
}

// ===============================================================================================================
// Pointwise ops (one RGB pixel in, one RGB pixel out)
// ===============================================================================================================
struct ExposureOp {            // F1: img_trans_torch_diff.py:60-64
  static constexpr int NP = 1;
  const float* p; int stride; float s;
  __device__ void load(int b) { s = expf(__fmul_rn(p[(long)b * stride], logf(2.0f))); }
  __device__ void fwd(const float* x, float* y) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = clamp01(__fmul_rn(x[c], s));
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float pre = __fmul_rn(x[c], s);
      float gg = in01(pre) ? g[c] : 0.f;
      gx[c] = gg * s;
      gp[0] += gg * x[c] * s * 0.6931471805599453f;
    }
  }
};

struct SaturationOp {          // F2: kornia.enhance.adjust_saturation (rgb_to_hsv -> s*=f, clamp -> hsv_to_rgb)
  static constexpr int NP = 1;
  const float* p; int stride; float F;
  __device__ void load(int b) { F = p[(long)b * stride]; }

  struct Mid { float M, mn, delta, dc, hnum, hsel, s, v, s2pre, s2, f; int a, imin, hi; };

  __device__ void forward_mid(const float* x, Mid& m, float* y) const {
    const float r = x[0], g = x[1], b = x[2];
    m.a = (r >= g && r >= b) ? 0 : (g >= b ? 1 : 2);              // first max (torch CPU tie rule)
    m.imin = (r <= g && r <= b) ? 0 : (g <= b ? 1 : 2);           // first min
    m.M = fmaxf(r, fmaxf(g, b));
    m.mn = fminf(r, fminf(g, b));
    m.delta = m.M - m.mn;
    m.v = m.M;
    m.s = m.delta / (m.M + 1e-8f);
    m.dc = (m.delta == 0.f) ? 1.f : m.delta;
    const float rc = m.M - r, gc = m.M - g, bc = m.M - b;
    m.hnum = (m.a == 0) ? (bc - gc) : (m.a == 1 ? __fadd_rn(rc - bc, __fmul_rn(2.0f, m.dc))
                                                 : __fadd_rn(gc - rc, __fmul_rn(4.0f, m.dc)));
    m.hsel = m.hnum / m.dc;
    // (hsel / 6) % 1 (python modulo): |hsel / 6| <= 5/6 < 1, so the remainder is the value itself -- no fmodf needed
    float hm = m.hsel / 6.0f;
    if (hm < 0.f) hm += 1.0f;
    const float h = __fmul_rn(kTwoPi, hm);
    m.s2pre = __fmul_rn(m.s, F);
    m.s2 = clamp01(m.s2pre);
    // hsv_to_rgb
    const float hn = h / kTwoPi;
    const float h6 = __fmul_rn(hn, 6.0f);
    // floor(h6) % 6 and h6 % 6 with h6 in [0, 6]: exact for one wrap (x - 6 is exact for x in [6, 12))
    const float fl = floorf(h6);
    const float him = fl >= 6.0f ? fl - 6.0f : fl;
    const float h6m = h6 >= 6.0f ? h6 - 6.0f : h6;
    m.f = h6m - him;
    m.hi = (int)him;
    const float v = m.v, s2 = m.s2, f = m.f;
    const float pp = __fmul_rn(v, 1.0f - s2);
    const float qq = __fmul_rn(v, 1.0f - __fmul_rn(f, s2));
    const float tt = __fmul_rn(v, 1.0f - __fmul_rn(1.0f - f, s2));
    switch (m.hi) {
      case 0: y[0] = v; y[1] = tt; y[2] = pp; break;
      case 1: y[0] = qq; y[1] = v; y[2] = pp; break;
      case 2: y[0] = pp; y[1] = v; y[2] = tt; break;
      case 3: y[0] = pp; y[1] = qq; y[2] = v; break;
      case 4: y[0] = tt; y[1] = pp; y[2] = v; break;
      default: y[0] = v; y[1] = pp; y[2] = qq; break;
    }
  }
  __device__ void fwd(const float* x, float* y) const {
    Mid m;
    forward_mid(x, m, y);
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = clamp01(y[c]);
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
    Mid m;
    float y[3];
    forward_mid(x, m, y);
    float go[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) go[c] = in01(y[c]) ? g[c] : 0.f;
    float g_v = 0.f, g_p = 0.f, g_q = 0.f, g_t = 0.f;
    switch (m.hi) {
      case 0: g_v = go[0]; g_t = go[1]; g_p = go[2]; break;
      case 1: g_q = go[0]; g_v = go[1]; g_p = go[2]; break;
      case 2: g_p = go[0]; g_v = go[1]; g_t = go[2]; break;
      case 3: g_p = go[0]; g_q = go[1]; g_v = go[2]; break;
      case 4: g_t = go[0]; g_p = go[1]; g_v = go[2]; break;
      default: g_v = go[0]; g_p = go[1]; g_q = go[2]; break;
    }
    const float v = m.v, s2 = m.s2, f = m.f;
    float gv = g_v + g_p * (1.0f - s2) + g_q * (1.0f - f * s2) + g_t * (1.0f - (1.0f - f) * s2);
    float gs2 = -v * (g_p + f * g_q + (1.0f - f) * g_t);
    float gf = v * s2 * (g_t - g_q);
    // f <- h6 <- hn <- h <- hmod <- hsel   (floor / integer parts carry no gradient)
    float ghsel = (((gf * 6.0f) / kTwoPi) * kTwoPi) / 6.0f;
    float gpre = in01(m.s2pre) ? gs2 : 0.f;
    float gs = gpre * F;
    gp[0] += gpre * m.s;
    float ghnum = ghsel / m.dc;
    float gdc = -ghsel * m.hsel / m.dc;
    float grc = 0.f, ggc = 0.f, gbc = 0.f;
    if (m.a == 0) { gbc += ghnum; ggc -= ghnum; }
    else if (m.a == 1) { grc += ghnum; gbc -= ghnum; gdc += 2.0f * ghnum; }
    else { ggc += ghnum; grc -= ghnum; gdc += 4.0f * ghnum; }
    float gM = grc + ggc + gbc;
    gx[0] = -grc; gx[1] = -ggc; gx[2] = -gbc;
    float gdelta = (m.delta != 0.f) ? gdc : 0.f;
    const float Me = m.M + 1e-8f;
    gdelta += gs / Me;
    gM += -gs * m.s / Me;
    gM += gv;
    gM += gdelta;
    const float gmn = -gdelta;
    gx[m.a] += gM;
    gx[m.imin] += gmn;
  }
};

template <int NCURVE>          // F3 (tone: NCURVE=1, shared by RGB) / F4 (color: NCURVE=3): img_trans_torch_diff.py:6-19
struct CurveOp {
  static constexpr int NP = 8 * NCURVE;
  const float* p; int stride; float w[NP];
  __device__ void load(int b) {
#pragma unroll
    for (int i = 0; i < NP; ++i) w[i] = p[(long)b * stride + i];
  }
  __device__ void fwd(const float* x, float* y) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* wc = w + (NCURVE == 3 ? 8 * c : 0);
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        tot = __fadd_rn(tot, __fmul_rn(fminf(fmaxf(x[c] - 0.125f * i, 0.f), 0.125f), wc[i]));
      y[c] = clamp01(fminf(tot, 1.0f));
    }
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* wc = w + (NCURVE == 3 ? 8 * c : 0);
      float* gpc = gp + (NCURVE == 3 ? 8 * c : 0);
      float tot = 0.f, seg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        seg[i] = fminf(fmaxf(x[c] - 0.125f * i, 0.f), 0.125f);
        tot = __fadd_rn(tot, __fmul_rn(seg[i], wc[i]));
      }
      const float gg = in01(tot) ? g[c] : 0.f;    // clamp(max=1) then clamp(0,1): passes iff 0 <= tot <= 1
      float dx = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        gpc[i] += gg * seg[i];
        const float u = x[c] - 0.125f * i;
        if (u >= 0.f && u <= 0.125f) dx += wc[i];
      }
      gx[c] = gg * dx;
    }
  }
};

struct ContrastOp {            // F5: kornia adjust_contrast_with_mean_subtraction (mean computed by gray_mean_kernel)
  static constexpr int NP = 2;                  // [0]: sum g*(x-mu) = d/df ; [1]: sum g = G (mean path)
  const float* p; int stride; const float* mu_all; float f, mu;
  __device__ void load(int b) { f = p[(long)b * stride]; mu = mu_all[b]; }
  __device__ void fwd(const float* x, float* y) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = clamp01(__fadd_rn(__fmul_rn(x[c], f), __fmul_rn(mu, 1.0f - f)));
  }
  // first backward pass: masked gradient gm (written in place of gx), reductions for df and the mean path
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float pre = __fadd_rn(__fmul_rn(x[c], f), __fmul_rn(mu, 1.0f - f));
      const float gg = in01(pre) ? g[c] : 0.f;
      gx[c] = gg * f;
      gp[0] += gg * (x[c] - mu);
      gp[1] += gg;
    }
  }
};

// exposure -> saturation -> tone -> colour as ONE pixel op (the head of the reference's default filter list,
// optimize_image_param.py:227): the three intermediate images never exist.  Each stage is the op above, clamp included,
// so the result is bit-identical to running them one after the other; backward recomputes the three intermediates per
// pixel and chains the four backward functions.  Parameters: 34 consecutive floats (1 + 1 + 8 + 24).
struct PrefixOp {
  static constexpr int NP = 34;
  ExposureOp e; SaturationOp s; CurveOp<1> t; CurveOp<3> c;
  __device__ void load(int b) { e.load(b); s.load(b); t.load(b); c.load(b); }
  __device__ void fwd(const float* x, float* y) const {
    float x1[3], x2[3], x3[3];
    e.fwd(x, x1); s.fwd(x1, x2); t.fwd(x2, x3); c.fwd(x3, y);
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
    float x1[3], x2[3], x3[3], g3[3], g2[3], g1[3];
    e.fwd(x, x1); s.fwd(x1, x2); t.fwd(x2, x3);
    c.bwd(x3, g, g3, gp + 10);
    t.bwd(x2, g3, g2, gp + 2);
    s.bwd(x1, g2, g1, gp + 1);
    e.bwd(x, g1, gx, gp);
  }
};
PrefixOp make_prefix_op(const float* p, int stride) {
  PrefixOp op;
  op.e = ExposureOp{p, stride, 0.f};
  op.s = SaturationOp{p + 1, stride, 0.f};
  op.t.p = p + 2; op.t.stride = stride;
  op.c.p = p + 10; op.c.stride = stride;
  return op;
}

// ---------------------------------------------------------------------------------------------------------------
// the remaining pointwise filters of apply_params
// ---------------------------------------------------------------------------------------------------------------
struct GammaOp {               // kornia.enhance.adjust_gamma(im, clamp(gamma, min=0), gain=1): clamp(x^gamma, 0, 1)
  static constexpr int NP = 1;                  // image_transformations.py:176-185
  const float* p; int stride; float gam;
  __device__ void load(int b) { gam = p[(long)b * stride]; }
  __device__ void fwd(const float* x, float* y) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = clamp01(powf(x[c], gam));
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float pre = powf(x[c], gam);
      const float gg = in01(pre) ? g[c] : 0.f;
      // torch pow_backward: d/dx = gamma * x^(gamma-1) (0 when gamma == 0); d/dgamma = x^gamma * ln x (0 at x == 0, gamma >= 0)
      gx[c] = gam == 0.f ? 0.f : gg * gam * powf(x[c], gam - 1.0f);
      gp[0] += (x[c] == 0.f && gam >= 0.f) ? 0.f : gg * pre * logf(x[c]);
    }
  }
};

struct BrightOp {              // kornia.enhance.adjust_brightness(im, clamp(p, 0, 1), clip_output=True): clamp(x + p, 0, 1)
  static constexpr int NP = 1;                  // image_transformations.py:136-143
  const float* p; int stride; float f;
  __device__ void load(int b) { f = p[(long)b * stride]; }
  __device__ void fwd(const float* x, float* y) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = clamp01(__fadd_rn(x[c], f));
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gg = in01(__fadd_rn(x[c], f)) ? g[c] : 0.f;
      gx[c] = gg;
      gp[0] += gg;
    }
  }
};

struct BwOp {                  // img_trans_torch_diff.py:67-70: lerp(im, rgb2lum(im), p); lum = 0.27 r + 0.67 g + 0.06 b
  static constexpr int NP = 1;                  // (color_transformations.py:74-81); then apply_params' clamp(0, 1)
  const float* p; int stride; float f;
  __device__ void load(int b) { f = p[(long)b * stride]; }
  __device__ float lum(const float* x) const {
    return __fadd_rn(__fadd_rn(__fmul_rn(0.27f, x[0]), __fmul_rn(0.67f, x[1])), __fmul_rn(0.06f, x[2]));
  }
  __device__ void fwd(const float* x, float* y) const {
    const float l = lum(x);
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = clamp01(__fadd_rn(__fmul_rn(1.0f - f, x[c]), __fmul_rn(f, l)));
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
    const float l = lum(x);
    const float wc[3] = {0.27f, 0.67f, 0.06f};
    float gg[3], gl = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float pre = __fadd_rn(__fmul_rn(1.0f - f, x[c]), __fmul_rn(f, l));
      gg[c] = in01(pre) ? g[c] : 0.f;
      gl += gg[c];
      gp[0] += gg[c] * (l - x[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) gx[c] = gg[c] * (1.0f - f) + gl * f * wc[c];
  }
};

struct HueOp {                 // kornia.enhance.adjust_hue(im, clamp(p, -pi, pi)): rgb_to_hsv -> h = fmod(h + p, 2 pi) -> hsv_to_rgb
  static constexpr int NP = 1;                  // image_transformations.py:166-173
  const float* p; int stride; float F;
  __device__ void load(int b) { F = p[(long)b * stride]; }

  struct Mid { float M, mn, delta, dc, hnum, hsel, s, v, f; int a, imin, hi; };

  __device__ void forward_mid(const float* x, Mid& m, float* y) const {
    const float r = x[0], g = x[1], b = x[2];
    m.a = (r >= g && r >= b) ? 0 : (g >= b ? 1 : 2);              // first max (torch CPU tie rule)
    m.imin = (r <= g && r <= b) ? 0 : (g <= b ? 1 : 2);           // first min
    m.M = fmaxf(r, fmaxf(g, b));
    m.mn = fminf(r, fminf(g, b));
    m.delta = m.M - m.mn;
    m.v = m.M;
    m.s = m.delta / (m.M + 1e-8f);
    m.dc = (m.delta == 0.f) ? 1.f : m.delta;
    const float rc = m.M - r, gc = m.M - g, bc = m.M - b;
    m.hnum = (m.a == 0) ? (bc - gc) : (m.a == 1 ? __fadd_rn(rc - bc, __fmul_rn(2.0f, m.dc))
                                                 : __fadd_rn(gc - rc, __fmul_rn(4.0f, m.dc)));
    m.hsel = m.hnum / m.dc;
    float hm = fmodf(m.hsel / 6.0f, 1.0f);
    if (hm != 0.f && hm < 0.f) hm += 1.0f;
    const float h = fmodf(__fadd_rn(__fmul_rn(kTwoPi, hm), F), kTwoPi);   // torch.fmod: sign of the dividend
    // hsv_to_rgb
    const float hn = h / kTwoPi;
    const float h6 = __fmul_rn(hn, 6.0f);
    float fl = floorf(h6);
    float him = fmodf(fl, 6.0f);
    if (him != 0.f && him < 0.f) him += 6.0f;
    float h6m = fmodf(h6, 6.0f);
    if (h6m != 0.f && h6m < 0.f) h6m += 6.0f;
    m.f = h6m - him;
    m.hi = (int)him;
    const float v = m.v, s2 = m.s, f = m.f;
    const float pp = __fmul_rn(v, 1.0f - s2);
    const float qq = __fmul_rn(v, 1.0f - __fmul_rn(f, s2));
    const float tt = __fmul_rn(v, 1.0f - __fmul_rn(1.0f - f, s2));
    switch (m.hi) {
      case 0: y[0] = v; y[1] = tt; y[2] = pp; break;
      case 1: y[0] = qq; y[1] = v; y[2] = pp; break;
      case 2: y[0] = pp; y[1] = v; y[2] = tt; break;
      case 3: y[0] = pp; y[1] = qq; y[2] = v; break;
      case 4: y[0] = tt; y[1] = pp; y[2] = v; break;
      default: y[0] = v; y[1] = pp; y[2] = qq; break;
    }
  }
  __device__ void fwd(const float* x, float* y) const {
    Mid m;
    forward_mid(x, m, y);
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = clamp01(y[c]);
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
    Mid m;
    float y[3];
    forward_mid(x, m, y);
    float go[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) go[c] = in01(y[c]) ? g[c] : 0.f;
    float g_v = 0.f, g_p = 0.f, g_q = 0.f, g_t = 0.f;
    switch (m.hi) {
      case 0: g_v = go[0]; g_t = go[1]; g_p = go[2]; break;
      case 1: g_q = go[0]; g_v = go[1]; g_p = go[2]; break;
      case 2: g_p = go[0]; g_v = go[1]; g_t = go[2]; break;
      case 3: g_p = go[0]; g_q = go[1]; g_v = go[2]; break;
      case 4: g_t = go[0]; g_p = go[1]; g_v = go[2]; break;
      default: g_v = go[0]; g_p = go[1]; g_q = go[2]; break;
    }
    const float v = m.v, s2 = m.s, f = m.f;
    float gv = g_v + g_p * (1.0f - s2) + g_q * (1.0f - f * s2) + g_t * (1.0f - (1.0f - f) * s2);
    const float gs = -v * (g_p + f * g_q + (1.0f - f) * g_t);
    const float gf = v * s2 * (g_t - g_q);
    // f <- h6 = 6 h / 2pi,  h = fmod(h0 + F, 2pi),  h0 = 2pi * fmod(hsel / 6, 1)
    const float gh = gf * 6.0f / kTwoPi;
    gp[0] += gh;
    const float ghsel = gh * kTwoPi / 6.0f;
    float ghnum = ghsel / m.dc;
    float gdc = -ghsel * m.hsel / m.dc;
    float grc = 0.f, ggc = 0.f, gbc = 0.f;
    if (m.a == 0) { gbc += ghnum; ggc -= ghnum; }
    else if (m.a == 1) { grc += ghnum; gbc -= ghnum; gdc += 2.0f * ghnum; }
    else { ggc += ghnum; grc -= ghnum; gdc += 4.0f * ghnum; }
    float gM = grc + ggc + gbc;
    gx[0] = -grc; gx[1] = -ggc; gx[2] = -gbc;
    float gdelta = (m.delta != 0.f) ? gdc : 0.f;
    const float Me = m.M + 1e-8f;
    gdelta += gs / Me;
    gM += -gs * m.s / Me;
    gM += gv;
    gM += gdelta;
    const float gmn = -gdelta;
    gx[m.a] += gM;
    gx[m.imin] += gmn;
  }
};

struct WbOp {                  // img_trans_torch_diff.py:51-57: clamp(lerp(im, im * 0.5 / (mean_HW(im) + 1e-9), p), 0, 1)
  static constexpr int NP = 4;                  // [0]: d/dp ; [1..3]: sum_px gg * x per channel (mean path)
  const float* p; int stride; const float* mean_all; float f, bal[3];
  __device__ void load(int b) {
    f = p[(long)b * stride];
#pragma unroll
    for (int c = 0; c < 3; ++c) bal[c] = 0.5f / mean_all[3 * b + c];
  }
  __device__ void fwd(const float* x, float* y) const {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      y[c] = clamp01(__fadd_rn(__fmul_rn(1.0f - f, x[c]), __fmul_rn(f, __fmul_rn(x[c], bal[c]))));
  }
  __device__ void bwd(const float* x, const float* g, float* gx, float* gp) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float wb = __fmul_rn(x[c], bal[c]);
      const float pre = __fadd_rn(__fmul_rn(1.0f - f, x[c]), __fmul_rn(f, wb));
      const float gg = in01(pre) ? g[c] : 0.f;
      gx[c] = gg * ((1.0f - f) + f * bal[c]);
      gp[0] += gg * (wb - x[c]);
      gp[1 + c] += gg * x[c];
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
template <class Op, int VEC, bool BWD, bool WRITE = true, int MINB = (BWD ? 3 : 4)>
__global__ void __launch_bounds__(kThreads, MINB) pointwise_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                            float* __restrict__ out, Op op, float* __restrict__ partial,
                                                            int HW, int chunk) {
  const int b = blockIdx.y;
  op.load(b);
  const long base = (long)b * 3 * HW;
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, HW);
  float acc[Op::NP > 0 ? Op::NP : 1];
#pragma unroll
  for (int i = 0; i < Op::NP; ++i) acc[i] = 0.f;
  for (int px = p0 + threadIdx.x * VEC; px < p1; px += kThreads * VEC) {
    float x[3][VEC], g[3][VEC], y[3][VEC];
    if (VEC == 4) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float4 v = *reinterpret_cast<const float4*>(in + base + (long)c * HW + px);
        x[c][0] = v.x; x[c][1] = v.y; x[c][2] = v.z; x[c][3] = v.w;
        if (BWD) {
          float4 gv = *reinterpret_cast<const float4*>(gout + base + (long)c * HW + px);
          g[c][0] = gv.x; g[c][1] = gv.y; g[c][2] = gv.z; g[c][3] = gv.w;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        x[c][0] = in[base + (long)c * HW + px];
        if (BWD) g[c][0] = gout[base + (long)c * HW + px];
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float xi[3] = {x[0][v], x[1][v], x[2][v]}, yo[3];
      if (BWD) {
        float gi[3] = {g[0][v], g[1][v], g[2][v]};
        op.bwd(xi, gi, yo, acc);
      } else {
        op.fwd(xi, yo);
      }
      y[0][v] = yo[0]; y[1][v] = yo[1]; y[2][v] = yo[2];
    }
    if (!WRITE) continue;                      // first stage of a chain: nothing upstream consumes d(in)
    if (VEC == 4) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        *reinterpret_cast<float4*>(out + base + (long)c * HW + px) = make_float4(y[c][0], y[c][1], y[c][2], y[c][3]);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) out[base + (long)c * HW + px] = y[c][0];
    }
  }
  if (BWD && Op::NP > 0) block_reduce_store<Op::NP>(acc, partial + ((long)b * gridDim.x + blockIdx.x) * Op::NP);
}

template <class Op, bool BWD>
int launch_pointwise(const float* in, const float* gout, float* out, Op op, float* partial, int B, int HW,
                     cudaStream_t st) {
  LaunchShape s = shape_for(HW, B, BWD);
  dim3 grid(s.nblk, B);
  if (HW % 4 == 0) pointwise_kernel<Op, 4, BWD><<<grid, kThreads, 0, st>>>(in, gout, out, op, partial, HW, s.chunk);
  else pointwise_kernel<Op, 1, BWD><<<grid, kThreads, 0, st>>>(in, gout, out, op, partial, HW, s.chunk);
  RGIE_LAUNCH_OK();
  return 0;
}

// gray mean (contrast): partial sums of 0.299r + 0.587g + 0.114b
template <int VEC>
__global__ void __launch_bounds__(kThreads) gray_sum_kernel(const float* __restrict__ in, float* __restrict__ partial,
                                                           int HW, int chunk) {
  const int b = blockIdx.y;
  const long base = (long)b * 3 * HW;
  const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, HW);
  float acc[1] = {0.f};
  for (int px = p0 + threadIdx.x * VEC; px < p1; px += kThreads * VEC) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float r = in[base + px + v], g = in[base + HW + px + v], bl = in[base + 2L * HW + px + v];
      acc[0] += __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, bl));
    }
  }
  block_reduce_store<1>(acc, partial + (long)b * gridDim.x + blockIdx.x);
}
__global__ void gray_mean_finalize(const float* __restrict__ partial, int nblk, float inv_hw, float* __restrict__ mu) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < nblk; ++k) s += partial[(long)b * nblk + k];
    mu[b] = s * inv_hw;
  }
}

// contrast backward, second pass: gin = gm (already g*mask*f) + w_c * (1-f) * G / HW
__global__ void __launch_bounds__(kThreads) contrast_bwd_mean_kernel(float* __restrict__ gin, const float* __restrict__ p,
                                                                    int stride, const float* __restrict__ sums,
                                                                    int sum_stride, int HW) {
  const int b = blockIdx.y;
  const float f = p[(long)b * stride];
  const float G = sums[(long)b * sum_stride + 1];
  const float k = (1.0f - f) * G / (float)HW;
  const float wc[3] = {0.299f, 0.587f, 0.114f};
  const long base = (long)b * 3 * HW;
  if ((HW & 3) == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4* g4 = reinterpret_cast<float4*>(gin + base + (long)c * HW);
      const float add = wc[c] * k;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW / 4; i += gridDim.x * blockDim.x) {
        float4 v = g4[i];
        v.x += add; v.y += add; v.z += add; v.w += add;
        g4[i] = v;
      }
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * HW; i += gridDim.x * blockDim.x) gin[base + i] += wc[i / HW] * k;
}

// white balance: per-channel sums (forward pre-pass), mean finalize (+1e-9), backward mean path
__global__ void __launch_bounds__(kThreads) chan_sum_kernel(const float* __restrict__ in, float* __restrict__ partial,
                                                           int HW, int chunk) {
  const int b = blockIdx.y;
  const long base = (long)b * 3 * HW;
  const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, HW);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int px = p0 + threadIdx.x; px < p1; px += kThreads) {
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += in[base + (long)c * HW + px];
  }
  block_reduce_store<3>(acc, partial + ((long)b * gridDim.x + blockIdx.x) * 3);
}
__global__ void chan_mean_finalize(const float* __restrict__ partial, int nblk, float inv_hw, float* __restrict__ mean) {
  const int b = blockIdx.x;
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int k = 0; k < nblk; ++k) s += partial[((long)b * nblk + k) * 3 + threadIdx.x];
    mean[3 * b + threadIdx.x] = s * inv_hw + 1e-9f;
  }
}
// gin[c] += d(loss)/d(mean_c) / HW with d/d(mean_c) = -p * (0.5 / mean_c^2) * sum_px gg * x_c
__global__ void __launch_bounds__(kThreads) wb_bwd_mean_kernel(float* __restrict__ gin, const float* __restrict__ p, int stride,
                                                              const float* __restrict__ mean, const float* __restrict__ sums,
                                                              int HW) {
  const int b = blockIdx.y;
  const float f = p[(long)b * stride];
  float k[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float mc = mean[3 * b + c];
    k[c] = -f * (0.5f / (mc * mc)) * sums[4 * b + 1 + c] / (float)HW;
  }
  const long base = (long)b * 3 * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * HW; i += gridDim.x * blockDim.x) gin[base + i] += k[i / HW];
}

// ===============================================================================================================
// F6 sharpness: kornia.enhance.sharpness -- 3x3 [[1,1,1],[1,5,1],[1,1,1]]/13 'valid' smoothing, border keeps input,
//               _blend_one(degenerate, input, factor) with its ==0 / ==1 / (0,1) / else branches
// ===============================================================================================================
__device__ __forceinline__ float sharp_conv(const float* __restrict__ pl, int y, int x, int W) {
  const float k1 = 1.0f / 13.0f, k5 = 5.0f / 13.0f;
  const float* r0 = pl + (long)(y - 1) * W + x;
  const float* r1 = r0 + W;
  const float* r2 = r1 + W;
  float s = 0.f;
  s = fmaf(k1, r0[-1], s); s = fmaf(k1, r0[0], s); s = fmaf(k1, r0[1], s);
  s = fmaf(k1, r1[-1], s); s = fmaf(k5, r1[0], s); s = fmaf(k1, r1[1], s);
  s = fmaf(k1, r2[-1], s); s = fmaf(k1, r2[0], s); s = fmaf(k1, r2[1], s);
  return s;
}

// mode: 0 -> factor==0 (result), 1 -> factor==1 (input), 2 -> 0<f<1 (no inner clamp), 3 -> clamp
__device__ __forceinline__ int sharp_mode(float f) { return f == 0.f ? 0 : (f == 1.f ? 1 : ((f > 0.f && f < 1.f) ? 2 : 3)); }

__global__ void __launch_bounds__(kThreads) sharp_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            const float* __restrict__ p, int stride, int H, int W) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float f = p[(long)b * stride];
  const int mode = sharp_mode(f);
  const float* pl = in + ((long)b * 3 + c) * H * W;
  float* po = out + ((long)b * 3 + c) * H * W;
  for (int y = blockIdx.x * ((H + gridDim.x - 1) / gridDim.x); y < min(H, (int)(blockIdx.x + 1) * (int)((H + gridDim.x - 1) / gridDim.x)); ++y)   // contiguous rows per block: vertical reuse stays in L1
  for (int x = threadIdx.x; x < W; x += blockDim.x) {       // row loop: no per-pixel division
    const int i = y * W + x;
    const float xin = pl[i];
    float result = xin;
    if (y >= 1 && y < H - 1 && x >= 1 && x < W - 1) result = clamp01(sharp_conv(pl, y, x, W));
    float o;
    if (mode == 0) o = result;
    else if (mode == 1) o = xin;
    else {
      o = __fadd_rn(result, __fmul_rn(xin - result, f));
      if (mode == 3) o = clamp01(o);
    }
    po[i] = clamp01(o);
  }
}

// backward pass A: gdeg (gradient w.r.t. the clamped 3x3 smoothing, zero on the border) and the direct d(in) term;
// partial sums of d(factor)
__global__ void __launch_bounds__(kThreads) sharp_bwd_a_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                              float* __restrict__ gdeg, float* __restrict__ gin,
                                                              const float* __restrict__ p, int stride,
                                                              float* __restrict__ partial, int H, int W) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float f = p[(long)b * stride];
  const int mode = sharp_mode(f);
  const long off = ((long)b * 3 + c) * H * W;
  const float* pl = in + off;
  float acc[1] = {0.f};
  for (int y = blockIdx.x * ((H + gridDim.x - 1) / gridDim.x); y < min(H, (int)(blockIdx.x + 1) * (int)((H + gridDim.x - 1) / gridDim.x)); ++y)   // contiguous rows per block: vertical reuse stays in L1
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const int i = y * W + x;
    const float xin = pl[i];
    const bool interior = y >= 1 && y < H - 1 && x >= 1 && x < W - 1;
    float conv = 0.f, result = xin;
    if (interior) { conv = sharp_conv(pl, y, x, W); result = clamp01(conv); }
    float g = gout[off + i];
    float g_res = 0.f, g_x = 0.f;
    if (mode == 0) { g = in01(result) ? g : 0.f; g_res = g; }
    else if (mode == 1) { g = in01(xin) ? g : 0.f; g_x = g; }
    else {
      const float o = __fadd_rn(result, __fmul_rn(xin - result, f));
      g = in01(o) ? g : 0.f;       // inner clamp (mode 3) and outer clamp have the same range
      g_res = g - g * f;
      g_x = g * f;
      acc[0] += g * (xin - result);
    }
    if (interior) gdeg[off + i] = in01(conv) ? g_res : 0.f;
    else { gdeg[off + i] = 0.f; g_x += g_res; }
    gin[off + i] = g_x;
  }
  block_reduce_store<1>(acc, partial + ((long)b * 3 + c) * gridDim.x + blockIdx.x);
}
// backward pass B: gin += 3x3 correlation of gdeg (symmetric kernel => transpose == same kernel), zero outside
__global__ void __launch_bounds__(kThreads) sharp_bwd_b_kernel(const float* __restrict__ gdeg, float* __restrict__ gin,
                                                              int H, int W) {
  const int b = blockIdx.z, c = blockIdx.y;
  const long off = ((long)b * 3 + c) * H * W;
  const float k1 = 1.0f / 13.0f, k5 = 5.0f / 13.0f;
  for (int y = blockIdx.x * ((H + gridDim.x - 1) / gridDim.x); y < min(H, (int)(blockIdx.x + 1) * (int)((H + gridDim.x - 1) / gridDim.x)); ++y)   // contiguous rows per block: vertical reuse stays in L1
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const int i = y * W + x;
    float s = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        s = fmaf((dy == 0 && dx == 0) ? k5 : k1, gdeg[off + (long)yy * W + xx], s);
      }
    }
    gin[off + i] += s;
  }
}

// ===============================================================================================================
// F7 gaussian blur: kornia.filters.gaussian_blur2d((25,25), sigma, 'reflect', separable) + clamp
// ===============================================================================================================
constexpr int kTaps = 25, kRad = 12;

// per-image 1-D kernel w and its sigma-derivative dw; returns true when the kernel is an exact delta (sigma tiny)
__device__ __forceinline__ bool gauss_weights(float sigma, float* w, float* dw) {
  float e[kTaps], de[kTaps], Z = 0.f, dZ = 0.f;
  const float s2 = sigma * sigma;
#pragma unroll
  for (int k = 0; k < kTaps; ++k) {
    const float x = (float)(k - kRad);
    e[k] = expf(-(x * x) / (2.0f * s2));
    de[k] = e[k] * (x * x) / (s2 * sigma);
    Z += e[k];
    dZ += de[k];
  }
  bool delta = true;
#pragma unroll
  for (int k = 0; k < kTaps; ++k) {
    w[k] = e[k] / Z;
    if (dw) dw[k] = (de[k] - w[k] * dZ) / Z;
    if (k != kRad && w[k] != 0.f) delta = false;
  }
  return delta && w[kRad] == 1.0f;
}
__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

// The per-image 1-D kernels are computed ONCE per launch group by blur_weights_kernel into wbuf[b] = {w[25], dw[25],
// is_delta} (a serial 25-term expf loop per thread block was the dominant cost of the pass kernels).  When the kernel is
// an exact delta -- the reference's start value sigma = 1e-4, where d/d(sigma) underflows to 0 so Adam never moves it --
// every pass degenerates to a float4 copy / clamp / mask.
constexpr int kWStride = 2 * kTaps + 2;
__global__ void blur_weights_kernel(const float* __restrict__ p, int stride, float* __restrict__ wbuf) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float lw[kTaps], ldw[kTaps];
    const bool delta = gauss_weights(p[(long)b * stride], lw, ldw);
    float* o = wbuf + (long)b * kWStride;
    for (int k = 0; k < kTaps; ++k) { o[k] = lw[k]; o[kTaps + k] = ldw[k]; }
    o[2 * kTaps] = delta ? 1.f : 0.f;
  }
}
struct BlurW {
  float w[kTaps], dw[kTaps];
  int is_delta;
};
__device__ __forceinline__ void load_blur_w(const float* __restrict__ wbuf, int b, BlurW* sw) {
  const float* o = wbuf + (long)b * kWStride;
  if (threadIdx.x < kTaps) { sw->w[threadIdx.x] = o[threadIdx.x]; sw->dw[threadIdx.x] = o[kTaps + threadIdx.x]; }
  if (threadIdx.x == 0) sw->is_delta = o[2 * kTaps] != 0.f;
  __syncthreads();
}

// horizontal pass: t = H_w(in) and (optionally) td = H_dw(in)
// Delta images (sigma ~ 0: the whole blur is the identity) are FINISHED here, the later passes skip them:
//   forward  (fin = out,  gout = null): out = clamp(in)
//   backward (fin = gin,  gout = g)   : gin = g * [0 <= in <= 1], d(sigma) = 0
__global__ void __launch_bounds__(kThreads) blur_h_kernel(const float* __restrict__ in, float* __restrict__ t,
                                                         float* __restrict__ td, const float* __restrict__ wbuf,
                                                         float* __restrict__ fin, const float* __restrict__ gout,
                                                         int H, int W) {
  __shared__ BlurW sw;
  const int b = blockIdx.z, c = blockIdx.y;
  load_blur_w(wbuf, b, &sw);
  const long off = ((long)b * 3 + c) * H * W;
  const int HW = H * W;
  if (sw.is_delta) {
    if ((HW & 3) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(in + off);
      const float4* g4 = gout ? reinterpret_cast<const float4*>(gout + off) : nullptr;
      float4* o4 = reinterpret_cast<float4*>(fin + off);
      // a pure streaming pass: four independent 16-byte loads per thread in flight before the first store (the one-load
      // loop ran at 43 % of the copy bandwidth, latency-bound)
      const int n4 = HW / 4, step = gridDim.x * blockDim.x;
      for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * step) {
        float4 v[4], g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * step;
          if (i < n4) { v[u] = s4[i]; if (g4) g[u] = g4[i]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * step;
          if (i >= n4) continue;
          if (g4) o4[i] = make_float4(in01(v[u].x) ? g[u].x : 0.f, in01(v[u].y) ? g[u].y : 0.f, in01(v[u].z) ? g[u].z : 0.f,
                                      in01(v[u].w) ? g[u].w : 0.f);
          else o4[i] = make_float4(clamp01(v[u].x), clamp01(v[u].y), clamp01(v[u].z), clamp01(v[u].w));
        }
      }
    } else {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const float v = in[off + i];
        fin[off + i] = gout ? (in01(v) ? gout[off + i] : 0.f) : clamp01(v);
      }
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float* row = in + off + (long)y * W;
    float s = 0.f, sd = 0.f;
#pragma unroll
    for (int k = 0; k < kTaps; ++k) {
      const float v = row[reflect(x + k - kRad, W)];
      s = fmaf(sw.w[k], v, s);
      sd = fmaf(sw.dw[k], v, sd);
    }
    t[off + i] = s;
    if (td) td[off + i] = sd;
  }
}
// vertical pass (forward): out = clamp(V_w(t))
__global__ void __launch_bounds__(kThreads) blur_v_kernel(const float* __restrict__ t, float* __restrict__ out,
                                                         const float* __restrict__ wbuf, int H, int W) {
  __shared__ BlurW sw;
  const int b = blockIdx.z, c = blockIdx.y;
  load_blur_w(wbuf, b, &sw);
  const long off = ((long)b * 3 + c) * H * W;
  const int HW = H * W;
  if (sw.is_delta) return;                       // finished by blur_h_kernel
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kTaps; ++k) s = fmaf(sw.w[k], t[off + (long)reflect(y + k - kRad, H) * W + x], s);
    out[off + i] = clamp01(s);
  }
}
// backward vertical: pre = V_w(t) -> mask; gm = g*mask; acc += gm * V_dw(t); writes gm (masked gradient)
__global__ void __launch_bounds__(kThreads) blur_bwd_mask_kernel(const float* __restrict__ t, const float* __restrict__ gout,
                                                                float* __restrict__ gm, const float* __restrict__ wbuf,
                                                                float* __restrict__ partial, int H, int W) {
  __shared__ BlurW sw;
  const int b = blockIdx.z, c = blockIdx.y;
  load_blur_w(wbuf, b, &sw);
  const long off = ((long)b * 3 + c) * H * W;
  const int HW = H * W;
  float acc[1] = {0.f};
  if (sw.is_delta) {
    // finished by blur_h_kernel; only the (zero) d(sigma) partial is written
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
      const int y = i / W, x = i - y * W;
      float s = 0.f, sd = 0.f;
#pragma unroll
      for (int k = 0; k < kTaps; ++k) {
        const float v = t[off + (long)reflect(y + k - kRad, H) * W + x];
        s = fmaf(sw.w[k], v, s);
        sd = fmaf(sw.dw[k], v, sd);
      }
      const float g = in01(s) ? gout[off + i] : 0.f;
      gm[off + i] = g;
      acc[0] += g * sd;
    }
  }
  block_reduce_store<1>(acc, partial + ((long)b * 3 + c) * gridDim.x + blockIdx.x);
}
// transpose of a reflect-padded 25-tap correlation along one axis (AXIS 0 = vertical, 1 = horizontal):
//   z[u] = sum_k w[k] * g[u - k + 12] on the padded domain (g = 0 outside), out[i] = z[i] + z[-i] + z[2(n-1)-i]
// optional: acc += out * other  (second d(sigma) term)
template <int AXIS>
__global__ void __launch_bounds__(kThreads) blur_bwd_t_kernel(const float* __restrict__ g, float* __restrict__ out,
                                                             const float* __restrict__ other, const float* __restrict__ wbuf,
                                                             float* __restrict__ partial, int H, int W) {
  __shared__ BlurW sw;
  const int b = blockIdx.z, c = blockIdx.y;
  load_blur_w(wbuf, b, &sw);
  const long off = ((long)b * 3 + c) * H * W;
  const int HW = H * W;
  const int n = AXIS == 0 ? H : W;
  float acc[1] = {0.f};
  if (sw.is_delta) {
    // finished by blur_h_kernel (gin = masked g, d(sigma) = 0)
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
      const int y = i / W, x = i - y * W;
      const int pos = AXIS == 0 ? y : x;
      float s = 0.f;
      int us[3] = {pos, -pos, 2 * (n - 1) - pos};
      bool ok[3] = {true, pos >= 1 && pos <= kRad, pos <= n - 2 && pos >= n - 1 - kRad};
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (!ok[q]) continue;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
          const int src = us[q] - k + kRad;
          if (src < 0 || src >= n) continue;
          const float v = AXIS == 0 ? g[off + (long)src * W + x] : g[off + (long)y * W + src];
          s = fmaf(sw.w[k], v, s);
        }
      }
      out[off + i] = s;
      if (other) acc[0] += s * other[off + i];
    }
  }
  if (other) block_reduce_store<1>(acc, partial + ((long)b * 3 + c) * gridDim.x + blockIdx.x);
}

// ===============================================================================================================
// F8 scale: kornia.geometry.transform.scale -> warp_affine -> affine_grid/grid_sample (bilinear, zeros, align_corners)
//   p = (sx, sy, cx, cy);  M = [[sx,0,(1-sx)cx],[0,sy,(1-sx)cy]]  (kornia reuses alpha=M00 for both translations)
// ===============================================================================================================
struct WarpCoef { float inv_sx, inv_sy, t02, t12; };
__device__ __forceinline__ WarpCoef warp_coef(const float* p, int H, int W) {
  const float sx = p[0], sy = p[1], cx = p[2], cy = p[3];
  const float a = 2.0f / (float)(W - 1), b = 2.0f / (float)(H - 1);
  const float tx = (1.0f - sx) * cx, ty = (1.0f - sx) * cy;
  WarpCoef k;
  k.inv_sx = 1.0f / sx;
  k.inv_sy = 1.0f / sy;
  k.t02 = -(sx + a * tx - 1.0f) / sx;
  k.t12 = -(sy + b * ty - 1.0f) / sy;
  return k;
}
__device__ __forceinline__ float lin_coord(int i, int n) {       // torch.linspace(-1, 1, n)[i]
  const float step = 2.0f / (float)(n - 1);
  return (i < n / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(n - 1 - i));
}

struct SampleCoord { int i0; float w1; };
__device__ __forceinline__ SampleCoord sample_coord(int i, int n, float inv_s, float t) {
  const float gn = lin_coord(i, n) * inv_s + t;
  const float ix = ((gn + 1.0f) * 0.5f) * (float)(n - 1);
  const float f = floorf(ix);
  SampleCoord c;
  c.i0 = (int)f;
  c.w1 = ix - f;
  return c;
}

// forward (BWD=false): out = clamp(bilinear(in)).  backward pass A (BWD=true): gm = g * [0 <= bilinear(in) <= 1] written
// to `out`, plus the per-block partial sums of d(sx, sy, cx, cy).
template <bool BWD>
__global__ void __launch_bounds__(kThreads) scale_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                        float* __restrict__ out, const float* __restrict__ p, int stride,
                                                        float* __restrict__ partial, int H, int W) {
  const int b = blockIdx.y;
  const float* pb = p + (long)b * stride;
  const WarpCoef k = warp_coef(pb, H, W);
  const float sx = pb[0], sy = pb[1], cx = pb[2], cy = pb[3];
  const float a = 2.0f / (float)(W - 1), bb = 2.0f / (float)(H - 1);
  const long base = (long)b * 3 * H * W;
  const int HW = H * W;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int y = blockIdx.x * ((H + gridDim.x - 1) / gridDim.x); y < min(H, (int)(blockIdx.x + 1) * (int)((H + gridDim.x - 1) / gridDim.x)); ++y)   // contiguous rows per block: vertical reuse stays in L1
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const int i = y * W + x;
    const SampleCoord sxc = sample_coord(x, W, k.inv_sx, k.t02), syc = sample_coord(y, H, k.inv_sy, k.t12);
    const int x0 = sxc.i0, y0 = syc.i0, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = sxc.w1, wx0 = 1.0f - wx1, wy1 = syc.w1, wy0 = 1.0f - wy1;
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
    float gix = 0.f, giy = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = in + base + (long)c * HW;
      const float v00 = (vy0 && vx0) ? pl[(long)y0 * W + x0] : 0.f;
      const float v01 = (vy0 && vx1) ? pl[(long)y0 * W + x1] : 0.f;
      const float v10 = (vy1 && vx0) ? pl[(long)y1 * W + x0] : 0.f;
      const float v11 = (vy1 && vx1) ? pl[(long)y1 * W + x1] : 0.f;
      const float o = v00 * (wx0 * wy0) + v01 * (wx1 * wy0) + v10 * (wx0 * wy1) + v11 * (wx1 * wy1);
      if (!BWD) {
        out[base + (long)c * HW + i] = clamp01(o);
      } else {
        const float g = in01(o) ? gout[base + (long)c * HW + i] : 0.f;
        out[base + (long)c * HW + i] = g;
        gix += g * ((v01 - v00) * wy0 + (v11 - v10) * wy1);
        giy += g * ((v10 - v00) * wx0 + (v11 - v01) * wx1);
      }
    }
    if (BWD) {
      const float xn = lin_coord(x, W), yn = lin_coord(y, H);
      const float ggx = gix * 0.5f * (float)(W - 1), ggy = giy * 0.5f * (float)(H - 1);   // d/d(grid x), d/d(grid y)
      acc[0] += ggx * ((-xn + a * cx - 1.0f) / (sx * sx)) + ggy * (bb * cy / sy);         // d/dsx
      acc[1] += ggy * ((-yn + bb * (1.0f - sx) * cy - 1.0f) / (sy * sy));                  // d/dsy
      acc[2] += ggx * (-a * (1.0f - sx) / sx);                                             // d/dcx
      acc[3] += ggy * (-bb * (1.0f - sx) / sy);                                            // d/dcy
    }
  }
  if (BWD) block_reduce_store<4>(acc, partial + ((long)b * gridDim.x + blockIdx.x) * 4);
}

// backward pass B: d(in)[v,u] = sum over the destination pixels whose bilinear footprint contains (v,u) -- a gather, so
// the result is deterministic (no atomics).  The warp is separable and monotone: the candidate destination columns /
// rows of a source column / row form a contiguous range found by inverting the affine coordinate map.
__global__ void __launch_bounds__(kThreads) scale_bwd_gather_kernel(const float* __restrict__ gm, float* __restrict__ gin,
                                                                   const float* __restrict__ p, int stride, int H, int W) {
  const int b = blockIdx.y;
  const WarpCoef k = warp_coef(p + (long)b * stride, H, W);
  const long base = (long)b * 3 * H * W;
  const int HW = H * W;
  // ix(x) ~= c0x + slope_x * x  (exact values are recomputed per candidate with sample_coord)
  const float gx0 = lin_coord(0, W) * k.inv_sx + k.t02, gx1 = lin_coord(W - 1, W) * k.inv_sx + k.t02;
  const float c0x = ((gx0 + 1.0f) * 0.5f) * (float)(W - 1);
  const float slx = (((gx1 + 1.0f) * 0.5f) * (float)(W - 1) - c0x) / (float)(W - 1);
  const float gy0 = lin_coord(0, H) * k.inv_sy + k.t12, gy1 = lin_coord(H - 1, H) * k.inv_sy + k.t12;
  const float c0y = ((gy0 + 1.0f) * 0.5f) * (float)(H - 1);
  const float sly = (((gy1 + 1.0f) * 0.5f) * (float)(H - 1) - c0y) / (float)(H - 1);
  for (int v = blockIdx.x * ((H + gridDim.x - 1) / gridDim.x); v < min(H, (int)(blockIdx.x + 1) * (int)((H + gridDim.x - 1) / gridDim.x)); ++v)
  for (int u = threadIdx.x; u < W; u += blockDim.x) {
    const int i = v * W + u;
    int xlo = (int)floorf(((float)(u - 1) - c0x) / slx) - 1, xhi = (int)ceilf(((float)(u + 1) - c0x) / slx) + 1;
    int ylo = (int)floorf(((float)(v - 1) - c0y) / sly) - 1, yhi = (int)ceilf(((float)(v + 1) - c0y) / sly) + 1;
    xlo = max(xlo, 0); xhi = min(xhi, W - 1); ylo = max(ylo, 0); yhi = min(yhi, H - 1);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      const SampleCoord cy = sample_coord(y, H, k.inv_sy, k.t12);
      float wy;
      if (cy.i0 == v) wy = 1.0f - cy.w1;
      else if (cy.i0 + 1 == v) wy = cy.w1;
      else continue;
      for (int x = xlo; x <= xhi; ++x) {
        const SampleCoord cx = sample_coord(x, W, k.inv_sx, k.t02);
        float wx;
        if (cx.i0 == u) wx = 1.0f - cx.w1;
        else if (cx.i0 + 1 == u) wx = cx.w1;
        else continue;
        const float w = wx * wy;
        const long o = base + (long)y * W + x;
        s0 = fmaf(w, gm[o], s0);
        s1 = fmaf(w, gm[o + HW], s1);
        s2 = fmaf(w, gm[o + 2L * HW], s2);
      }
    }
    gin[base + i] = s0; gin[base + HW + i] = s1; gin[base + 2L * HW + i] = s2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Table-driven scale kernels.  The warp is separable: pixel (y, x) samples source row (i0y(y), w1y(y)) and source column
// (i0x(x), w1x(x)), both monotone in their index (sx, sy > 0).  Each block builds the column table of its image in
// shared memory once (W coordinate evaluations, amortised over its band of rows) instead of evaluating two coordinates
// per pixel, and the gather kernel inverts the tables by binary search -- source column u receives from the contiguous
// destination range {x : i0x(x) in {u - 1, u}} -- instead of re-deriving candidates with divisions per pixel.
// Same arithmetic per coordinate (sample_coord) as scale_kernel / scale_bwd_gather_kernel, which remain the reference
// form (RGIE_SCALE_TAB=0 or RGIE_SCALE_COL=0) and the path for tables that would not fit shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScaleRows = 16;          // destination (or source) rows per block
__device__ __forceinline__ void build_coord_table(int* __restrict__ ti0, float* __restrict__ tw1, int n, float inv_s, float t) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const SampleCoord c = sample_coord(i, n, inv_s, t);
    ti0[i] = c.i0; tw1[i] = c.w1;
  }
}

// Column-marching form of the same pass: a thread owns ONE destination column (its source columns and weights are
// loop-invariant registers) and walks down a band of rows; the two source rows of the previous destination row stay in
// registers, and because consecutive destination rows sample source rows at most one apart when sy >= 1 (the reference
// clamps the scale factors to >= 1, optimize_image_param.py:279-280) a row step needs 0 or 1 new source rows
// (6 loads per pixel instead of 12).  Any other step reloads both rows, so every sx, sy > 0 stays correct.
// Values and operation order per pixel are those of scale_kernel: results are bit-identical.
constexpr int kScaleColRows = 32;
template <bool BWD>
__global__ void __launch_bounds__(kThreads) scale_col_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                            float* __restrict__ out, const float* __restrict__ p, int stride,
                                                            float* __restrict__ partial, int H, int W) {
  const int b = blockIdx.z;
  const float* pb = p + (long)b * stride;
  const WarpCoef k = warp_coef(pb, H, W);
  const int x = blockIdx.x * kThreads + threadIdx.x;
  const bool active = x < W;
  const int ya = blockIdx.y * kScaleColRows, yb = min(ya + kScaleColRows, H);
  const long base = (long)b * 3 * H * W;
  const int HW = H * W;
  const SampleCoord sxc = sample_coord(active ? x : 0, W, k.inv_sx, k.t02);
  const int x0 = sxc.i0, x1 = x0 + 1;
  const float wx1 = sxc.w1, wx0 = 1.0f - wx1;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
  const float xn = lin_coord(active ? x : 0, W);
  float top[3][2], bot[3][2];                       // source rows cur, cur + 1 at columns x0, x1 (zero outside the image)
  int cur = -0x40000000;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  auto load_row = [&](int r, float (*dst)[2]) {
    const bool vr = r >= 0 && r < H;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* row = in + base + (long)c * HW + (long)r * W;
      dst[c][0] = (vr && vx0) ? row[x0] : 0.f;
      dst[c][1] = (vr && vx1) ? row[x1] : 0.f;
    }
  };
  if (active)
  for (int y = ya; y < yb; ++y) {
    const SampleCoord syc = sample_coord(y, H, k.inv_sy, k.t12);
    const int y0 = syc.i0;
    const float wy1 = syc.w1, wy0 = 1.0f - wy1;
    if (y0 == cur + 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { top[c][0] = bot[c][0]; top[c][1] = bot[c][1]; }
      load_row(y0 + 1, bot);
    } else if (y0 != cur) {
      load_row(y0, top);
      load_row(y0 + 1, bot);
    }
    cur = y0;
    const int i = y * W + x;
    float gix = 0.f, giy = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v00 = top[c][0], v01 = top[c][1], v10 = bot[c][0], v11 = bot[c][1];
      const float o = v00 * (wx0 * wy0) + v01 * (wx1 * wy0) + v10 * (wx0 * wy1) + v11 * (wx1 * wy1);
      if (!BWD) {
        out[base + (long)c * HW + i] = clamp01(o);
      } else {
        const float g = in01(o) ? gout[base + (long)c * HW + i] : 0.f;
        out[base + (long)c * HW + i] = g;
        gix += g * ((v01 - v00) * wy0 + (v11 - v10) * wy1);
        giy += g * ((v10 - v00) * wx0 + (v11 - v01) * wx1);
      }
    }
    if (BWD) {
      const float ggx = gix * 0.5f * (float)(W - 1), ggy = giy * 0.5f * (float)(H - 1);   // d/d(grid x), d/d(grid y)
      a0 += ggx; a1 = fmaf(ggx, xn, a1);
      b0 += ggy; b1 = fmaf(ggy, lin_coord(y, H), b1);
    }
  }
  if (BWD) {
    float acc[4] = {a0, a1, b0, b1};
    __shared__ float sums[4];
    block_reduce_store<4>(acc, sums);
    __syncthreads();
    if (threadIdx.x == 0) {
      const float sx = pb[0], sy = pb[1], cx = pb[2], cy = pb[3];
      const float a = 2.0f / (float)(W - 1), bb = 2.0f / (float)(H - 1);
      float* dst = partial + ((long)b * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x) * 4;
      dst[0] = (-sums[1] + (a * cx - 1.0f) * sums[0]) / (sx * sx) + sums[2] * (bb * cy / sy);      // d/dsx
      dst[1] = (-sums[3] + (bb * (1.0f - sx) * cy - 1.0f) * sums[2]) / (sy * sy);                  // d/dsy
      dst[2] = sums[0] * (-a * (1.0f - sx) / sx);                                                 // d/dcx
      dst[3] = sums[2] * (-bb * (1.0f - sx) / sy);                                                // d/dcy
    }
  }
}
bool scale_col_on() {
  static const int env = getenv("RGIE_SCALE_COL") ? atoi(getenv("RGIE_SCALE_COL")) : 1;
  return env != 0;
}

// first index i in [0, n) with tab[i] >= key (n when none); tab is non-decreasing
__device__ __forceinline__ int lower_bound_i(const int* __restrict__ tab, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (tab[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// backward pass B: d(in)[v, u] = sum_{y in Y(v)} sum_{x in X(u)} wy * wx * gm[y, x]; a deterministic gather
__global__ void __launch_bounds__(kThreads) scale_gather_tab_kernel(const float* __restrict__ gm, float* __restrict__ gin,
                                                                   const float* __restrict__ p, int stride, int H, int W) {
  extern __shared__ __align__(16) int s_tab[];              // [W] i0x, [W] w1x, [H] i0y, [H] w1y, [W+2] lbx, [rows+2] lby
  int* xi0 = s_tab;
  float* xw1 = reinterpret_cast<float*>(s_tab + W);
  int* yi0 = s_tab + 2 * W;
  float* yw1 = reinterpret_cast<float*>(s_tab + 2 * W + H);
  int* lbx = s_tab + 2 * W + 2 * H;                         // lbx[key + 1] = first x with i0x(x) >= key, key in [-1, W]
  int* lby = lbx + W + 2;                                   // lby[key - va + 1] likewise for keys [va - 1, vb]
  const int b = blockIdx.y;
  const WarpCoef k = warp_coef(p + (long)b * stride, H, W);
  build_coord_table(xi0, xw1, W, k.inv_sx, k.t02);
  build_coord_table(yi0, yw1, H, k.inv_sy, k.t12);
  __syncthreads();
  const int va = blockIdx.x * kScaleRows, vb = min(va + kScaleRows, H);
  for (int i = threadIdx.x; i < W + 2; i += blockDim.x) lbx[i] = lower_bound_i(xi0, W, i - 1);
  for (int i = threadIdx.x; i < vb - va + 2; i += blockDim.x) lby[i] = lower_bound_i(yi0, H, va + i - 1);
  __syncthreads();
  const long base = (long)b * 3 * H * W;
  const int HW = H * W;
  for (int v = va; v < vb; ++v) {
    const int ylo = lby[v - va], yhi = lby[v - va + 2];     // [ylo, yhi): i0y in {v-1, v}
    for (int u = threadIdx.x; u < W; u += blockDim.x) {
      const int xlo = lbx[u], xhi = lbx[u + 2];             // [xlo, xhi): i0x in {u-1, u}
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      for (int y = ylo; y < yhi; ++y) {
        const float wy = yi0[y] == v ? 1.0f - yw1[y] : yw1[y];
        for (int x = xlo; x < xhi; ++x) {
          const float wx = xi0[x] == u ? 1.0f - xw1[x] : xw1[x];
          const float w = wx * wy;
          const long o = base + (long)y * W + x;
          s0 = fmaf(w, gm[o], s0);
          s1 = fmaf(w, gm[o + HW], s1);
          s2 = fmaf(w, gm[o + 2L * HW], s2);
        }
      }
      const int i = v * W + u;
      gin[base + i] = s0; gin[base + HW + i] = s1; gin[base + 2L * HW + i] = s2;
    }
  }
}
constexpr size_t scale_gather_smem(int H, int W) { return (size_t)(3 * W + 2 * H + kScaleRows + 4) * sizeof(int); }
bool scale_tab_ok(int H, int W) {
  static const int env = getenv("RGIE_SCALE_TAB") ? atoi(getenv("RGIE_SCALE_TAB")) : 1;
  return env != 0 && scale_gather_smem(H, W) <= 44 * 1024;
}

// ===============================================================================================================
// affine: kornia.geometry.transform.affine(im, M[2x3], padding_mode='border') + clamp   (image_transformations.py:198-206)
//   warp_affine: theta = inv(N M3 N^-1)[:2] with N the pixel -> [-1,1] normalisation, affine_grid(align_corners=True),
//   grid_sample(bilinear, border, align_corners=True).  Border padding clips the sampling coordinate to [0, size-1]; the
//   gradient through a clipped coordinate is zero (ATen clip_coordinates_set_grad).
// Forward: one gather per pixel.  Backward: d(theta) by block reductions (6 sums), d(image) by atomic scatter-add of the
// four bilinear corners (summation order is not fixed), then d(M) from d(theta) through the 3x3 inverse in a one-thread
// kernel per image.
// ===============================================================================================================
struct AffCoef { float t00, t01, t02, t10, t11, t12; };
__device__ __forceinline__ AffCoef affine_theta(const float* m, int H, int W) {
  const float a = 2.0f / (float)(W - 1), b = 2.0f / (float)(H - 1), ia = 0.5f * (float)(W - 1), ib = 0.5f * (float)(H - 1);
  const float A00 = m[0], A01 = m[1] * (a * ib), A02 = a * (m[0] * ia + m[1] * ib + m[2]) - 1.0f;
  const float A10 = m[3] * (b * ia), A11 = m[4], A12 = b * (m[3] * ia + m[4] * ib + m[5]) - 1.0f;
  const float det = A00 * A11 - A01 * A10;
  AffCoef t;
  t.t00 = A11 / det; t.t01 = -A01 / det; t.t10 = -A10 / det; t.t11 = A00 / det;
  t.t02 = -(t.t00 * A02 + t.t01 * A12);
  t.t12 = -(t.t10 * A02 + t.t11 * A12);
  return t;
}

template <bool BWD>
__global__ void __launch_bounds__(kThreads) affine_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                         float* __restrict__ out, const float* __restrict__ p, int stride,
                                                         float* __restrict__ partial, int H, int W) {
  const int b = blockIdx.y;
  const AffCoef k = affine_theta(p + (long)b * stride, H, W);
  const long base = (long)b * 3 * H * W;
  const int HW = H * W;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int y = blockIdx.x * ((H + gridDim.x - 1) / gridDim.x); y < min(H, (int)(blockIdx.x + 1) * (int)((H + gridDim.x - 1) / gridDim.x)); ++y)
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const int i = y * W + x;
    const float xn = lin_coord(x, W), yn = lin_coord(y, H);
    const float xs = xn * k.t00 + yn * k.t01 + k.t02, ys = xn * k.t10 + yn * k.t11 + k.t12;
    float ix = ((xs + 1.0f) * 0.5f) * (float)(W - 1), iy = ((ys + 1.0f) * 0.5f) * (float)(H - 1);
    // ATen clip_coordinates_set_grad: the coordinate gradient is zero for in <= 0 and in >= size - 1
    const bool cx = ix <= 0.f || ix >= (float)(W - 1), cy = iy <= 0.f || iy >= (float)(H - 1);
    ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx, wx0 = 1.0f - wx1, wy1 = iy - fy, wy0 = 1.0f - wy1;
    const bool vx1 = x1 < W, vy1 = y1 < H;                     // x0, y0 are always inside after the clip
    float gix = 0.f, giy = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = in + base + (long)c * HW;
      const float v00 = pl[(long)y0 * W + x0];
      const float v01 = vx1 ? pl[(long)y0 * W + x1] : 0.f;
      const float v10 = vy1 ? pl[(long)y1 * W + x0] : 0.f;
      const float v11 = (vy1 && vx1) ? pl[(long)y1 * W + x1] : 0.f;
      const float o = v00 * (wx0 * wy0) + v01 * (wx1 * wy0) + v10 * (wx0 * wy1) + v11 * (wx1 * wy1);
      if (!BWD) {
        out[base + (long)c * HW + i] = clamp01(o);
      } else {
        const float g = in01(o) ? gout[base + (long)c * HW + i] : 0.f;
        float* go = out + base + (long)c * HW;                  // d(image), zeroed by the launcher
        atomicAdd(go + (long)y0 * W + x0, g * (wx0 * wy0));
        if (vx1) atomicAdd(go + (long)y0 * W + x1, g * (wx1 * wy0));
        if (vy1) atomicAdd(go + (long)y1 * W + x0, g * (wx0 * wy1));
        if (vy1 && vx1) atomicAdd(go + (long)y1 * W + x1, g * (wx1 * wy1));
        gix += g * ((v01 - v00) * wy0 + (v11 - v10) * wy1);
        giy += g * ((v10 - v00) * wx0 + (v11 - v01) * wx1);
      }
    }
    if (BWD) {
      const float gxs = cx ? 0.f : gix * 0.5f * (float)(W - 1), gys = cy ? 0.f : giy * 0.5f * (float)(H - 1);
      acc[0] += gxs * xn; acc[1] += gxs * yn; acc[2] += gxs;
      acc[3] += gys * xn; acc[4] += gys * yn; acc[5] += gys;
    }
  }
  if (BWD) block_reduce_store<6>(acc, partial + ((long)b * gridDim.x + blockIdx.x) * 6);
}

// d(M) from d(theta): theta = inv(A)[:2], A = N M3 N^-1  =>  dA = -T^T dT T^T, then the chain through A(M)
__global__ void affine_param_grad_kernel(const float* __restrict__ dtheta, const float* __restrict__ p, int stride,
                                         float* __restrict__ gp, int gp_stride, int B, int H, int W) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* m = p + (long)b * stride;
  const AffCoef k = affine_theta(m, H, W);
  const float T[3][3] = {{k.t00, k.t01, k.t02}, {k.t10, k.t11, k.t12}, {0.f, 0.f, 1.f}};
  const float* g = dtheta + 6L * b;
  const float G[3][3] = {{g[0], g[1], g[2]}, {g[3], g[4], g[5]}, {0.f, 0.f, 0.f}};
  float TG[3][3], dA[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int q = 0; q < 3; ++q) s += T[q][i] * G[q][j];          // T^T G
      TG[i][j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int q = 0; q < 3; ++q) s += TG[i][q] * T[j][q];          // (T^T G) T^T
      dA[i][j] = -s;
    }
  const float a = 2.0f / (float)(W - 1), bb = 2.0f / (float)(H - 1), ia = 0.5f * (float)(W - 1), ib = 0.5f * (float)(H - 1);
  float* o = gp + (long)b * gp_stride;
  o[0] = dA[0][0] + dA[0][2] * (a * ia);
  o[1] = (dA[0][1] + dA[0][2]) * (a * ib);
  o[2] = dA[0][2] * a;
  o[3] = (dA[1][0] + dA[1][2]) * (bb * ia);
  o[4] = dA[1][1] + dA[1][2] * (bb * ib);
  o[5] = dA[1][2] * bb;
}

// ---------------------------------------------------------------------------------------------------------------
// Marching kernels (W % 4 == 0): a thread owns FOUR adjacent columns and walks down a band of rows with the three input
// rows it needs (6 values each: its 4 columns + one neighbour either side) rolling through registers, so every input
// element is fetched once per band as part of a 16-byte load.  Backward: the gradient of the clamped 3x3 smoothing
// (gdeg) goes through a 4-row ring in shared memory (one __syncthreads per row), the transposed 3x3 reads it from
// there, and the direct term of the previous row waits in registers -- in + g are read once, gin is written once (3N),
// no intermediate in HBM.  The 9-term fmaf chains are those of sharp_conv / sharp_bwd_b_kernel, in the same order (the
// [0 <= conv <= 1] mask of a saturated neighbourhood depends on it).  (An earlier form that staged whole row bands of in,
// g, gdeg and the direct term in shared memory -- 80 KB per block, scalar 3x3 reads -- measured 0.58 ms backward against 0.53
// for the two global-memory passes and 0.19 for this one; it is no longer in the file.)
// ---------------------------------------------------------------------------------------------------------------
struct Row6 { float v[6]; };     // columns x0-1 .. x0+4 of one row (zero outside the image: only read for interior pixels)
__device__ __forceinline__ Row6 load_row6(const float* __restrict__ row, int x0, int W, bool valid) {
  Row6 r;
  if (!valid) {
#pragma unroll
    for (int i = 0; i < 6; ++i) r.v[i] = 0.f;
    return r;
  }
  const float4 q = *reinterpret_cast<const float4*>(row + x0);
  r.v[1] = q.x; r.v[2] = q.y; r.v[3] = q.z; r.v[4] = q.w;
  r.v[0] = x0 > 0 ? row[x0 - 1] : 0.f;
  r.v[5] = x0 + 4 < W ? row[x0 + 4] : 0.f;
  return r;
}
__device__ __forceinline__ float conv_rows(const Row6& a, const Row6& b, const Row6& c, int j) {   // pixel column x0 + j
  const float k1 = 1.0f / 13.0f, k5 = 5.0f / 13.0f;
  float s = 0.f;
  s = fmaf(k1, a.v[j], s); s = fmaf(k1, a.v[j + 1], s); s = fmaf(k1, a.v[j + 2], s);
  s = fmaf(k1, b.v[j], s); s = fmaf(k5, b.v[j + 1], s); s = fmaf(k1, b.v[j + 2], s);
  s = fmaf(k1, c.v[j], s); s = fmaf(k1, c.v[j + 1], s); s = fmaf(k1, c.v[j + 2], s);
  return s;
}

__global__ void __launch_bounds__(1024) sharp_fwd_march_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              const float* __restrict__ p, int stride, int H, int W, int R) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float f = p[(long)b * stride];
  const int mode = sharp_mode(f);
  const long off = ((long)b * 3 + c) * H * W;
  const int x0 = 4 * threadIdx.x;
  if (x0 >= W) return;
  const int y0 = blockIdx.x * R, y1 = min(y0 + R, H);
  const float* pl = in + off;
  Row6 ra = load_row6(pl + (long)(y0 - 1) * W, x0, W, y0 >= 1);
  Row6 rb = load_row6(pl + (long)y0 * W, x0, W, true);
  for (int y = y0; y < y1; ++y) {
    const Row6 rc = load_row6(pl + (long)(y + 1) * W, x0, W, y + 1 < H);
    float o4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + j;
      const float xin = rb.v[j + 1];
      float result = xin;
      if (y >= 1 && y < H - 1 && x >= 1 && x < W - 1) result = clamp01(conv_rows(ra, rb, rc, j));
      float o;
      if (mode == 0) o = result;
      else if (mode == 1) o = xin;
      else {
        o = __fadd_rn(result, __fmul_rn(xin - result, f));
        if (mode == 3) o = clamp01(o);
      }
      o4[j] = clamp01(o);
    }
    *reinterpret_cast<float4*>(out + off + (long)y * W + x0) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    ra = rb; rb = rc;
  }
}

__global__ void __launch_bounds__(1024) sharp_bwd_march_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                               float* __restrict__ gin, const float* __restrict__ p,
                                                               int stride, float* __restrict__ partial, int H, int W, int R) {
  extern __shared__ __align__(16) float ring[];            // 4 rows of (W + 8) floats: gdeg with 4 zero columns either side
  __shared__ float red[32];
  const int b = blockIdx.z, c = blockIdx.y;
  const float f = p[(long)b * stride];
  const int mode = sharp_mode(f);
  const long off = ((long)b * 3 + c) * H * W;
  const int Wp = W + 8;
  const int x0 = 4 * threadIdx.x;
  const bool active = x0 < W;
  const int y0 = blockIdx.x * R, y1 = min(y0 + R, H);
  const int ga = max(y0 - 1, 0), gb = min(y1 + 1, H);      // rows whose smoothing gradient this band needs
  for (int i = threadIdx.x; i < 4 * Wp; i += blockDim.x) ring[i] = 0.f;
  __syncthreads();
  const float* pl = in + off;
  const float k1 = 1.0f / 13.0f, k5 = 5.0f / 13.0f;
  Row6 ra, rb;
  if (active) {
    ra = load_row6(pl + (long)(ga - 1) * W, x0, W, ga >= 1);
    rb = load_row6(pl + (long)ga * W, x0, W, true);
  }
  float acc = 0.f;
  float hold[4] = {0.f, 0.f, 0.f, 0.f};                    // direct d(in) term of the previous row
  for (int y = ga; y <= gb; ++y) {                         // iteration y == gb only emits row gb - 1 (below it: zeros)
    float gx[4] = {0.f, 0.f, 0.f, 0.f};
    if (active && y < gb) {
      const Row6 rc = load_row6(pl + (long)(y + 1) * W, x0, W, y + 1 < H);
      const float4 g4 = *reinterpret_cast<const float4*>(gout + off + (long)y * W + x0);
      const float gq[4] = {g4.x, g4.y, g4.z, g4.w};
      const bool own = y >= y0 && y < y1;
      float deg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + j;
        const float xin = rb.v[j + 1];
        const bool interior = y >= 1 && y < H - 1 && x >= 1 && x < W - 1;
        float conv = 0.f, result = xin;
        if (interior) { conv = conv_rows(ra, rb, rc, j); result = clamp01(conv); }
        float g = gq[j], g_res = 0.f, g_x = 0.f;
        if (mode == 0) { g = in01(result) ? g : 0.f; g_res = g; }
        else if (mode == 1) { g = in01(xin) ? g : 0.f; g_x = g; }
        else {
          const float o = __fadd_rn(result, __fmul_rn(xin - result, f));
          g = in01(o) ? g : 0.f;
          g_res = g - g * f;
          g_x = g * f;
          if (own) acc += g * (xin - result);
        }
        if (interior) deg[j] = in01(conv) ? g_res : 0.f;
        else { deg[j] = 0.f; g_x += g_res; }
        gx[j] = g_x;
      }
      *reinterpret_cast<float4*>(ring + (y & 3) * Wp + 4 + x0) = make_float4(deg[0], deg[1], deg[2], deg[3]);
      ra = rb; rb = rc;
    }
    __syncthreads();
    const int ye = y - 1;                                  // row to emit: needs gdeg rows ye-1, ye, ye+1 (= y)
    if (active && ye >= y0 && ye < y1) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = ye + dy;
        if (yy < 0 || yy >= H) continue;
        const float* rr = ring + (yy & 3) * Wp + 4 + x0;
        const float4 q = *reinterpret_cast<const float4*>(rr);
        const float v[6] = {rr[-1], q.x, q.y, q.z, q.w, rr[4]};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[j] = fmaf(k1, v[j], s[j]);
          s[j] = fmaf(dy == 0 ? k5 : k1, v[j + 1], s[j]);
          s[j] = fmaf(k1, v[j + 2], s[j]);
        }
      }
      *reinterpret_cast<float4*>(gin + off + (long)ye * W + x0) =
          make_float4(hold[0] + s[0], hold[1] + s[1], hold[2] + s[2], hold[3] + s[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) hold[j] = gx[j];
  }
  // block sum of d(factor) over the band's own rows (any block size that is a multiple of 32)
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    partial[((long)b * 3 + c) * gridDim.x + blockIdx.x] = t;
  }
}
struct SharpMarch { int R, nb, threads; size_t smem; };
SharpMarch sharp_march(int H, int W) {
  SharpMarch m;
  m.R = H >= 256 ? 32 : 16;
  m.nb = ceil_div(H, m.R);
  m.threads = ceil_div(ceil_div(W, 4), 32) * 32;
  m.smem = (size_t)4 * (W + 8) * sizeof(float);
  static const int env = getenv("RGIE_SHARP_MARCH") ? atoi(getenv("RGIE_SHARP_MARCH")) : 1;
  if (!env || (W & 3) != 0 || m.threads > 1024 || m.nb > 64) m.R = 0;       // fall back (partials: 3 * nb <= 192 per image)
  return m;
}

int plane_blocks(int HW) { int n = ceil_div(HW, kThreads * 4); return n > 64 ? 64 : (n < 1 ? 1 : n); }

}  // namespace
}  // namespace rgie

using namespace rgie;

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" {

long rgie_filter_ws_floats(int B, int H, int W) {
  // worst case over all filters: partials (B * kMaxBlk * 24) + three full-size temporaries (blur backward)
  return (long)B * kMaxBlk * 24 + 3L * B * 3 * H * W + 64;
}

int rgie_filter_param_count(int kind) {
  switch (kind) {
    case RGIE_F_EXPOSURE: case RGIE_F_SATURATION: case RGIE_F_CONTRAST: case RGIE_F_SHARP: case RGIE_F_BLUR: return 1;
    case RGIE_F_GAMMA: case RGIE_F_BRIGHT: case RGIE_F_BW: case RGIE_F_HUE: case RGIE_F_WB: return 1;
    case RGIE_F_AFFINE: return 6;
    case RGIE_F_TONE: return 8;
    case RGIE_F_COLOR: return 24;
    case RGIE_F_SCALE: return 4;
  }
  return -1;
}

int rgie_filter_fwd(int kind, const float* in, float* out, const float* p, int p_stride, int B, int H, int W,
                    float* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  RGIE_CHECK(B > 0 && H > 0 && W > 0, "rgie_filter_fwd: bad shape");
  switch (kind) {
    case RGIE_F_EXPOSURE: { ExposureOp op{p, p_stride, 0.f}; return launch_pointwise<ExposureOp, false>(in, nullptr, out, op, nullptr, B, HW, st); }
    case RGIE_F_SATURATION: { SaturationOp op{p, p_stride, 0.f}; return launch_pointwise<SaturationOp, false>(in, nullptr, out, op, nullptr, B, HW, st); }
    case RGIE_F_TONE: { CurveOp<1> op; op.p = p; op.stride = p_stride; return launch_pointwise<CurveOp<1>, false>(in, nullptr, out, op, nullptr, B, HW, st); }
    case RGIE_F_COLOR: { CurveOp<3> op; op.p = p; op.stride = p_stride; return launch_pointwise<CurveOp<3>, false>(in, nullptr, out, op, nullptr, B, HW, st); }
    case RGIE_F_CONTRAST: {
      RGIE_CHECK(ws != nullptr, "rgie_filter_fwd(contrast): workspace required");
      LaunchShape s = shape_for(HW);
      float* partial = ws;                 // [B, nblk]
      float* mu = ws + (long)B * kMaxBlk;  // [B]
      dim3 grid(s.nblk, B);
      if (HW % 4 == 0) gray_sum_kernel<4><<<grid, kThreads, 0, st>>>(in, partial, HW, s.chunk);
      else gray_sum_kernel<1><<<grid, kThreads, 0, st>>>(in, partial, HW, s.chunk);
      RGIE_LAUNCH_OK();
      gray_mean_finalize<<<B, 32, 0, st>>>(partial, s.nblk, 1.0f / (float)HW, mu);
      RGIE_LAUNCH_OK();
      ContrastOp op{p, p_stride, mu, 0.f, 0.f};
      return launch_pointwise<ContrastOp, false>(in, nullptr, out, op, nullptr, B, HW, st);
    }
    case RGIE_F_GAMMA: { GammaOp op{p, p_stride, 0.f}; return launch_pointwise<GammaOp, false>(in, nullptr, out, op, nullptr, B, HW, st); }
    case RGIE_F_BRIGHT: { BrightOp op{p, p_stride, 0.f}; return launch_pointwise<BrightOp, false>(in, nullptr, out, op, nullptr, B, HW, st); }
    case RGIE_F_BW: { BwOp op{p, p_stride, 0.f}; return launch_pointwise<BwOp, false>(in, nullptr, out, op, nullptr, B, HW, st); }
    case RGIE_F_HUE: { HueOp op{p, p_stride, 0.f}; return launch_pointwise<HueOp, false>(in, nullptr, out, op, nullptr, B, HW, st); }
    case RGIE_F_WB: {
      RGIE_CHECK(ws != nullptr, "rgie_filter_fwd(wb): workspace required");
      LaunchShape s = shape_for(HW);
      float* partial = ws;                      // [B, nblk, 3]
      float* mean = ws + (long)B * kMaxBlk * 4; // [B, 3]
      dim3 grid(s.nblk, B);
      chan_sum_kernel<<<grid, kThreads, 0, st>>>(in, partial, HW, s.chunk);
      RGIE_LAUNCH_OK();
      chan_mean_finalize<<<B, 32, 0, st>>>(partial, s.nblk, 1.0f / (float)HW, mean);
      RGIE_LAUNCH_OK();
      WbOp op; op.p = p; op.stride = p_stride; op.mean_all = mean;
      return launch_pointwise<WbOp, false>(in, nullptr, out, op, nullptr, B, HW, st);
    }
    case RGIE_F_SHARP: {
      RGIE_CHECK(H >= 3 && W >= 3, "sharp: image too small");
      const SharpMarch sm = sharp_march(H, W);
      if (sm.R > 0) {
        sharp_fwd_march_kernel<<<dim3(sm.nb, 3, B), sm.threads, 0, st>>>(in, out, p, p_stride, H, W, sm.R);
        RGIE_LAUNCH_OK();
        return 0;
      }
      dim3 grid(plane_blocks(HW), 3, B);
      sharp_fwd_kernel<<<grid, kThreads, 0, st>>>(in, out, p, p_stride, H, W);
      RGIE_LAUNCH_OK();
      return 0;
    }
    case RGIE_F_BLUR: {
      RGIE_CHECK(ws != nullptr, "rgie_filter_fwd(blur): workspace required");
      RGIE_CHECK(H > kRad && W > kRad, "blur: reflect padding needs H,W > 12");
      float* wbuf = ws;                         // [B, kWStride]
      float* t = ws + (long)B * kMaxBlk * 24;
      dim3 grid(plane_blocks(HW), 3, B);
      blur_weights_kernel<<<B, 32, 0, st>>>(p, p_stride, wbuf);
      RGIE_LAUNCH_OK();
      blur_h_kernel<<<grid, kThreads, 0, st>>>(in, t, nullptr, wbuf, out, nullptr, H, W);
      RGIE_LAUNCH_OK();
      blur_v_kernel<<<grid, kThreads, 0, st>>>(t, out, wbuf, H, W);
      RGIE_LAUNCH_OK();
      return 0;
    }
    case RGIE_F_SCALE: {
      RGIE_CHECK(H >= 2 && W >= 2, "scale: image too small");
      // column-marching kernel: 0.104 ms against 0.135 ms for the direct-coordinate kernel below (64 x 512^2; a per-block
      // coordinate table in shared memory measured 0.152 ms and is no longer in the file).  RGIE_SCALE_COL=0: kernel below
      if (scale_col_on()) {
        scale_col_kernel<false><<<dim3(ceil_div(W, kThreads), ceil_div(H, kScaleColRows), B), kThreads, 0, st>>>(in, nullptr, out, p, p_stride, nullptr, H, W);
        RGIE_LAUNCH_OK();
        return 0;
      }
      dim3 grid(plane_blocks(HW), B);
      scale_kernel<false><<<grid, kThreads, 0, st>>>(in, nullptr, out, p, p_stride, nullptr, H, W);
      RGIE_LAUNCH_OK();
      return 0;
    }
    case RGIE_F_AFFINE: {
      RGIE_CHECK(H >= 2 && W >= 2, "affine: image too small");
      dim3 grid(plane_blocks(HW), B);
      affine_kernel<false><<<grid, kThreads, 0, st>>>(in, nullptr, out, p, p_stride, nullptr, H, W);
      RGIE_LAUNCH_OK();
      return 0;
    }
  }
  return fail("rgie_filter_fwd: unknown filter kind");
}

int rgie_filter_bwd(int kind, const float* in, const float* gout, float* gin, const float* p, int p_stride,
                    float* gp, int gp_stride, int B, int H, int W, float* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  RGIE_CHECK(B > 0 && H > 0 && W > 0 && ws != nullptr, "rgie_filter_bwd: bad arguments");
  LaunchShape s = shape_for(HW, B, true);
  float* partial = ws;
  switch (kind) {
    case RGIE_F_EXPOSURE: {
      ExposureOp op{p, p_stride, 0.f};
      if (int rc = launch_pointwise<ExposureOp, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 1, gp, gp_stride);
      break;
    }
    case RGIE_F_SATURATION: {
      SaturationOp op{p, p_stride, 0.f};
      if (int rc = launch_pointwise<SaturationOp, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 1, gp, gp_stride);
      break;
    }
    case RGIE_F_TONE: {
      CurveOp<1> op; op.p = p; op.stride = p_stride;
      if (int rc = launch_pointwise<CurveOp<1>, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 8, gp, gp_stride);
      break;
    }
    case RGIE_F_COLOR: {
      CurveOp<3> op; op.p = p; op.stride = p_stride;
      if (int rc = launch_pointwise<CurveOp<3>, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 24, gp, gp_stride);
      break;
    }
    case RGIE_F_CONTRAST: {
      // mean is recomputed (cheap) so that backward does not depend on forward-call state
      float* mu = ws + (long)B * kMaxBlk * 2;
      float* sums = mu + B;      // [B,2]
      dim3 grid(s.nblk, B);
      if (HW % 4 == 0) gray_sum_kernel<4><<<grid, kThreads, 0, st>>>(in, partial, HW, s.chunk);
      else gray_sum_kernel<1><<<grid, kThreads, 0, st>>>(in, partial, HW, s.chunk);
      RGIE_LAUNCH_OK();
      gray_mean_finalize<<<B, 32, 0, st>>>(partial, s.nblk, 1.0f / (float)HW, mu);
      RGIE_LAUNCH_OK();
      ContrastOp op{p, p_stride, mu, 0.f, 0.f};
      if (int rc = launch_pointwise<ContrastOp, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 2, sums, 2);
      RGIE_LAUNCH_OK();
      dim3 g2(plane_blocks(3 * HW), B);
      contrast_bwd_mean_kernel<<<g2, kThreads, 0, st>>>(gin, p, p_stride, sums, 2, HW);
      RGIE_LAUNCH_OK();
      copy_strided_kernel<<<ceil_div(B, 128), 128, 0, st>>>(sums, 2, gp, gp_stride, B);   // sums[b,0] -> gp[b]
      break;
    }
    case RGIE_F_GAMMA: {
      GammaOp op{p, p_stride, 0.f};
      if (int rc = launch_pointwise<GammaOp, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 1, gp, gp_stride);
      break;
    }
    case RGIE_F_BRIGHT: {
      BrightOp op{p, p_stride, 0.f};
      if (int rc = launch_pointwise<BrightOp, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 1, gp, gp_stride);
      break;
    }
    case RGIE_F_BW: {
      BwOp op{p, p_stride, 0.f};
      if (int rc = launch_pointwise<BwOp, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 1, gp, gp_stride);
      break;
    }
    case RGIE_F_HUE: {
      HueOp op{p, p_stride, 0.f};
      if (int rc = launch_pointwise<HueOp, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 1, gp, gp_stride);
      break;
    }
    case RGIE_F_WB: {
      // channel means are recomputed (cheap) so that backward does not depend on forward-call state
      float* mean = ws + (long)B * kMaxBlk * 4;   // [B,3]
      float* sums = mean + 3L * B;                // [B,4]
      dim3 grid(s.nblk, B);
      chan_sum_kernel<<<grid, kThreads, 0, st>>>(in, partial, HW, s.chunk);
      RGIE_LAUNCH_OK();
      chan_mean_finalize<<<B, 32, 0, st>>>(partial, s.nblk, 1.0f / (float)HW, mean);
      RGIE_LAUNCH_OK();
      WbOp op; op.p = p; op.stride = p_stride; op.mean_all = mean;
      if (int rc = launch_pointwise<WbOp, true>(in, gout, gin, op, partial, B, HW, st)) return rc;
      finalize_partials<<<B, 32, 0, st>>>(partial, s.nblk, 4, sums, 4);
      RGIE_LAUNCH_OK();
      dim3 g2(plane_blocks(3 * HW), B);
      wb_bwd_mean_kernel<<<g2, kThreads, 0, st>>>(gin, p, p_stride, mean, sums, HW);
      RGIE_LAUNCH_OK();
      copy_strided_kernel<<<ceil_div(B, 128), 128, 0, st>>>(sums, 4, gp, gp_stride, B);   // sums[b,0] -> gp[b]
      break;
    }
    case RGIE_F_SHARP: {
      RGIE_CHECK(H >= 3 && W >= 3, "sharp: image too small");
      const SharpMarch sm = sharp_march(H, W);
      if (sm.R > 0) {
        sharp_bwd_march_kernel<<<dim3(sm.nb, 3, B), sm.threads, sm.smem, st>>>(in, gout, gin, p, p_stride, partial, H, W, sm.R);
        RGIE_LAUNCH_OK();
        finalize_partials<<<B, 32, 0, st>>>(partial, 3 * sm.nb, 1, gp, gp_stride);
        break;
      }
      const int nb = plane_blocks(HW);
      float* gdeg = ws + (long)B * kMaxBlk * 24;
      dim3 grid(nb, 3, B);
      sharp_bwd_a_kernel<<<grid, kThreads, 0, st>>>(in, gout, gdeg, gin, p, p_stride, partial, H, W);
      RGIE_LAUNCH_OK();
      sharp_bwd_b_kernel<<<grid, kThreads, 0, st>>>(gdeg, gin, H, W);
      RGIE_LAUNCH_OK();
      finalize_partials<<<B, 32, 0, st>>>(partial, 3 * nb, 1, gp, gp_stride);
      break;
    }
    case RGIE_F_BLUR: {
      RGIE_CHECK(H > kRad && W > kRad, "blur: reflect padding needs H,W > 12");
      const int nb = plane_blocks(HW);
      const long plane = (long)B * 3 * HW;
      float* part2 = ws + (long)B * kMaxBlk * 8;
      float* wbuf = ws + (long)B * kMaxBlk * 16;   // [B, kWStride] (kWStride = 52 <= 8 * kMaxBlk)
      float* t = ws + (long)B * kMaxBlk * 24;
      float* td = t + plane;
      float* gm = td + plane;
      dim3 grid(nb, 3, B);
      blur_weights_kernel<<<B, 32, 0, st>>>(p, p_stride, wbuf);
      RGIE_LAUNCH_OK();
      blur_h_kernel<<<grid, kThreads, 0, st>>>(in, t, td, wbuf, gin, gout, H, W);
      RGIE_LAUNCH_OK();
      blur_bwd_mask_kernel<<<grid, kThreads, 0, st>>>(t, gout, gm, wbuf, partial, H, W);
      RGIE_LAUNCH_OK();
      // t is dead now: reuse it for V^T gm
      blur_bwd_t_kernel<0><<<grid, kThreads, 0, st>>>(gm, t, td, wbuf, part2, H, W);
      RGIE_LAUNCH_OK();
      blur_bwd_t_kernel<1><<<grid, kThreads, 0, st>>>(t, gin, nullptr, wbuf, nullptr, H, W);
      RGIE_LAUNCH_OK();
      // d(sigma) = sum(partial) + sum(part2): both live in one contiguous [B, 2*3*nb] view when nb == kMaxBlk*12/(3*nb)...
      // keep it simple: finalize each into a 2-float scratch then add
      float* two = gm;   // gm is dead after blur_bwd_t_kernel<0>; [B,2]
      finalize_partials<<<B, 32, 0, st>>>(partial, 3 * nb, 1, two, 2);
      finalize_partials<<<B, 32, 0, st>>>(part2, 3 * nb, 1, two + 1, 2);
      RGIE_LAUNCH_OK();
      finalize_partials<<<B, 32, 0, st>>>(two, 2, 1, gp, gp_stride);
      break;
    }
    case RGIE_F_SCALE: {
      RGIE_CHECK(H >= 2 && W >= 2, "scale: image too small");
      const dim3 gcol(ceil_div(W, kThreads), ceil_div(H, kScaleColRows), B);
      if (scale_col_on() && scale_tab_ok(H, W) && (long)gcol.x * gcol.y <= kMaxBlk * 6) {
        // pass A: column-marching kernel (masked gradient + four plain sums per block); pass B: table-driven gather
        const int nbt = ceil_div(H, kScaleRows);
        float* gmt = ws + (long)B * kMaxBlk * 24;
        scale_col_kernel<true><<<gcol, kThreads, 0, st>>>(in, gout, gmt, p, p_stride, partial, H, W);
        RGIE_LAUNCH_OK();
        scale_gather_tab_kernel<<<dim3(nbt, B), kThreads, scale_gather_smem(H, W), st>>>(gmt, gin, p, p_stride, H, W);
        RGIE_LAUNCH_OK();
        finalize_partials<<<B, 32, 0, st>>>(partial, (int)(gcol.x * gcol.y), 4, gp, gp_stride);
        break;
      }
      const int nb = plane_blocks(HW);
      float* gm = ws + (long)B * kMaxBlk * 24;
      dim3 grid(nb, B);
      scale_kernel<true><<<grid, kThreads, 0, st>>>(in, gout, gm, p, p_stride, partial, H, W);
      RGIE_LAUNCH_OK();
      scale_bwd_gather_kernel<<<grid, kThreads, 0, st>>>(gm, gin, p, p_stride, H, W);
      RGIE_LAUNCH_OK();
      finalize_partials<<<B, 32, 0, st>>>(partial, nb, 4, gp, gp_stride);
      break;
    }
    case RGIE_F_AFFINE: {
      RGIE_CHECK(H >= 2 && W >= 2, "affine: image too small");
      const int nb = plane_blocks(HW);
      float* dtheta = ws + (long)B * kMaxBlk * 8;      // [B, 6]
      dim3 grid(nb, B);
      RGIE_CUDA_OK(cudaMemsetAsync(gin, 0, (size_t)B * 3 * HW * sizeof(float), st));
      affine_kernel<true><<<grid, kThreads, 0, st>>>(in, gout, gin, p, p_stride, partial, H, W);
      RGIE_LAUNCH_OK();
      finalize_partials<<<B, 32, 0, st>>>(partial, nb, 6, dtheta, 6);
      RGIE_LAUNCH_OK();
      affine_param_grad_kernel<<<ceil_div(B, 128), 128, 0, st>>>(dtheta, p, p_stride, gp, gp_stride, B, H, W);
      break;
    }
    default:
      return fail("rgie_filter_bwd: unknown filter kind");
  }
  RGIE_LAUNCH_OK();
  return 0;
}

// exposure -> saturation -> tone -> colour in one pass each way (see PrefixOp).  `p` points at image 0's exposure value,
// followed by saturation, the 8 tone and the 24 colour values (the layout of the reference's default parameter vector);
// backward writes the 34 parameter gradients only -- the chain's input is the fixed original image, nothing consumes d(in).
int rgie_filter_prefix_fwd(const float* in, float* out, const float* p, int p_stride, int B, int H, int W, void* stream) {
  RGIE_CHECK(B > 0 && H > 0 && W > 0, "rgie_filter_prefix_fwd: bad shape");
  const int HW = H * W;
  LaunchShape s = shape_for(HW);
  dim3 grid(s.nblk, B);
  PrefixOp op = make_prefix_op(p, p_stride);
  cudaStream_t st = (cudaStream_t)stream;
  if (HW % 4 == 0) pointwise_kernel<PrefixOp, 4, false, true, 3><<<grid, kThreads, 0, st>>>(in, nullptr, out, op, nullptr, HW, s.chunk);
  else pointwise_kernel<PrefixOp, 1, false, true, 3><<<grid, kThreads, 0, st>>>(in, nullptr, out, op, nullptr, HW, s.chunk);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_filter_prefix_bwd(const float* in, const float* gout, const float* p, int p_stride, float* gp, int gp_stride,
                           int B, int H, int W, float* ws, void* stream) {
  RGIE_CHECK(B > 0 && H > 0 && W > 0 && ws != nullptr, "rgie_filter_prefix_bwd: bad arguments");
  const int HW = H * W;
  LaunchShape s = shape_for(HW, B, true);
  dim3 grid(s.nblk, B);
  PrefixOp op = make_prefix_op(p, p_stride);
  cudaStream_t st = (cudaStream_t)stream;
  if (HW % 4 == 0) pointwise_kernel<PrefixOp, 4, true, false, 2><<<grid, kThreads, 0, st>>>(in, gout, nullptr, op, ws, HW, s.chunk);
  else pointwise_kernel<PrefixOp, 1, true, false, 2><<<grid, kThreads, 0, st>>>(in, gout, nullptr, op, ws, HW, s.chunk);
  RGIE_LAUNCH_OK();
  finalize_partials<<<B, 64, 0, st>>>(ws, s.nblk, PrefixOp::NP, gp, gp_stride);
  RGIE_LAUNCH_OK();
  return 0;
}

}  // extern "C"

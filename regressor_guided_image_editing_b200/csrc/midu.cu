// MiDU guidance head (Stable-Diffusion variant): src/guidance_classifier/MiduClassifier.py:145-160
//   Conv(1280->256,3,p1) ReLU MaxPool2 Conv(256->128,3,p1) ReLU AdaptiveAvgPool(2,2) Flatten Linear(512,64) ReLU Linear(64,n)
// forward + input-gradient backward (d score / d feature), so the caller's autograd continues into its UNet
// (pipelines/InversionResamplingStableDiffusionPipeline.py:132-134).  The two convolutions run on the row-shifted GEMM
// (tcgen05 in bf16 mode, CUDA cores in fp32 parity mode); the tail is one small kernel per direction.
#include <string.h>
#include <vector>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "rgie.h"

namespace rgie {
namespace {

// feat NCHW fp32 -> padded NHWC T
template <typename T>
__global__ void __launch_bounds__(256) midu_pack_kernel(const float* __restrict__ feat, T* __restrict__ x, Geom g, int C, long total) {
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long q = idx / C;
    const int j = (int)(q % g.W); q /= g.W;
    const int i = (int)(q % g.H);
    const int n = (int)(q / g.H);
    x[geom_row(g, 0, n, i, j) * C + c] = from_f<T>(feat[(((long)n * C + c) * g.H + i) * g.W + j]);
  }
}
// d(feat): padded NHWC fp32 -> NCHW fp32
__global__ void __launch_bounds__(256) midu_unpack_kernel(const float* __restrict__ dx, float* __restrict__ dfeat, Geom g, int C, long total) {
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % g.W);
    long q = idx / g.W;
    const int i = (int)(q % g.H); q /= g.H;
    const int c = (int)(q % C);
    const int n = (int)(q / C);
    dfeat[idx] = dx[geom_row(g, 0, n, i, j) * C + c];
  }
}
// 2x2 stride-2 max pool (first max wins) between two padded layouts
template <typename T>
__global__ void __launch_bounds__(256) pool2_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, uint8_t* __restrict__ arg,
                                                       Geom gi, Geom go, int C, long total) {
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long q = idx / C;
    const int j = (int)(q % go.W); q /= go.W;
    const int i = (int)(q % go.H);
    const int n = (int)(q / go.H);
    float best = 0.f; int code = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v = to_f<T>(in[geom_row(gi, 0, n, 2 * i + (k >> 1), 2 * j + (k & 1)) * C + c]);
      if (k == 0 || v > best) { best = v; code = k; }
    }
    out[geom_row(go, 0, n, i, j) * C + c] = from_f<T>(best);
    arg[idx] = (uint8_t)code;
  }
}
// dIn[n,y,x,c] = (in > 0) * (argmax of its window == this position ? dOut : 0)
template <typename T>
__global__ void __launch_bounds__(256) pool2_bwd_kernel(const T* __restrict__ dout, const uint8_t* __restrict__ arg,
                                                       const T* __restrict__ in, T* __restrict__ din, Geom gi, Geom go, int C,
                                                       long total) {
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long q = idx / C;
    const int x = (int)(q % gi.W); q /= gi.W;
    const int y = (int)(q % gi.H);
    const int n = (int)(q / gi.H);
    const long r = geom_row(gi, 0, n, y, x) * C + c;
    const int i = y >> 1, j = x >> 1, code = (y & 1) * 2 + (x & 1);
    float g = 0.f;
    if (to_f<T>(in[r]) > 0.f && arg[(((long)n * go.H + i) * go.W + j) * C + c] == code)
      g = to_f<T>(dout[geom_row(go, 0, n, i, j) * C + c]);
    din[r] = from_f<T>(g);
  }
}
// AdaptiveAvgPool(2,2) over the HxH map (H even) + Flatten (c*4 + i*2 + j) + Linear(512,64) + ReLU + Linear(64,n_out)
template <typename T>
__global__ void __launch_bounds__(256) midu_tail_fwd_kernel(const T* __restrict__ h, Geom g, int C, const float* __restrict__ w7,
                                                           const float* __restrict__ b7, const float* __restrict__ w9,
                                                           const float* __restrict__ b9, int n_out, float* __restrict__ v_out,
                                                           float* __restrict__ u_out, float* __restrict__ pred) {
  __shared__ float v[512];
  __shared__ float u[64];
  const int n = blockIdx.x;
  const int hb = g.H / 2;
  for (int e = threadIdx.x; e < C * 4; e += blockDim.x) {
    const int c = e >> 2, i = (e >> 1) & 1, j = e & 1;
    float s = 0.f;
    for (int a = 0; a < hb; ++a)
      for (int b = 0; b < hb; ++b) s += to_f<T>(h[geom_row(g, 0, n, i * hb + a, j * hb + b) * C + c]);
    v[e] = s / (float)(hb * hb);
    v_out[(long)n * C * 4 + e] = v[e];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = b7[threadIdx.x];
    for (int e = 0; e < C * 4; ++e) s = fmaf(w7[(long)threadIdx.x * C * 4 + e], v[e], s);
    u[threadIdx.x] = fmaxf(s, 0.f);
    u_out[(long)n * 64 + threadIdx.x] = u[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x < n_out) {
    float s = b9[threadIdx.x];
    for (int e = 0; e < 64; ++e) s = fmaf(w9[threadIdx.x * 64 + e], u[e], s);
    pred[(long)n * n_out + threadIdx.x] = s;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) midu_tail_bwd_kernel(const float* __restrict__ dpred, int n_out, const float* __restrict__ w9,
                                                           const float* __restrict__ w7, const float* __restrict__ u,
                                                           const T* __restrict__ h, T* __restrict__ dh, Geom g, int C) {
  __shared__ float du[64];
  __shared__ float dv[512];
  const int n = blockIdx.x;
  const int hb = g.H / 2;
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int k = 0; k < n_out; ++k) s = fmaf(w9[k * 64 + threadIdx.x], dpred[(long)n * n_out + k], s);
    du[threadIdx.x] = u[(long)n * 64 + threadIdx.x] > 0.f ? s : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < C * 4; e += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < 64; ++k) s = fmaf(w7[(long)k * C * 4 + e], du[k], s);
    dv[e] = s / (float)(hb * hb);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < C * g.H * g.W; e += blockDim.x) {
    const int c = e % C;
    const int p = e / C;
    const int y = p / g.W, x = p % g.W;
    const long r = geom_row(g, 0, n, y, x) * C + c;
    const float gv = dv[c * 4 + (y / hb) * 2 + (x / hb)];
    dh[r] = from_f<T>(to_f<T>(h[r]) > 0.f ? gv : 0.f);
  }
}

int grid_for(long total) {
  long g = (total + 255) / 256;
  const long cap = 148L * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace
}  // namespace rgie

using namespace rgie;

struct RgieMiduHead {
  int precision = 0, dtype = 0, esz = 4;
  int B = 0, hw = 0, n_out = 2;
  Geom gA, gB;
  std::vector<void*> allocs;
  void *w0 = nullptr, *w0t = nullptr, *w3 = nullptr, *w3t = nullptr;
  float *b0 = nullptr, *b3 = nullptr, *w7 = nullptr, *b7 = nullptr, *w9 = nullptr, *b9 = nullptr;
  void *x = nullptr, *h0 = nullptr, *p0 = nullptr, *h1 = nullptr;
  void *dh1 = nullptr, *dp0 = nullptr, *dh0 = nullptr;
  float *dx = nullptr, *v = nullptr, *u = nullptr;
  uint8_t* arg = nullptr;
  GemmDesc d[4];
  GemmPlanSm100 plan[4];
};

namespace {
int midu_alloc(RgieMiduHead* M, void** p, size_t bytes) {
  RGIE_CUDA_OK(cudaMalloc(p, bytes ? bytes : 16));
  RGIE_CUDA_OK(cudaMemset(*p, 0, bytes ? bytes : 16));
  M->allocs.push_back(*p);
  return 0;
}
int midu_upload(RgieMiduHead* M, const std::vector<float>& h, void** dptr) {
  if (int rc = midu_alloc(M, dptr, h.size() * M->esz)) return rc;
  if (M->dtype == 0) {
    RGIE_CUDA_OK(cudaMemcpy(*dptr, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  } else {
    std::vector<__nv_bfloat16> hb(h.size());
    for (size_t i = 0; i < h.size(); ++i) hb[i] = __float2bfloat16(h[i]);
    RGIE_CUDA_OK(cudaMemcpy(*dptr, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  }
  return 0;
}
int midu_upload_f32(RgieMiduHead* M, const float* h, size_t n, float** dptr) {
  if (int rc = midu_alloc(M, (void**)dptr, n * 4)) return rc;
  RGIE_CUDA_OK(cudaMemcpy(*dptr, h, n * 4, cudaMemcpyHostToDevice));
  return 0;
}
// PyTorch conv weight [co, ci, 3, 3] -> forward [co, 9*ci] (t = r*3+s) and dgrad [ci, 9*co]
void pack3x3(const float* w, int co, int ci, std::vector<float>& f, std::vector<float>& t) {
  f.assign((size_t)co * 9 * ci, 0.f);
  t.assign((size_t)ci * 9 * co, 0.f);
  for (int n = 0; n < co; ++n)
    for (int c = 0; c < ci; ++c)
      for (int k = 0; k < 9; ++k) {
        const float v = w[((size_t)n * ci + c) * 9 + k];
        f[(size_t)n * 9 * ci + (size_t)k * ci + c] = v;
        const int kk = (k / 3) * 3 + (2 - k % 3);          // dgrad tap slot: taps of a kernel row reversed (ascending offsets)
        t[(size_t)c * 9 * co + (size_t)kk * co + n] = v;
      }
}
int midu_run(RgieMiduHead* M, int i, cudaStream_t st) {
  if (M->precision == RGIE_PREC_BF16) return run_gemm_sm100(M->plan[i], st);
  return launch_gemm_simt(M->d[i], M->dtype, st);
}
}  // namespace

extern "C" {

void rgie_midu_destroy(RgieMiduHead* M) {
  if (!M) return;
  for (void* p : M->allocs) cudaFree(p);
  delete M;
}

int rgie_midu_create(const float* const* h_tensors, int n_tensors, int n_out, int max_batch, int hw, int precision,
                     RgieMiduHead** out) {
  RGIE_CHECK(h_tensors && out, "rgie_midu_create: null argument");
  RGIE_CHECK(n_tensors == 8, "rgie_midu_create: expected 8 tensors (0.w,0.b,3.w,3.b,7.w,7.b,9.w,9.b)");
  RGIE_CHECK(hw == 8, "rgie_midu_create: the SD head takes 8x8 mid-block features (SDXL variant: SURVEY.md 8f rank 4)");
  RGIE_CHECK(n_out >= 1 && n_out <= 64 && max_batch >= 1 && precision >= 0 && precision <= 2, "rgie_midu_create: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("rgie_midu_create: no CUDA device (there is no CPU fallback)");
  RgieMiduHead* M = new RgieMiduHead();
  struct Guard { RgieMiduHead* m; bool ok = false; ~Guard() { if (!ok) rgie_midu_destroy(m); } } guard{M};
  M->precision = precision; M->dtype = precision == RGIE_PREC_FP32 ? 0 : 1; M->esz = M->dtype == 0 ? 4 : 2;
  M->B = max_batch; M->hw = hw; M->n_out = n_out;
  const int B = max_batch, esz = M->esz;
  M->gA = make_geom(1, B, hw, hw, 1, 1, 1, 1);
  M->gB = make_geom(1, B, hw / 2, hw / 2, 1, 1, 1, 1);
  std::vector<float> f, t;
  pack3x3(h_tensors[0], 256, 1280, f, t);
  if (int rc = midu_upload(M, f, &M->w0)) return rc;
  if (int rc = midu_upload(M, t, &M->w0t)) return rc;
  pack3x3(h_tensors[2], 128, 256, f, t);
  if (int rc = midu_upload(M, f, &M->w3)) return rc;
  if (int rc = midu_upload(M, t, &M->w3t)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[1], 256, &M->b0)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[3], 128, &M->b3)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[4], 64 * 512, &M->w7)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[5], 64, &M->b7)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[6], (size_t)n_out * 64, &M->w9)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[7], n_out, &M->b9)) return rc;
  const long rA = M->gA.rows(), rB = M->gB.rows();
  if (int rc = midu_alloc(M, &M->x, (size_t)rA * 1280 * esz)) return rc;
  if (int rc = midu_alloc(M, &M->h0, (size_t)rA * 256 * esz)) return rc;
  if (int rc = midu_alloc(M, &M->p0, (size_t)rB * 256 * esz)) return rc;
  if (int rc = midu_alloc(M, &M->h1, (size_t)rB * 128 * esz)) return rc;
  if (int rc = midu_alloc(M, &M->dh1, (size_t)rB * 128 * esz)) return rc;
  if (int rc = midu_alloc(M, &M->dp0, (size_t)rB * 256 * esz)) return rc;
  if (int rc = midu_alloc(M, &M->dh0, (size_t)rA * 256 * esz)) return rc;
  if (int rc = midu_alloc(M, (void**)&M->dx, (size_t)rA * 1280 * 4)) return rc;
  if (int rc = midu_alloc(M, (void**)&M->v, (size_t)B * 512 * 4)) return rc;
  if (int rc = midu_alloc(M, (void**)&M->u, (size_t)B * 64 * 4)) return rc;
  if (int rc = midu_alloc(M, (void**)&M->arg, (size_t)B * (hw / 2) * (hw / 2) * 256)) return rc;

  auto conv = [&](GemmDesc& d, const void* A, const Geom& g, int ci, const void* W, int co, float* bias, int relu,
                  const void* mask, int ld_mask, void* D, int d_fp32, bool transpose) {
    memset(&d, 0, sizeof(d));
    d.A = A; d.a_rows = g.rows(); d.Cin = ci; d.Wt = W; d.n_pad = co; d.ntaps = 9;
    for (int k = 0; k < 9; ++k) {
      if (!transpose) d.row_off[k] = (long)(k / 3 - 1) * g.P + (k % 3 - 1);
      else {
        const int t = (k / 3) * 3 + (2 - k % 3);             // slot k holds kernel tap t (see pack3x3)
        d.row_off[k] = -((long)(t / 3 - 1) * g.P + (t % 3 - 1));
      }
    }
    d.m_begin = 0; d.m_end = g.rows(); d.Cout = co;
    d.src = g; d.dst_kind = DST_SAME; d.dst = g; d.D = D; d.ldd = co; d.d_fp32 = d_fp32;
    d.bias = bias; d.relu = relu; d.mask = mask; d.ld_mask = ld_mask;
  };
  conv(M->d[0], M->x, M->gA, 1280, M->w0, 256, M->b0, 1, nullptr, 0, M->h0, 0, false);
  conv(M->d[1], M->p0, M->gB, 256, M->w3, 128, M->b3, 1, nullptr, 0, M->h1, 0, false);
  conv(M->d[2], M->dh1, M->gB, 128, M->w3t, 256, nullptr, 0, nullptr, 0, M->dp0, 0, true);
  conv(M->d[3], M->dh0, M->gA, 256, M->w0t, 1280, nullptr, 0, nullptr, 0, M->dx, 1, true);
  if (precision == RGIE_PREC_BF16)
    for (int i = 0; i < 4; ++i)
      if (int rc = build_gemm_sm100(M->d[i], &M->plan[i])) return rc;
  RGIE_CUDA_OK(cudaDeviceSynchronize());
  guard.ok = true;
  *out = M;
  return 0;
}

int rgie_midu_forward(RgieMiduHead* M, const float* feat, int B, float* pred, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RGIE_CHECK(M && feat && pred, "rgie_midu_forward: null argument");
  RGIE_CHECK(B == M->B, "rgie_midu_forward: batch must equal the max_batch the handle was created with");
  const long t_pack = (long)B * M->hw * M->hw * 1280;
  const long t_pool = (long)B * (M->hw / 2) * (M->hw / 2) * 256;
  if (M->dtype == 0) midu_pack_kernel<float><<<grid_for(t_pack), 256, 0, st>>>(feat, (float*)M->x, M->gA, 1280, t_pack);
  else midu_pack_kernel<__nv_bfloat16><<<grid_for(t_pack), 256, 0, st>>>(feat, (__nv_bfloat16*)M->x, M->gA, 1280, t_pack);
  RGIE_LAUNCH_OK();
  if (int rc = midu_run(M, 0, st)) return rc;
  if (M->dtype == 0) pool2_fwd_kernel<float><<<grid_for(t_pool), 256, 0, st>>>((const float*)M->h0, (float*)M->p0, M->arg, M->gA, M->gB, 256, t_pool);
  else pool2_fwd_kernel<__nv_bfloat16><<<grid_for(t_pool), 256, 0, st>>>((const __nv_bfloat16*)M->h0, (__nv_bfloat16*)M->p0, M->arg, M->gA, M->gB, 256, t_pool);
  RGIE_LAUNCH_OK();
  if (int rc = midu_run(M, 1, st)) return rc;
  if (M->dtype == 0) midu_tail_fwd_kernel<float><<<B, 256, 0, st>>>((const float*)M->h1, M->gB, 128, M->w7, M->b7, M->w9, M->b9, M->n_out, M->v, M->u, pred);
  else midu_tail_fwd_kernel<__nv_bfloat16><<<B, 256, 0, st>>>((const __nv_bfloat16*)M->h1, M->gB, 128, M->w7, M->b7, M->w9, M->b9, M->n_out, M->v, M->u, pred);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_midu_backward(RgieMiduHead* M, const float* dpred, float* dfeat, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RGIE_CHECK(M && dpred && dfeat, "rgie_midu_backward: null argument");
  const int B = M->B;
  const long t_in = (long)B * M->hw * M->hw * 256;
  const long t_un = (long)B * 1280 * M->hw * M->hw;
  if (M->dtype == 0) midu_tail_bwd_kernel<float><<<B, 256, 0, st>>>(dpred, M->n_out, M->w9, M->w7, M->u, (const float*)M->h1, (float*)M->dh1, M->gB, 128);
  else midu_tail_bwd_kernel<__nv_bfloat16><<<B, 256, 0, st>>>(dpred, M->n_out, M->w9, M->w7, M->u, (const __nv_bfloat16*)M->h1, (__nv_bfloat16*)M->dh1, M->gB, 128);
  RGIE_LAUNCH_OK();
  if (int rc = midu_run(M, 2, st)) return rc;
  if (M->dtype == 0) pool2_bwd_kernel<float><<<grid_for(t_in), 256, 0, st>>>((const float*)M->dp0, M->arg, (const float*)M->h0, (float*)M->dh0, M->gA, M->gB, 256, t_in);
  else pool2_bwd_kernel<__nv_bfloat16><<<grid_for(t_in), 256, 0, st>>>((const __nv_bfloat16*)M->dp0, M->arg, (const __nv_bfloat16*)M->h0, (__nv_bfloat16*)M->dh0, M->gA, M->gB, 256, t_in);
  RGIE_LAUNCH_OK();
  if (int rc = midu_run(M, 3, st)) return rc;
  midu_unpack_kernel<<<grid_for(t_un), 256, 0, st>>>(M->dx, dfeat, M->gA, 1280, t_un);
  RGIE_LAUNCH_OK();
  return 0;
}

}  // extern "C"

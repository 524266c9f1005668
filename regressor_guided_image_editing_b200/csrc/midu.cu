// MiDU guidance head: src/guidance_classifier/MiduClassifier.py:121-161
//   SD   (:145-160): Conv(1280->256,3,p1) ReLU MaxPool2 Conv(256->128,3,p1) ReLU AdaptiveAvgPool(2,2) Flatten
//                    Linear(512,64) ReLU Linear(64,n)                                  on 8x8 mid-block features
//   SDXL (:125-143): [Conv(3,p1) ReLU MaxPool2] x4 with 1280->512->256->128->64, Flatten, Linear(256,128) ReLU
//                    Linear(128,n)                                                     on 32x32 mid-block features
// Both are a list of conv stages (3x3 conv + ReLU, optional 2x2 max pool) followed by one tail kernel
// (2x2 adaptive average + flatten + two linears).
// forward + input-gradient backward (d score / d feature), so the caller's autograd continues into its UNet
// (pipelines/InversionResamplingStableDiffusionPipeline.py:132-134).  The two convolutions run on the row-shifted GEMM
// (tcgen05 in bf16 mode, CUDA cores in fp32 parity mode); the tail is one small kernel per direction.
#include <string.h>
#include <vector>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "gemm_tc32.cuh"
#include "rgie.h"
#include <stdlib.h>

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
// AdaptiveAvgPool(2,2) over the HxH map (H even; H == 2: identity) + Flatten (c*4 + i*2 + j) + Linear(4C, NH) + ReLU +
// Linear(NH, n_out)                                                           (SD: C = 128, NH = 64; SDXL: C = 64, NH = 128)
template <typename T>
__global__ void __launch_bounds__(256) midu_tail_fwd_kernel(const T* __restrict__ h, Geom g, int C, int NH,
                                                           const float* __restrict__ w7,
                                                           const float* __restrict__ b7, const float* __restrict__ w9,
                                                           const float* __restrict__ b9, int n_out, float* __restrict__ v_out,
                                                           float* __restrict__ u_out, float* __restrict__ pred) {
  __shared__ float v[512];
  __shared__ float u[128];
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
  if (threadIdx.x < NH) {
    float s = b7[threadIdx.x];
    for (int e = 0; e < C * 4; ++e) s = fmaf(w7[(long)threadIdx.x * C * 4 + e], v[e], s);
    u[threadIdx.x] = fmaxf(s, 0.f);
    u_out[(long)n * NH + threadIdx.x] = u[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x < n_out) {
    float s = b9[threadIdx.x];
    for (int e = 0; e < NH; ++e) s = fmaf(w9[threadIdx.x * NH + e], u[e], s);
    pred[(long)n * n_out + threadIdx.x] = s;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) midu_tail_bwd_kernel(const float* __restrict__ dpred, int n_out, const float* __restrict__ w9,
                                                           const float* __restrict__ w7, const float* __restrict__ u,
                                                           const T* __restrict__ h, T* __restrict__ dh, Geom g, int C, int NH) {
  __shared__ float du[128];
  __shared__ float dv[512];
  const int n = blockIdx.x;
  const int hb = g.H / 2;
  if (threadIdx.x < NH) {
    float s = 0.f;
    for (int k = 0; k < n_out; ++k) s = fmaf(w9[k * NH + threadIdx.x], dpred[(long)n * n_out + k], s);
    du[threadIdx.x] = u[(long)n * NH + threadIdx.x] > 0.f ? s : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < C * 4; e += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < NH; ++k) s = fmaf(w7[(long)k * C * 4 + e], du[k], s);
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

constexpr int kMaxStages = 4;
struct MiduStage {
  int ci = 0, co = 0, hw = 0;
  bool pool = false;
  Geom g, gp;                      // conv geometry / pooled geometry
  void *w = nullptr, *wt = nullptr;
  float* b = nullptr;
  void *in = nullptr, *h = nullptr, *p = nullptr;      // padded NHWC: conv input, conv output (post ReLU), pooled output
  void *dh = nullptr, *dp = nullptr;                   // gradients w.r.t. h / p
  uint8_t* arg = nullptr;
  GemmDesc d_fwd, d_bwd;
  GemmPlanSm100 p_fwd, p_bwd;
  GemmPlanTc32 t_fwd, t_bwd;       // fp32 mode: bf16x3 tensor-core plans
};

struct RgieMiduHead {
  int precision = 0, dtype = 0, esz = 4;
  int tc32 = 0;                    // fp32 mode on the tensor cores (gemm_tc32.cu); RGIE_FP32_SIMT=1 keeps the CUDA-core GEMM
  int B = 0, hw = 0, n_out = 2, n_stages = 0, C_tail = 0, NH = 0;
  MiduStage st[kMaxStages];
  Geom g_tail;
  void* tail_in = nullptr;         // last stage's p (pooled) or h
  void* d_tail_in = nullptr;
  std::vector<void*> allocs;
  float *w7 = nullptr, *b7 = nullptr, *w9 = nullptr, *b9 = nullptr;
  float *dx = nullptr, *v = nullptr, *u = nullptr;
};

namespace {
int midu_alloc(RgieMiduHead* M, void** p, size_t bytes) {
  RGIE_CUDA_OK(cudaMalloc(p, bytes ? bytes : 16));
  RGIE_CUDA_OK(cudaMemset(*p, 0, bytes ? bytes : 16));
  M->allocs.push_back(*p);
  return 0;
}
int midu_upload(RgieMiduHead* M, const std::vector<float>& h, void** dptr) {
  if (M->tc32) {
    std::vector<__nv_bfloat16> planes(3 * h.size());
    split_weights_bf16x3(h.data(), h.size(), planes.data());
    if (int rc = midu_alloc(M, dptr, planes.size() * 2)) return rc;
    RGIE_CUDA_OK(cudaMemcpy(*dptr, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice));
    return 0;
  }
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
void conv_desc(GemmDesc& d, const void* A, const Geom& g, int ci, const void* W, int co, float* bias, int relu, void* D,
               int d_fp32, bool transpose) {
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
  d.bias = bias; d.relu = relu;
}
int midu_run(RgieMiduHead* M, const GemmDesc& d, const GemmPlanSm100& plan, const GemmPlanTc32& tplan, cudaStream_t st) {
  if (M->precision == RGIE_PREC_BF16) return run_gemm_sm100(plan, st);
  if (M->tc32) return run_gemm_tc32(tplan, st);
  return launch_gemm_simt(d, M->dtype, st);
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
  RGIE_CHECK((n_tensors == 8 && hw == 8) || (n_tensors == 12 && hw == 32),
             "rgie_midu_create: expected 8 tensors on 8x8 features (SD head) or 12 tensors on 32x32 features (SDXL head)");
  RGIE_CHECK(n_out >= 1 && n_out <= 64 && max_batch >= 1 && precision >= 0 && precision <= 2, "rgie_midu_create: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("rgie_midu_create: no CUDA device (there is no CPU fallback)");
  RgieMiduHead* M = new RgieMiduHead();
  struct Guard { RgieMiduHead* m; bool ok = false; ~Guard() { if (!ok) rgie_midu_destroy(m); } } guard{M};
  M->precision = precision; M->dtype = precision == RGIE_PREC_FP32 ? 0 : 1; M->esz = M->dtype == 0 ? 4 : 2;
  static const int env_simt = getenv("RGIE_FP32_SIMT") ? atoi(getenv("RGIE_FP32_SIMT")) : 0;
  M->tc32 = (precision == RGIE_PREC_FP32 && !env_simt) ? 1 : 0;
  M->B = max_batch; M->hw = hw; M->n_out = n_out;
  const int B = max_batch, esz = M->esz;
  const bool sdxl = n_tensors == 12;
  // stage list
  const int chans_sd[3] = {1280, 256, 128}, chans_xl[5] = {1280, 512, 256, 128, 64};
  const int* ch = sdxl ? chans_xl : chans_sd;
  M->n_stages = sdxl ? 4 : 2;
  int cur = hw;
  for (int s = 0; s < M->n_stages; ++s) {
    MiduStage& S = M->st[s];
    S.ci = ch[s]; S.co = ch[s + 1]; S.hw = cur;
    S.pool = sdxl || s == 0;                         // SD: only the first conv is followed by a max pool
    S.g = make_geom(1, B, cur, cur, 1, 1, 1, 1);
    if (S.pool) { cur /= 2; S.gp = make_geom(1, B, cur, cur, 1, 1, 1, 1); }
    std::vector<float> f, t;
    pack3x3(h_tensors[2 * s], S.co, S.ci, f, t);
    if (int rc = midu_upload(M, f, &S.w)) return rc;
    if (int rc = midu_upload(M, t, &S.wt)) return rc;
    if (int rc = midu_upload_f32(M, h_tensors[2 * s + 1], S.co, &S.b)) return rc;
  }
  M->C_tail = ch[M->n_stages];
  M->NH = sdxl ? 128 : 64;
  const int t0 = 2 * M->n_stages;
  if (int rc = midu_upload_f32(M, h_tensors[t0], (size_t)M->NH * M->C_tail * 4, &M->w7)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[t0 + 1], M->NH, &M->b7)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[t0 + 2], (size_t)n_out * M->NH, &M->w9)) return rc;
  if (int rc = midu_upload_f32(M, h_tensors[t0 + 3], n_out, &M->b9)) return rc;
  // buffers
  for (int s = 0; s < M->n_stages; ++s) {
    MiduStage& S = M->st[s];
    const long r = S.g.rows();
    if (s == 0) { if (int rc = midu_alloc(M, &S.in, (size_t)r * S.ci * esz)) return rc; }
    else S.in = M->st[s - 1].pool ? M->st[s - 1].p : M->st[s - 1].h;
    if (int rc = midu_alloc(M, &S.h, (size_t)r * S.co * esz)) return rc;
    if (int rc = midu_alloc(M, &S.dh, (size_t)r * S.co * esz)) return rc;
    if (S.pool) {
      const long rp = S.gp.rows();
      if (int rc = midu_alloc(M, &S.p, (size_t)rp * S.co * esz)) return rc;
      if (int rc = midu_alloc(M, &S.dp, (size_t)rp * S.co * esz)) return rc;
      if (int rc = midu_alloc(M, (void**)&S.arg, (size_t)B * S.gp.H * S.gp.W * S.co)) return rc;
    }
  }
  const MiduStage& L = M->st[M->n_stages - 1];
  M->g_tail = L.pool ? L.gp : L.g;
  M->tail_in = L.pool ? L.p : L.h;
  M->d_tail_in = L.pool ? L.dp : L.dh;
  if (int rc = midu_alloc(M, (void**)&M->dx, (size_t)M->st[0].g.rows() * 1280 * 4)) return rc;
  if (int rc = midu_alloc(M, (void**)&M->v, (size_t)B * 512 * 4)) return rc;
  if (int rc = midu_alloc(M, (void**)&M->u, (size_t)B * 128 * 4)) return rc;
  // GEMM descriptors: forward conv + ReLU; backward: d(conv input) = conv^T(d h); its destination is the previous stage's
  // dp (pooled gradient) / dh, or the fp32 d(feature) buffer for the first stage
  for (int s = 0; s < M->n_stages; ++s) {
    MiduStage& S = M->st[s];
    conv_desc(S.d_fwd, S.in, S.g, S.ci, S.w, S.co, S.b, 1, S.h, 0, false);
    void* dst = s == 0 ? (void*)M->dx : (M->st[s - 1].pool ? M->st[s - 1].dp : M->st[s - 1].dh);
    conv_desc(S.d_bwd, S.dh, S.g, S.co, S.wt, S.ci, nullptr, 0, dst, s == 0 ? 1 : 0, true);
    if (s > 0 && !M->st[s - 1].pool) {
      // no pool in between: the previous ReLU's mask is applied by this GEMM's epilogue
      S.d_bwd.mask = M->st[s - 1].h; S.d_bwd.ld_mask = M->st[s - 1].co;
    }
    if (precision == RGIE_PREC_BF16) {
      if (int rc = build_gemm_sm100(S.d_fwd, &S.p_fwd)) return rc;
      if (int rc = build_gemm_sm100(S.d_bwd, &S.p_bwd)) return rc;
    } else if (M->tc32) {
      if (int rc = build_gemm_tc32(S.d_fwd, &S.t_fwd)) return rc;
      if (int rc = build_gemm_tc32(S.d_bwd, &S.t_bwd)) return rc;
    }
  }
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
  if (M->dtype == 0) midu_pack_kernel<float><<<grid_for(t_pack), 256, 0, st>>>(feat, (float*)M->st[0].in, M->st[0].g, 1280, t_pack);
  else midu_pack_kernel<__nv_bfloat16><<<grid_for(t_pack), 256, 0, st>>>(feat, (__nv_bfloat16*)M->st[0].in, M->st[0].g, 1280, t_pack);
  RGIE_LAUNCH_OK();
  for (int s = 0; s < M->n_stages; ++s) {
    MiduStage& S = M->st[s];
    if (int rc = midu_run(M, S.d_fwd, S.p_fwd, S.t_fwd, st)) return rc;
    if (S.pool) {
      const long t_pool = (long)B * S.gp.H * S.gp.W * S.co;
      if (M->dtype == 0) pool2_fwd_kernel<float><<<grid_for(t_pool), 256, 0, st>>>((const float*)S.h, (float*)S.p, S.arg, S.g, S.gp, S.co, t_pool);
      else pool2_fwd_kernel<__nv_bfloat16><<<grid_for(t_pool), 256, 0, st>>>((const __nv_bfloat16*)S.h, (__nv_bfloat16*)S.p, S.arg, S.g, S.gp, S.co, t_pool);
      RGIE_LAUNCH_OK();
    }
  }
  if (M->dtype == 0) midu_tail_fwd_kernel<float><<<B, 256, 0, st>>>((const float*)M->tail_in, M->g_tail, M->C_tail, M->NH, M->w7, M->b7, M->w9, M->b9, M->n_out, M->v, M->u, pred);
  else midu_tail_fwd_kernel<__nv_bfloat16><<<B, 256, 0, st>>>((const __nv_bfloat16*)M->tail_in, M->g_tail, M->C_tail, M->NH, M->w7, M->b7, M->w9, M->b9, M->n_out, M->v, M->u, pred);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_midu_backward(RgieMiduHead* M, const float* dpred, float* dfeat, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RGIE_CHECK(M && dpred && dfeat, "rgie_midu_backward: null argument");
  const int B = M->B;
  // tail: d(tail input), masked by (tail input > 0) -- the ReLU of the last conv (a max pool of post-ReLU values is
  // positive exactly where its arg-max is)
  if (M->dtype == 0) midu_tail_bwd_kernel<float><<<B, 256, 0, st>>>(dpred, M->n_out, M->w9, M->w7, M->u, (const float*)M->tail_in, (float*)M->d_tail_in, M->g_tail, M->C_tail, M->NH);
  else midu_tail_bwd_kernel<__nv_bfloat16><<<B, 256, 0, st>>>(dpred, M->n_out, M->w9, M->w7, M->u, (const __nv_bfloat16*)M->tail_in, (__nv_bfloat16*)M->d_tail_in, M->g_tail, M->C_tail, M->NH);
  RGIE_LAUNCH_OK();
  for (int s = M->n_stages - 1; s >= 0; --s) {
    MiduStage& S = M->st[s];
    if (S.pool) {
      const long t_in = (long)B * S.hw * S.hw * S.co;
      if (M->dtype == 0) pool2_bwd_kernel<float><<<grid_for(t_in), 256, 0, st>>>((const float*)S.dp, S.arg, (const float*)S.h, (float*)S.dh, S.g, S.gp, S.co, t_in);
      else pool2_bwd_kernel<__nv_bfloat16><<<grid_for(t_in), 256, 0, st>>>((const __nv_bfloat16*)S.dp, S.arg, (const __nv_bfloat16*)S.h, (__nv_bfloat16*)S.dh, S.g, S.gp, S.co, t_in);
      RGIE_LAUNCH_OK();
    }
    if (int rc = midu_run(M, S.d_bwd, S.p_bwd, S.t_bwd, st)) return rc;
  }
  const long t_un = (long)B * 1280 * M->hw * M->hw;
  midu_unpack_kernel<<<grid_for(t_un), 256, 0, st>>>(M->dx, dfeat, M->st[0].g, 1280, t_un);
  RGIE_LAUNCH_OK();
  return 0;
}

}  // extern "C"

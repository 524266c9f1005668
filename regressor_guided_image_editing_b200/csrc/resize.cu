// Antialiased bilinear resize (ATen _upsample_bilinear2d_aa, align_corners=False) and its transpose.
// Reference call site: torchvision transforms.Resize(480, antialias=True) inside load_model_eval
// (src/baselines/models/EmotionPredictionModel.py:36-37).
//
// Separable: horizontal pass into `tmp` [planes, in_h, out_w], then vertical pass.  Tap tables (per output index:
// first source index, tap count, normalised fp32 weights; and the CSR transpose for backward) are built once on the
// host with the same float arithmetic as ATen's `_compute_indices_weights_aa` and live in the handle.
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "rgie.h"

namespace rgie {

// output rows per block of the fused kernels: a group of R outputs needs ~R * in/out + taps source rows, so the halo (rows
// staged and filtered by two neighbouring blocks) shrinks from ~50 % at R = 8 to ~25 % at R = 16.  RGIE_RESIZE_ROWS overrides.
// threads per block of the fused kernels.  512 (one column sweep of a 480-wide row instead of two, twice the warps) was tried:
// 0.247 / 0.273 ms against 0.213 / 0.257 ms at 256 threads (64 x 3 planes, 512 -> 480) -- the block is bound by its
// load -> filter -> filter -> store phases, not by idle lanes.  RGIE_RESIZE_THREADS=512 keeps the experiment reachable.
static const int kFuseThreads = getenv("RGIE_RESIZE_THREADS") && atoi(getenv("RGIE_RESIZE_THREADS")) == 512 ? 512 : 256;
static const int kFuseRows = getenv("RGIE_RESIZE_ROWS") ? (atoi(getenv("RGIE_RESIZE_ROWS")) > 0 ? atoi(getenv("RGIE_RESIZE_ROWS")) : 8) : 8;
constexpr int kMaxTapsReg = 8;

struct AxisTable {
  int in_size = 0, out_size = 0, kmax = 0, tmax = 0;
  int span = 0;             // max source rows needed by a group of kFuseRows consecutive outputs (fused kernel strip height)
  int tspan = 0;            // max OUTPUT rows touched by a group of kFuseRows consecutive inputs (transpose)
  int* xmin = nullptr;      // [out]
  int* xsize = nullptr;     // [out]
  float* w = nullptr;       // [out, kmax]
  int* t_start = nullptr;   // [in + 1]  CSR over inputs
  int* t_out = nullptr;     // [nnz] output index
  float* t_w = nullptr;     // [nnz] weight
};

static int build_axis(int in_size, int out_size, AxisTable* t) {
  t->in_size = in_size;
  t->out_size = out_size;
  const float scale = (float)((double)in_size / (double)out_size);
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const int kmax = (int)ceilf(support) * 2 + 1;
  t->kmax = kmax;
  std::vector<int> xmin(out_size), xsize(out_size);
  std::vector<float> w((size_t)out_size * kmax, 0.f);
  std::vector<std::vector<std::pair<int, float>>> tr(in_size);
  for (int i = 0; i < out_size; ++i) {
    const float center = scale * ((float)i + 0.5f);
    int lo = (int)(center - support + 0.5f);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5f);
    if (hi > in_size) hi = in_size;
    xmin[i] = lo;
    xsize[i] = hi - lo;
    float tot = 0.f;
    for (int j = 0; j < hi - lo; ++j) {
      const float arg = ((float)(j + lo) - center + 0.5f) * invscale;
      const float v = fabsf(arg) < 1.0f ? 1.0f - fabsf(arg) : 0.f;
      w[(size_t)i * kmax + j] = v;
      tot += v;
    }
    for (int j = 0; j < hi - lo; ++j) {
      if (tot != 0.f) w[(size_t)i * kmax + j] /= tot;
      tr[lo + j].push_back({i, w[(size_t)i * kmax + j]});
    }
  }
  for (int o0 = 0; o0 < out_size; o0 += kFuseRows) {
    const int o1 = (o0 + kFuseRows < out_size ? o0 + kFuseRows : out_size) - 1;
    const int need = xmin[o1] + xsize[o1] - xmin[o0];
    if (need > t->span) t->span = need;
  }
  std::vector<int> ts(in_size + 1, 0), to;
  std::vector<float> tw;
  int tmax = 0;
  for (int s = 0; s < in_size; ++s) {
    ts[s] = (int)to.size();
    for (auto& e : tr[s]) { to.push_back(e.first); tw.push_back(e.second); }
    if ((int)tr[s].size() > tmax) tmax = (int)tr[s].size();
  }
  ts[in_size] = (int)to.size();
  t->tmax = tmax;
  for (int s0 = 0; s0 < in_size; s0 += kFuseRows) {
    const int s1 = s0 + kFuseRows < in_size ? s0 + kFuseRows : in_size;
    if (ts[s1] > ts[s0]) {
      const int need = to[ts[s1] - 1] - to[ts[s0]] + 1;
      if (need > t->tspan) t->tspan = need;
    }
  }
  if (to.empty()) { to.push_back(0); tw.push_back(0.f); }
  RGIE_CUDA_OK(cudaMalloc(&t->xmin, sizeof(int) * out_size));
  RGIE_CUDA_OK(cudaMalloc(&t->xsize, sizeof(int) * out_size));
  RGIE_CUDA_OK(cudaMalloc(&t->w, sizeof(float) * w.size()));
  RGIE_CUDA_OK(cudaMalloc(&t->t_start, sizeof(int) * ts.size()));
  RGIE_CUDA_OK(cudaMalloc(&t->t_out, sizeof(int) * to.size()));
  RGIE_CUDA_OK(cudaMalloc(&t->t_w, sizeof(float) * tw.size()));
  RGIE_CUDA_OK(cudaMemcpy(t->xmin, xmin.data(), sizeof(int) * out_size, cudaMemcpyHostToDevice));
  RGIE_CUDA_OK(cudaMemcpy(t->xsize, xsize.data(), sizeof(int) * out_size, cudaMemcpyHostToDevice));
  RGIE_CUDA_OK(cudaMemcpy(t->w, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
  RGIE_CUDA_OK(cudaMemcpy(t->t_start, ts.data(), sizeof(int) * ts.size(), cudaMemcpyHostToDevice));
  RGIE_CUDA_OK(cudaMemcpy(t->t_out, to.data(), sizeof(int) * to.size(), cudaMemcpyHostToDevice));
  RGIE_CUDA_OK(cudaMemcpy(t->t_w, tw.data(), sizeof(float) * tw.size(), cudaMemcpyHostToDevice));
  return 0;
}
static void free_axis(AxisTable* t) {
  cudaFree(t->xmin); cudaFree(t->xsize); cudaFree(t->w); cudaFree(t->t_start); cudaFree(t->t_out); cudaFree(t->t_w);
}

// out[plane, y, xo] = sum_j w[xo, j] * in[plane, y, xmin[xo] + j]        (AXIS 1: along x)
// out[plane, yo, x] = sum_j w[yo, j] * in[plane, ymin[yo] + j, x]        (AXIS 0: along y)
constexpr int kRowsPerBlock = 8;
// One thread block = kRowsPerBlock OUTPUT ROWS (row = plane * rows_out + r; no per-element division), threads along x so
// every load / store of a warp is contiguous (AXIS 0) or within a few cache lines (AXIS 1: neighbouring outputs share taps).
template <int AXIS>
__global__ void __launch_bounds__(256) resample_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                          const int* __restrict__ xmin, const int* __restrict__ xsize,
                                                          const float* __restrict__ w, int kmax, int in_size, int out_size,
                                                          int other, long rows) {
  // AXIS 1: in [planes, other, in_size] -> out [planes, other, out_size]   (row = plane * other + r)
  // AXIS 0: in [planes, in_size, other] -> out [planes, out_size, other]   (row = plane * out_size + o)
  for (long row = (long)blockIdx.x * kRowsPerBlock; row < min((long)(blockIdx.x + 1) * kRowsPerBlock, rows); ++row) {
    if (AXIS == 1) {
      const float* src = in + row * (long)in_size;
      float* dst = out + row * (long)out_size;
      for (int o = threadIdx.x; o < out_size; o += blockDim.x) {
        const int lo = xmin[o], n = xsize[o];
        const float* wk = w + (long)o * kmax;
        float s = 0.f;
        for (int j = 0; j < n; ++j) s = fmaf(wk[j], src[lo + j], s);
        dst[o] = s;
      }
    } else {
      const long plane = row / out_size;
      const int o = (int)(row - plane * out_size);
      const int lo = xmin[o], n = xsize[o];
      const float* wk = w + (long)o * kmax;
      const float* src = in + (plane * in_size + lo) * (long)other;
      float* dst = out + row * (long)other;
      for (int r = threadIdx.x; r < other; r += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < n; ++j) s = fmaf(wk[j], src[(long)j * other + r], s);
        dst[r] = s;
      }
    }
  }
}
// transpose: gin[plane, .., s] = sum_e t_w[e] * gout[plane, .., t_out[e]]    (one block = one row of gin)
template <int AXIS>
__global__ void __launch_bounds__(256) resample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin,
                                                          const int* __restrict__ t_start, const int* __restrict__ t_out,
                                                          const float* __restrict__ t_w, int in_size, int out_size,
                                                          int other, long rows) {
  // AXIS 1: gout [planes, other, out_size] -> gin [planes, other, in_size]   (row = plane * other + r)
  // AXIS 0: gout [planes, out_size, other] -> gin [planes, in_size, other]   (row = plane * in_size + s)
  for (long row = (long)blockIdx.x * kRowsPerBlock; row < min((long)(blockIdx.x + 1) * kRowsPerBlock, rows); ++row) {
    if (AXIS == 1) {
      const float* src = gout + row * (long)out_size;
      float* dst = gin + row * (long)in_size;
      for (int s_idx = threadIdx.x; s_idx < in_size; s_idx += blockDim.x) {
        float s = 0.f;
        const int e0 = t_start[s_idx], e1 = t_start[s_idx + 1];
        for (int e = e0; e < e1; ++e) s = fmaf(t_w[e], src[t_out[e]], s);
        dst[s_idx] = s;
      }
    } else {
      const long plane = row / in_size;
      const int s_idx = (int)(row - plane * in_size);
      const int e0 = t_start[s_idx], e1 = t_start[s_idx + 1];
      const float* src = gout + plane * (long)out_size * other;
      float* dst = gin + row * (long)other;
      for (int r = threadIdx.x; r < other; r += blockDim.x) {
        float s = 0.f;
        for (int e = e0; e < e1; ++e) s = fmaf(t_w[e], src[(long)t_out[e] * other + r], s);
        dst[r] = s;
      }
    }
  }
}

// ---- fused separable passes: the intermediate [rows, out_w] strip lives in shared memory, so a forward resize moves
// N_in + N_out bytes and nothing else.  Per-element arithmetic (tap order, fmaf chain, fp32 rounding of the intermediate)
// is exactly that of the two-pass kernels above, so results are bit-identical.
// Both kernels first stage the source rows they need in shared memory with coalesced 16-byte loads (every global element is
// read once per block, all loads independent), then run the two separable passes out of shared memory.
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int nrows, int width) {
  const long total = (long)nrows * width;                 // rows are contiguous in global memory
  if ((width & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (long i = threadIdx.x; i < total / 4; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (long i = threadIdx.x; i < total; i += blockDim.x) dst[i] = src[i];
  }
}

// forward: one block = kFuseRows output rows of one plane; needs input rows [ymin[o0], ymin[o1] + ysize[o1])
__global__ void __launch_bounds__(512) resize_fused_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              const int* __restrict__ xmin, const int* __restrict__ xsize,
                                                              const float* __restrict__ xw, int xk,
                                                              const int* __restrict__ ymin, const int* __restrict__ ysize,
                                                              const float* __restrict__ yw, int yk, int in_h, int in_w,
                                                              int out_h, int out_w, int blocks_per_plane, int max_rows, int kFuseRows) {
  extern __shared__ __align__(16) float smem_f[];
  float* rows = smem_f;                                  // [nrows, in_w]   source rows
  float* strip = smem_f + (long)max_rows * in_w;         // [nrows, out_w]  after the horizontal pass
  const int plane = blockIdx.x / blocks_per_plane;
  const int o0 = (blockIdx.x - plane * blocks_per_plane) * kFuseRows;
  const int o1 = min(o0 + kFuseRows, out_h) - 1;
  const int y_lo = ymin[o0], nrows = ymin[o1] + ysize[o1] - y_lo;
  stage_rows(rows, in + ((long)plane * in_h + y_lo) * in_w, nrows, in_w);
  __syncthreads();
  for (int xo = threadIdx.x; xo < out_w; xo += blockDim.x) {
    const int lo = xmin[xo], n = xsize[xo];
    float wk[kMaxTapsReg];
#pragma unroll
    for (int j = 0; j < kMaxTapsReg; ++j) wk[j] = j < n ? xw[(long)xo * xk + j] : 0.f;
    for (int r = 0; r < nrows; ++r) {
      const float* row = rows + r * in_w + lo;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxTapsReg; ++j)
        if (j < n) acc = fmaf(wk[j], row[j], acc);
      strip[r * out_w + xo] = acc;
    }
  }
  __syncthreads();
  float* dst = out + ((long)plane * out_h + o0) * out_w;
  for (int o = o0; o <= o1; ++o) {
    const int lo = ymin[o] - y_lo, n = ysize[o];
    const float* wk = yw + (long)o * yk;
    for (int x = threadIdx.x; x < out_w; x += blockDim.x) {
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(wk[j], strip[(lo + j) * out_w + x], acc);
      dst[(long)(o - o0) * out_w + x] = acc;
    }
  }
}

// backward: one block = kFuseRows rows of gin of one plane; needs gout rows [yo[ys[s0]], yo[ys[s0 + ns] - 1]]
__global__ void __launch_bounds__(512) resize_fused_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin,
                                                              const int* __restrict__ ys, const int* __restrict__ yo,
                                                              const float* __restrict__ yw, const int* __restrict__ xs,
                                                              const int* __restrict__ xo_tab, const float* __restrict__ xw,
                                                              int in_h, int in_w, int out_h, int out_w, int blocks_per_plane,
                                                              int max_rows, int kFuseRows) {
  extern __shared__ __align__(16) float smem_f[];
  float* rows = smem_f;                                  // [nrows, out_w]      gout rows
  float* strip = smem_f + (long)max_rows * out_w;        // [kFuseRows, out_w]  after the vertical transpose
  const int plane = blockIdx.x / blocks_per_plane;
  const int s0 = (blockIdx.x - plane * blocks_per_plane) * kFuseRows;
  const int ns = min(s0 + kFuseRows, in_h) - s0;
  const int e_lo = ys[s0], e_hi = ys[s0 + ns];
  const int r_lo = e_hi > e_lo ? yo[e_lo] : 0;
  const int nrows = e_hi > e_lo ? yo[e_hi - 1] - r_lo + 1 : 0;
  stage_rows(rows, gout + ((long)plane * out_h + r_lo) * out_w, nrows, out_w);
  __syncthreads();
  for (int r = 0; r < ns; ++r) {
    const int e0 = ys[s0 + r], e1 = ys[s0 + r + 1];
    for (int x = threadIdx.x; x < out_w; x += blockDim.x) {
      float acc = 0.f;
      for (int e = e0; e < e1; ++e) acc = fmaf(yw[e], rows[(yo[e] - r_lo) * out_w + x], acc);
      strip[r * out_w + x] = acc;
    }
  }
  __syncthreads();
  float* dst = gin + ((long)plane * in_h + s0) * in_w;
  for (int x = threadIdx.x; x < in_w; x += blockDim.x) {
    const int e0 = xs[x], e1 = xs[x + 1];
    for (int r = 0; r < ns; ++r) {
      float acc = 0.f;
      for (int e = e0; e < e1; ++e) acc = fmaf(xw[e], strip[r * out_w + xo_tab[e]], acc);
      dst[(long)r * in_w + x] = acc;
    }
  }
}

}  // namespace rgie

using namespace rgie;

struct RgieResize {
  AxisTable ax_w, ax_h;
  int in_h, in_w, out_h, out_w;
  int fwd_rows = 0, bwd_rows = 0;   // source rows a block of the fused forward / backward kernel stages at most
  bool fused = false;
};

extern "C" {

int rgie_resize_create(int in_h, int in_w, int out_h, int out_w, RgieResize** out) {
  RGIE_CHECK(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && out != nullptr, "rgie_resize_create: bad shape");
  RgieResize* r = new RgieResize();
  r->in_h = in_h; r->in_w = in_w; r->out_h = out_h; r->out_w = out_w;
  if (int rc = build_axis(in_w, out_w, &r->ax_w)) { delete r; return rc; }
  if (int rc = build_axis(in_h, out_h, &r->ax_h)) { delete r; return rc; }
  {
    // fused single-kernel path: the shared-memory strip must fit and the taps of one output column fit the register array
    r->fwd_rows = r->ax_h.span;
    r->bwd_rows = r->ax_h.tspan;
    const size_t smem_f = (size_t)r->fwd_rows * (in_w + out_w) * sizeof(float);
    const size_t smem_b = (size_t)(r->bwd_rows + kFuseRows) * out_w * sizeof(float);
    static const bool env_off = getenv("RGIE_RESIZE_FUSED") && atoi(getenv("RGIE_RESIZE_FUSED")) == 0;
    r->fused = !env_off && r->ax_w.kmax <= kMaxTapsReg && smem_f <= 100 * 1024 && smem_b <= 100 * 1024;
    if (r->fused) {
      RGIE_CUDA_OK(cudaFuncSetAttribute(resize_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      RGIE_CUDA_OK(cudaFuncSetAttribute(resize_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
  }
  *out = r;
  return 0;
}

void rgie_resize_destroy(RgieResize* r) {
  if (!r) return;
  free_axis(&r->ax_w);
  free_axis(&r->ax_h);
  delete r;
}

int rgie_resize_fwd(const RgieResize* r, const float* in, float* out, int planes, float* tmp, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RGIE_CHECK(r && in && out && tmp && planes > 0, "rgie_resize_fwd: bad arguments");
  if (r->fused) {
    const int bpp = (r->out_h + kFuseRows - 1) / kFuseRows;
    resize_fused_fwd_kernel<<<planes * bpp, kFuseThreads, (size_t)r->fwd_rows * (r->in_w + r->out_w) * sizeof(float), st>>>(
        in, out, r->ax_w.xmin, r->ax_w.xsize, r->ax_w.w, r->ax_w.kmax, r->ax_h.xmin, r->ax_h.xsize, r->ax_h.w, r->ax_h.kmax,
        r->in_h, r->in_w, r->out_h, r->out_w, bpp, r->fwd_rows, kFuseRows);
    RGIE_LAUNCH_OK();
    return 0;
  }
  const long rows_in = (long)planes * r->in_h, rows_out = (long)planes * r->out_h;
  const int g_in = (int)((rows_in + kRowsPerBlock - 1) / kRowsPerBlock), g_out = (int)((rows_out + kRowsPerBlock - 1) / kRowsPerBlock);
  resample_fwd_kernel<1><<<g_in, 256, 0, st>>>(in, tmp, r->ax_w.xmin, r->ax_w.xsize, r->ax_w.w, r->ax_w.kmax, r->in_w, r->out_w,
                                               r->in_h, rows_in);
  RGIE_LAUNCH_OK();
  resample_fwd_kernel<0><<<g_out, 256, 0, st>>>(tmp, out, r->ax_h.xmin, r->ax_h.xsize, r->ax_h.w, r->ax_h.kmax, r->in_h, r->out_h,
                                                r->out_w, rows_out);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_resize_bwd(const RgieResize* r, const float* gout, float* gin, int planes, float* tmp, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RGIE_CHECK(r && gout && gin && tmp && planes > 0, "rgie_resize_bwd: bad arguments");
  // transpose of (vertical o horizontal) = horizontal^T o vertical^T ; tmp: [planes, in_h, out_w]
  if (r->fused) {
    const int bpp = (r->in_h + kFuseRows - 1) / kFuseRows;
    resize_fused_bwd_kernel<<<planes * bpp, kFuseThreads, (size_t)(r->bwd_rows + kFuseRows) * r->out_w * sizeof(float), st>>>(
        gout, gin, r->ax_h.t_start, r->ax_h.t_out, r->ax_h.t_w, r->ax_w.t_start, r->ax_w.t_out, r->ax_w.t_w, r->in_h, r->in_w,
        r->out_h, r->out_w, bpp, r->bwd_rows, kFuseRows);
    RGIE_LAUNCH_OK();
    return 0;
  }
  const long rows_in = (long)planes * r->in_h;
  const int g_in = (int)((rows_in + kRowsPerBlock - 1) / kRowsPerBlock);
  resample_bwd_kernel<0><<<g_in, 256, 0, st>>>(gout, tmp, r->ax_h.t_start, r->ax_h.t_out, r->ax_h.t_w, r->in_h, r->out_h, r->out_w,
                                               rows_in);
  RGIE_LAUNCH_OK();
  resample_bwd_kernel<1><<<g_in, 256, 0, st>>>(tmp, gin, r->ax_w.t_start, r->ax_w.t_out, r->ax_w.t_w, r->in_w, r->out_w, r->in_h,
                                               rows_in);
  RGIE_LAUNCH_OK();
  return 0;
}

}  // extern "C"

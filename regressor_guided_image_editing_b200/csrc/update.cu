// Small per-problem kernels around the regressor: parameter-vector transform (reference clamps), loss head,
// fused Adam + on-device best-x tracking, and the regressor-guidance (normalised SGD) update.
// These are launch-latency-bound by nature (B x 41 floats); they exist to keep the whole step on the device with no
// host synchronisation (the reference syncs twice per step: baselines/optimize_image.py:78,90).
#include "common.cuh"
#include "rgie.h"

namespace rgie {
namespace {

// default filter list layout (optimize_image_param.py:227): exposure 0 | saturation 1 | tone 2..9 | color 10..33 |
// contrast 34 | sharp 35 | blur 36 | scale 37..40 (sx, sy, cx, cy)
constexpr int kNP = 41;

__global__ void params_default_fwd_kernel(const float* __restrict__ x, float* __restrict__ p, int B, float input_size) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kNP) return;
  const int k = i % kNP;
  float v = x[i];
  if (k == 1 || k == 35 || k == 36) v = fmaxf(v, 0.f);            // image_transformations.py:98,195,120
  else if (k == 37 || k == 38) v = fmaxf(v, 1.0f);                 // optimize_image_param.py:279
  else if (k == 39 || k == 40) v = fminf(fmaxf(v, 0.f), input_size);   // :280
  else if (k == 34) v = v < 0.f ? 0.f : v;                         // :291
  p[i] = v;
}
__global__ void params_default_bwd_kernel(const float* __restrict__ x, float* __restrict__ gp, int B, float input_size) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kNP) return;
  const int k = i % kNP;
  const float v = x[i];
  bool pass = true;
  if (k == 1 || k == 35 || k == 36) pass = v >= 0.f;
  else if (k == 37 || k == 38) pass = v >= 1.0f;
  else if (k == 39 || k == 40) pass = v >= 0.f && v <= input_size;
  else if (k == 34) pass = !(v < 0.f);
  if (!pass) gp[i] = 0.f;
}

__global__ void va_head_kernel(const float* __restrict__ logits, int B, int reps, int nc, int sigmoid,
                               const float* __restrict__ target, float tv_def, float ta_def, int use_mask, float scale,
                               float* __restrict__ preds, float* __restrict__ loss, float* __restrict__ dlogits) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float l = 0.f;
  for (int k = 0; k < nc; ++k) {
    float s = 0.f;
    for (int r = 0; r < reps; ++r) s += logits[((long)b * reps + r) * nc + k];
    const float z = s / (float)reps;
    const float pr = sigmoid ? 1.0f / (1.0f + expf(-z)) : z;
    preds[(long)b * nc + k] = pr;
    float dz = 0.f;
    if (k < 2 && ((use_mask >> k) & 1)) {
      const float t = target ? target[(long)b * 2 + k] : (k == 0 ? tv_def : ta_def);
      const float e = t - pr;
      l += e * e;
      const float dpr = -2.0f * e * scale;
      dz = sigmoid ? dpr * pr * (1.0f - pr) : dpr;
    }
    if (dlogits)
      for (int r = 0; r < reps; ++r) dlogits[((long)b * reps + r) * nc + k] = dz / (float)reps;
  }
  if (loss) loss[b] = scale * l;
}

__global__ void adam_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int B, int n, float step_size, float bc2_sqrt, float w1, float beta2,
                            float w2, float eps, const float* __restrict__ loss, float* __restrict__ best_loss,
                            float* __restrict__ best_x, int* __restrict__ best_step, int step) {
  const int b = blockIdx.x;
  bool better = false;
  if (loss && best_loss) better = loss[b] < best_loss[b];      // strict '<' (optimize_image.py:78); NaN never wins
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long idx = (long)b * n + i;
    const float xi = x[idx];
    if (better && best_x) best_x[idx] = xi;                    // snapshot BEFORE the update (:78-81)
    const float gi = g[idx];
    // torch _single_tensor_adam: exp_avg.lerp_(grad, 1-beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    float mi = m[idx];
    mi = __fadd_rn(mi, __fmul_rn(w1, gi - mi));
    float vi = __fadd_rn(__fmul_rn(v[idx], beta2), __fmul_rn(__fmul_rn(w2, gi), gi));
    m[idx] = mi;
    v[idx] = vi;
    const float denom = __fadd_rn(sqrtf(vi) / bc2_sqrt, eps);
    x[idx] = __fadd_rn(xi, __fmul_rn(-step_size, mi) / denom);
  }
  __syncthreads();
  if (better && threadIdx.x == 0) {
    best_loss[b] = loss[b];
    if (best_step) best_step[b] = step;
  }
}

// graph-replayable variant: step-dependent scalars come from a device table indexed by a device step counter
__global__ void adam_sched_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, int B, int n, const float* __restrict__ sched,
                                  const int* __restrict__ step_ptr, float w1, float beta2, float w2, float eps,
                                  const float* __restrict__ loss, float* __restrict__ best_loss,
                                  float* __restrict__ best_x, int* __restrict__ best_step) {
  const int step = *step_ptr;
  const float step_size = sched[2 * step], bc2_sqrt = sched[2 * step + 1];
  const int b = blockIdx.x;
  bool better = false;
  if (loss && best_loss) better = loss[b] < best_loss[b];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long idx = (long)b * n + i;
    const float xi = x[idx];
    if (better && best_x) best_x[idx] = xi;
    const float gi = g[idx];
    float mi = m[idx];
    mi = __fadd_rn(mi, __fmul_rn(w1, gi - mi));
    float vi = __fadd_rn(__fmul_rn(v[idx], beta2), __fmul_rn(__fmul_rn(w2, gi), gi));
    m[idx] = mi;
    v[idx] = vi;
    const float denom = __fadd_rn(sqrtf(vi) / bc2_sqrt, eps);
    x[idx] = __fadd_rn(xi, __fmul_rn(-step_size, mi) / denom);
  }
  __syncthreads();
  if (better && threadIdx.x == 0) {
    best_loss[b] = loss[b];
    if (best_step) best_step[b] = step;
  }
}
__global__ void record_kernel(const float* __restrict__ src, float* __restrict__ table, const int* __restrict__ step_ptr, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) table[(long)(*step_ptr) * n + i] = src[i];
}
__global__ void counter_add_kernel(int* counter, int delta) { *counter += delta; }

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long per, float* __restrict__ part) {
  __shared__ float red[8];
  const int b = blockIdx.y;
  float s = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    const float v = g[(long)b * per + i];
    s = fmaf(v, v, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    part[(long)b * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) guidance_axpy_kernel(float* __restrict__ x, const float* __restrict__ g, long per,
                                                           const float* __restrict__ part, int nparts, float scale,
                                                           int normalize) {
  const int b = blockIdx.y;
  float k = scale;
  if (normalize) {
    float t = 0.f;
    for (int w = 0; w < nparts; ++w) t += part[(long)b * nparts + w];
    k = scale / (sqrtf(t) + 1e-10f);
  }
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    const long idx = (long)b * per + i;
    x[idx] = x[idx] - k * g[idx];
  }
}

}  // namespace
}  // namespace rgie

using namespace rgie;

extern "C" {

int rgie_params_default_fwd(const float* x, float* p, int B, float input_size, void* stream) {
  RGIE_CHECK(x && p && B > 0, "rgie_params_default_fwd: bad arguments");
  params_default_fwd_kernel<<<ceil_div((long)B * kNP, 128), 128, 0, (cudaStream_t)stream>>>(x, p, B, input_size);
  RGIE_LAUNCH_OK();
  return 0;
}
int rgie_params_default_bwd(const float* x, float* gp, int B, float input_size, void* stream) {
  RGIE_CHECK(x && gp && B > 0, "rgie_params_default_bwd: bad arguments");
  params_default_bwd_kernel<<<ceil_div((long)B * kNP, 128), 128, 0, (cudaStream_t)stream>>>(x, gp, B, input_size);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_va_head(const float* logits, int B, int reps, int num_classes, int sigmoid, const float* target,
                 float tv_default, float ta_default, int use_mask, float scale, float* preds, float* loss,
                 float* dlogits, void* stream) {
  RGIE_CHECK(logits && preds && B > 0 && reps > 0 && num_classes > 0, "rgie_va_head: bad arguments");
  va_head_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(logits, B, reps, num_classes, sigmoid, target,
                                                                   tv_default, ta_default, use_mask, scale, preds, loss,
                                                                   dlogits);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_adam_step(float* x, const float* g, float* m, float* v, int B, int n, float step_size, float bc2_sqrt,
                   float one_minus_beta1, float beta2, float one_minus_beta2, float eps, const float* loss, float* best_loss, float* best_x,
                   int* best_step, int step, void* stream) {
  RGIE_CHECK(x && g && m && v && B > 0 && n > 0, "rgie_adam_step: bad arguments");
  adam_kernel<<<B, 64, 0, (cudaStream_t)stream>>>(x, g, m, v, B, n, step_size, bc2_sqrt, one_minus_beta1, beta2,
                                                  one_minus_beta2, eps, loss,
                                                  best_loss, best_x, best_step, step);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_adam_step_sched(float* x, const float* g, float* m, float* v, int B, int n, const float* sched,
                        const int* step_ptr, float one_minus_beta1, float beta2, float one_minus_beta2, float eps,
                        const float* loss, float* best_loss, float* best_x, int* best_step, void* stream) {
  RGIE_CHECK(x && g && m && v && sched && step_ptr && B > 0 && n > 0, "rgie_adam_step_sched: bad arguments");
  adam_sched_kernel<<<B, 64, 0, (cudaStream_t)stream>>>(x, g, m, v, B, n, sched, step_ptr, one_minus_beta1, beta2,
                                                        one_minus_beta2, eps, loss, best_loss, best_x, best_step);
  RGIE_LAUNCH_OK();
  return 0;
}
int rgie_record(const float* src, float* table, const int* step_ptr, int n, void* stream) {
  RGIE_CHECK(src && table && step_ptr && n > 0, "rgie_record: bad arguments");
  record_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(src, table, step_ptr, n);
  RGIE_LAUNCH_OK();
  return 0;
}
int rgie_counter_add(int* counter, int delta, void* stream) {
  RGIE_CHECK(counter != nullptr, "rgie_counter_add: null");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  RGIE_LAUNCH_OK();
  return 0;
}

int rgie_guidance_update(float* x, const float* g, int n_problems, long per_problem, float scale, int normalize,
                         float* ws, void* stream) {
  RGIE_CHECK(x && g && ws && n_problems > 0 && per_problem > 0, "rgie_guidance_update: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int nparts = ceil_div(per_problem, 256 * 8);
  if (nparts > 128) nparts = 128;
  if (nparts < 1) nparts = 1;
  dim3 grid(nparts, n_problems);
  if (normalize) {
    sumsq_partial_kernel<<<grid, 256, 0, st>>>(g, per_problem, ws);
    RGIE_LAUNCH_OK();
  }
  guidance_axpy_kernel<<<grid, 256, 0, st>>>(x, g, per_problem, ws, nparts, scale, normalize);
  RGIE_LAUNCH_OK();
  return 0;
}

}  // extern "C"

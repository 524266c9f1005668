// CUDA-core backend of the row-shifted GEMM (common.cuh: GemmDesc).
//
// Role: (1) the fp32 "parity mode" of the regressor (fp32 activations/weights, fp32 accumulate) that is compared
// against the CPU oracle at <=1e-3 pixel tolerance, and (2) the on-device cross-check of the tcgen05 backend
// (same descriptor, bf16 operands).  It is not the throughput path.
#include "common.cuh"

namespace rgie {

template <typename T>
__device__ __forceinline__ void epilogue_store(const GemmDesc& d, long m, long dest, int n0, const float* acc, int cnt) {
  // acc[0..cnt) are columns n0..n0+cnt of row m
  const T* res = reinterpret_cast<const T*>(d.res);
  const T* mask = reinterpret_cast<const T*>(d.mask);
  for (int c = 0; c < cnt; ++c) {
    int n = n0 + c;
    if (n >= d.Cout) break;
    float v = acc[c];
    if (d.bias) v += d.bias[n];
    if (res && m < d.res_rows) v += to_f<T>(res[m * d.ld_res + n]);
    if (d.relu) v = fmaxf(v, 0.f);
    if (mask) v = (to_f<T>(mask[m * d.ld_mask + n]) > 0.f) ? v : 0.f;
    if (d.mask_bits) v = ((d.mask_bits[bits_index(m, n >> 5, d.ld_mb)] >> (n & 31)) & 1u) ? v : 0.f;
    float stored = v;
    if (d.d_fp32) reinterpret_cast<float*>(d.D)[dest * d.ldd + n] = v;
    else { const T o = from_f<T>(v); reinterpret_cast<T*>(d.D)[dest * d.ldd + n] = o; stored = to_f<T>(o); }
    if (d.D_bits && stored > 0.f) atomicOr(d.D_bits + bits_index(dest, n >> 5, d.ld_db), 1u << (n & 31));   // words zeroed by the launcher
  }
}

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDesc d) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const T* A = reinterpret_cast<const T*>(d.A);
  const T* W = reinterpret_cast<const T*>(d.Wt);
  const int tid = threadIdx.x;
  const long m0 = d.m_begin + (long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int Ktot = d.ntaps * d.Cin + (d.A2 ? d.Cin2 : 0);
  const int lr = tid >> 2;         // 0..63 : row inside the tile (A) / output channel (W)
  const int lk = (tid & 3) * 4;    // 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;   // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nsrc = d.ntaps + (d.A2 ? 1 : 0);      // the optional second operand is one more "tap" with its own matrix
  for (int t = 0; t < nsrc; ++t) {
    const bool second = t >= d.ntaps;
    const T* At = second ? reinterpret_cast<const T*>(d.A2) : A;
    const int cin = second ? d.Cin2 : d.Cin;
    const long arow = m0 + lr + (second ? 0 : d.row_off[t]);
    const bool a_ok = arow >= 0 && arow < (second ? d.a2_rows : d.a_rows) && (m0 + lr) < d.m_end;
    const int wn = n0 + lr;
    const bool w_ok = wn < d.n_pad;
    for (int c0 = 0; c0 < cin; c0 += TK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
      if (a_ok) {
        const T* p = At + arow * (second || d.a_ld == 0 ? cin : d.a_ld) + c0 + lk;
#pragma unroll
        for (int q = 0; q < 4; ++q) av[q] = to_f<T>(p[q]);
      }
      if (w_ok) {
        const T* p = W + (long)wn * Ktot + (long)t * d.Cin + c0 + lk;
#pragma unroll
        for (int q = 0; q < 4; ++q) wv[q] = to_f<T>(p[q]);
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        As[lk + q][lr] = av[q];
        Bs[lk + q][lr] = wv[q];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long m = m0 + ty * 4 + i;
    if (m >= d.m_end) continue;
    long dest = map_row(d.src, d.dst_kind, d.dst, m);
    if (dest < 0) continue;
    epilogue_store<T>(d, m, dest, n0 + tx * 4, acc[i], 4);
  }
}

int launch_gemm_simt(const GemmDesc& d, int dtype, cudaStream_t st) {
  RGIE_CHECK(d.Cin % TK == 0, "gemm_simt: Cin must be a multiple of 16");
  RGIE_CHECK(d.ntaps >= 1 && d.ntaps <= kMaxTaps, "gemm_simt: ntaps out of range");
  RGIE_CHECK(d.A2 == nullptr || d.Cin2 % TK == 0, "gemm_simt: Cin2 must be a multiple of 16");
  RGIE_CHECK((d.mask_bits == nullptr && d.D_bits == nullptr) || d.Cout % 32 == 0, "gemm_simt: bit masks need Cout % 32 == 0");
  long M = d.m_end - d.m_begin;
  if (M <= 0) return 0;
  if (d.D_bits) {
    const long rows = d.dst_kind == DST_TO_PLAIN ? (long)d.src.n_img * d.src.H * d.src.W : d.dst.rows();
    RGIE_CUDA_OK(cudaMemsetAsync(d.D_bits, 0, (size_t)bits_words(rows, d.ld_db) * 4, st));
  }
  dim3 grid(ceil_div(M, TM), ceil_div(d.Cout, TN));
  if (dtype == 0) gemm_simt_kernel<float><<<grid, 256, 0, st>>>(d);
  else gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(d);
  RGIE_LAUNCH_OK();
  return 0;
}

}  // namespace rgie

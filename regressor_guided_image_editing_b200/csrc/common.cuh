// Shared definitions for the rgie sm_100a extension (C-ABI in include/rgie.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace rgie {

// ---------------------------------------------------------------------------------------------------------------
// error plumbing: C-ABI functions return an int status and stash a thread-local message (never throw across the ABI)
// ---------------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int  fail(const std::string& msg);          // sets the message, returns RGIE_ERR (1)

#define RGIE_CUDA_OK(expr)                                                                           \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::rgie::fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                          std::to_string(__LINE__));                                                 \
  } while (0)

#define RGIE_CHECK(cond, msg)                                       \
  do {                                                              \
    if (!(cond)) return ::rgie::fail(std::string("check failed: ") + (msg)); \
  } while (0)

void count_launch();
#define RGIE_LAUNCH_OK()                  \
  do {                                    \
    ::rgie::count_launch();               \
    RGIE_CUDA_OK(cudaGetLastError());     \
  } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// Per-device one-time setup at a call site (cudaFuncSetAttribute is a per-device property): a bit per device ordinal,
// updated atomically so that two host threads driving two GPUs cannot race.
struct DeviceOnce {
  unsigned long long done_bits = 0ull;
  bool needed() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    return ((__atomic_load_n(&done_bits, __ATOMIC_ACQUIRE) >> dev) & 1ull) == 0ull;
  }
  void done() {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) __atomic_fetch_or(&done_bits, 1ull << dev, __ATOMIC_RELEASE);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Pixel-row geometry of an activation matrix [rows, channels] (channels contiguous, NHWC-like).
// A tensor is `planes` stacked planes; each plane holds n_img images of (H + pad_t + pad_b) x P pixel rows,
// the valid HxW window sitting at (pad_t, pad_l).  Pad pixels are kept at zero (buffers are zeroed once at
// creation and the epilogues never write pad rows), which is what turns a 3x3 convolution into nine row-shifted
// GEMM operand loads.  planes == 4 is the 2x2 phase split (space-to-depth by parity) used around stride-2 convs.
// ---------------------------------------------------------------------------------------------------------------
// Division of a non-negative 31-bit integer by a runtime constant: q = umulhi(n, mul) >> shift with
// mul = ceil(2^(31 + ceil(log2 d)) / d) < 2^32 (exact for every n < 2^31); d == 1 is flagged by mul == 0.
// The epilogue decodes one pixel row per thread per tile; 64-bit hardware-emulated divisions there cost more
// instructions than the rest of the epilogue.
struct FastDiv {
  uint32_t d, mul, shift;
  __host__ __device__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return mul ? (__umulhi(n, mul) >> shift) : n;
#else
    return mul ? (uint32_t)(((uint64_t)n * mul) >> 32) >> shift : n;
#endif
  }
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d; f.mul = 0; f.shift = 0;
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;                    // ceil(log2 d) >= 1
  const unsigned __int128 one = 1;
  f.mul = (uint32_t)(((one << (31 + l)) + d - 1) / d);
  f.shift = l - 1;
  return f;
}

struct Geom {
  int planes;   // 1 or 4
  int n_img;
  int H, W;     // valid extent of one plane
  int pad_t, pad_l;
  int P;        // pitch in pixels
  int S;        // pixel rows per image per plane
  FastDiv fd_plane, fd_S, fd_P;   // dividers by plane_rows(), S, P (plane_rows() < 2^31)
  __host__ __device__ long plane_rows() const { return (long)n_img * S; }
  __host__ __device__ long rows() const { return (long)planes * n_img * S; }
};

static inline Geom make_geom(int planes, int n_img, int H, int W, int pad_t, int pad_b, int pad_l, int pad_r) {
  Geom g;
  g.planes = planes; g.n_img = n_img; g.H = H; g.W = W; g.pad_t = pad_t; g.pad_l = pad_l;
  g.P = W + pad_l + pad_r;
  g.S = g.P * (H + pad_t + pad_b);
  g.fd_plane = make_fastdiv((uint32_t)g.plane_rows());
  g.fd_S = make_fastdiv((uint32_t)g.S);
  g.fd_P = make_fastdiv((uint32_t)g.P);
  return g;
}

enum DstKind : int {
  DST_SAME = 0,        // dest row = m (valid rows only)
  DST_TO_PHASE = 1,    // src: 1 plane (H,W)  -> dst: 4 planes (H/2,W/2), plane = (i&1)*2 + (j&1)
  DST_FROM_PHASE = 2,  // src: 4 planes (H,W) -> dst: 1 plane (2H,2W)
  DST_TO_PLAIN = 3     // dst: dense [n_img, H, W] rows, no padding
};

// decode row m of geometry g; returns false for pad rows
__host__ __device__ inline bool geom_decode(const Geom& g, long m, int& plane, int& n, int& i, int& j) {
  // all row indices are < 2^31 (checked where the tensors are created)
  const uint32_t mu = (uint32_t)m;
  plane = g.planes == 1 ? 0 : (int)g.fd_plane.div(mu);
  const uint32_t mm = mu - (uint32_t)plane * (uint32_t)g.plane_rows();
  n = (int)g.fd_S.div(mm);
  int rem = (int)(mm - (uint32_t)n * (uint32_t)g.S);
  int ri = (int)g.fd_P.div((uint32_t)rem);
  i = ri - g.pad_t;
  j = rem - ri * g.P - g.pad_l;
  return plane < g.planes && n < g.n_img && i >= 0 && i < g.H && j >= 0 && j < g.W;
}

__host__ __device__ inline long geom_row(const Geom& g, int plane, int n, int i, int j) {
  return (long)plane * g.plane_rows() + (long)n * g.S + (long)(i + g.pad_t) * g.P + (j + g.pad_l);
}

// destination row for source row m, or -1 when m is a pad row (skipped)
__host__ __device__ inline long map_row(const Geom& src, int kind, const Geom& dst, long m) {
  int plane, n, i, j;
  if (!geom_decode(src, m, plane, n, i, j)) return -1;
  switch (kind) {
    case DST_SAME: return m;
    case DST_TO_PHASE: return geom_row(dst, (i & 1) * 2 + (j & 1), n, i >> 1, j >> 1);
    case DST_FROM_PHASE: return geom_row(dst, 0, n, 2 * i + (plane >> 1), 2 * j + (plane & 1));
    default: return ((long)n * src.H + i) * src.W + j;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The universal "row-shifted GEMM":
//   D[map(m), n] = epi( sum_t sum_c A[m + row_off[t], c] * Wt[n, t*Cin + c]  +  sum_c A2[m, c] * Wt[n, ntaps*Cin + c] )
// The optional second operand A2 (rows >= a2_rows contribute nothing) concatenates a second contraction along K: it
// fuses the 1x1 downsample branch of a bottleneck into its conv3 (forward) / conv1 input gradient (backward), so the
// branch output is never written to or re-read from HBM.
// Both backends (tcgen05 in gemm_sm100.cu, CUDA-core fp32/bf16 in gemm_simt.cu) consume this descriptor.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 16;

struct GemmDesc {
  // operands
  const void* A;      long a_rows;  int Cin;     // A: [a_rows, Cin] row-major
  int a_ld;                                       // elements between consecutive rows of A; 0 = Cin (dense).  a_ld < Cin: rows
                                                  // OVERLAP (row m = the Cin elements starting at element m * a_ld) -- conv1 reads
                                                  // its 16-channel packed pixels as 4-pixel windows without replicating them
  const void* Wt;     int  n_pad;                 // Wt: [n_pad, ntaps*Cin] row-major (K-major), n_pad >= Cout
  int ntaps;          long row_off[kMaxTaps];
  const void* A2;     long a2_rows; int Cin2;    // optional second operand [a2_rows, Cin2] (row offset 0), or null
  long m_begin, m_end;                            // rows enumerated (tiles start at m_begin)
  int Cout;
  // epilogue
  Geom src; int dst_kind; Geom dst;
  void* D;            int ldd;       int d_fp32;  // output (T or float), leading dim in elements
  const float* bias;                               // [Cout] or null
  const void* res;    int ld_res;    long res_rows; // + res[m, n] for m < res_rows (same dtype as A)
  const void* mask;   int ld_mask;                 // * (mask[m, n] > 0)          (same dtype as A)
  const uint32_t* mask_bits; int ld_mb;            // * bit n of row m (1 bit per element, ld in 32-bit words); alternative to mask
  uint32_t* D_bits;   int ld_db;                   // also emit (stored value > 0) as 1 bit per element at row map(m) (tcgen05 backend)
  int relu;
};

// 1-bit-per-element masks (sign bits of an activation): word w (32 channels) of pixel row m.  Blocked by 32 rows so that
// the 32 lanes of a warp (32 consecutive rows, same word) touch one contiguous 128-byte segment.
// An array for `rows` rows of `words` words holds bits_words(rows, words) 32-bit words.
__host__ __device__ inline long bits_index(long m, int w, int words) { return ((m >> 5) * words + w) * 32 + (m & 31); }
__host__ __device__ inline long bits_words(long rows, int words) { return ((rows + 31) / 32) * 32 * (long)words; }

// element helpers
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// launchers (defined in the respective .cu files); dtype: 0 = fp32 activations, 1 = bf16 activations
int launch_gemm_simt(const GemmDesc& d, int dtype, cudaStream_t st);
int launch_gemm_sm100(const GemmDesc& d, cudaStream_t st);   // bf16 only, tcgen05/TMEM/TMA

}  // namespace rgie

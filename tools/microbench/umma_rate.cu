// Microbenchmark: issue rate of tcgen05.mma cta_group::1 kind::f16 (bf16, M=128, K=16) from shared-memory operands
// (SWIZZLE_128B K-major tiles, contents irrelevant) for several N, one CTA per SM, no loads.  Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t tm, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tm), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: 4 K-slices of one 128x64 A tile per "k-block" (descriptor +32 B), accumulate into one accumulator
// mode 1: same K-slice every time; mode 2: alternate between two accumulators
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t sa = smem_u32(smem), sb = sa + 4 * 16384;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t ad = make_desc(sa + (i & 3) * 16384), bd = make_desc(sb + (i & 3) * 32768);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int ks = mode == 1 ? 0 : kk;
        const uint32_t d = tm + (mode == 2 ? ((kk & 1) * 256) : 0);
        mma(d, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 4 * 16384 + 4 * 32768 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {16, 32, 64, 128, 256}) {
      if (mode == 2 && N > 256) continue;
      for (int grid : {1, 148}) {
        k<<<grid, 128, smem>>>(N, iters, mode, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("mode %d N %3d grid %3d : %7.1f cycles/MMA  (%s)\n", mode, N, grid, (double)c / (iters * 4.0), cudaGetErrorString(e));
      }
    }
  return 0;
}

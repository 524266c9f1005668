// Microbenchmark: issue rate of tcgen05.mma cta_group::2 kind::f16 (bf16, M=256 across a CTA pair, K=16) from shared-memory
// operands for several N, clusters of 2 CTAs, no loads.  Prints cycles per MMA as seen by the issuing (leader) thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_2cta umma_rate_2cta.cu && ./umma_rate_2cta
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
__device__ __forceinline__ void mma2(uint32_t tm, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tm), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int N, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bar;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t sa = smem_u32(smem), sb = sa + 4 * 16384;        // A: 128 x 64 per CTA, B: (N/2) x 64 per CTA
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t ad = make_desc(sa + (i & 3) * 16384), bd = make_desc(sb + (i & 3) * 16384);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) mma2(tm, ad + 2 * kk, bd + 2 * kk, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 4 * 16384 + 4 * 16384 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int N : {32, 64, 128, 256}) {
    for (int grid : {2, 148}) {
      k<<<grid, 128, smem>>>(N, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("cta_group::2 M 256 N %3d grid %3d : %7.1f cycles/MMA  (%s)\n", N, grid, (double)c / (iters * 4.0), cudaGetErrorString(e));
    }
  }
  return 0;
}

// Does cuTensorMapEncodeTiled accept a row stride SMALLER than the row extent (overlapping rows), and does the TMA unit
// deliver the overlapped windows?  A [lines, P, 16] bf16 pixel buffer is described as [lines, P, 64]: "pixel" p of the view
// is the 128-byte window over memory pixels p .. p+3 (the conv1 operand without the 4x horizontal-tap replication).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_overlap tma_overlap.cu -lcuda && ./tma_overlap
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap tm, uint16_t* out, int c1, int c2, int box_px, int box_lines) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem), ba = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t bytes = (uint32_t)(box_px * box_lines * 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(bytes));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(sa), "l"((uint64_t)&tm), "r"(0), "r"(c1), "r"(c2), "r"(ba) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(ba) : "memory");
  for (int i = threadIdx.x; i < (int)bytes / 2; i += blockDim.x) out[i] = ((uint16_t*)smem)[i];
}

int main() {
  const int P = 228, L = 40, C = 16, BOXP = 8, BOXL = 5;
  std::vector<uint16_t> h((size_t)L * P * C + 64);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i * 2654435761u >> 16);
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&o, BOXP * BOXL * 128);
  CUtensorMap tm;
  cuuint64_t gdim[3] = {64, (cuuint64_t)P, (cuuint64_t)L};
  cuuint64_t gstride[2] = {(cuuint64_t)C * 2, (cuuint64_t)P * C * 2};
  cuuint32_t box[3] = {64, BOXP, BOXL};
  cuuint32_t estr[3] = {1, 1, 1};
  cuInit(0);
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (row stride 32 B < row extent 128 B): CUresult %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  const int c1 = 219, c2 = 3;      // a box that also runs past the end of the line (px 219..226 of P = 228 -> in range; memory px + 3 wraps)
  k<<<1, 128, BOXP * BOXL * 128 + 1024>>>(tm, o, c1, c2, BOXP, BOXL);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint16_t> out(BOXP * BOXL * 64);
  cudaMemcpy(out.data(), o, out.size() * 2, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int l = 0; l < BOXL; ++l)
    for (int p = 0; p < BOXP; ++p)
      for (int c = 0; c < 64; ++c) {
        const int row = l * BOXP + p;                                   // 128-byte row inside the box
        const int piece = (c / 8) ^ (row & 7);                          // SWIZZLE_128B: 16-byte piece index XOR (row mod 8)
        const uint16_t got = out[(size_t)row * 64 + piece * 8 + (c & 7)];
        const uint16_t want = h[((size_t)(c2 + l) * P + (c1 + p)) * C + c];   // = memory pixel (c1 + p + c / 16), channel c % 16
        bad += got != want;
      }
  printf("mismatches: %ld of %d\n", bad, BOXP * BOXL * 64);
  return bad != 0;
}

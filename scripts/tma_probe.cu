// Bring-up probe (not part of the library): which un-swizzled / swizzled TMA box shapes and out-of-range start coordinates execute on this GPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/tma_probe scripts/tma_probe.cu -lcuda ; run on the GPU box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int box0, int box1, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = smem_u32(&bar), dst = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(box0 * box1 * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(&tm), "r"(b), "r"(c0), "r"(c1), "r"(0), "r"(0), "r"(0) : "memory");
  }
  uint32_t done = 0;
  for (int spin = 0; !done && spin < 100000; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b) : "memory");
  __syncthreads();
  const float* s = reinterpret_cast<const float*>(smem + (dst - smem_u32(smem)));
  for (int i = threadIdx.x; i < box0 * box1; i += blockDim.x) out[i] = done ? s[i] : -12345.f;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  const int W = 160, Hh = 100;   // global matrix [Hh][W], value = row*1000 + col
  std::vector<float> h(W * Hh);
  for (int r = 0; r < Hh; ++r) for (int c = 0; c < W; ++c) h[r * W + c] = r * 1000.f + c;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 256 * 256 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  struct Cfg { int sw, box0, box1, c0, c1; } cfgs[] = {
      {1, 32, 32, 4, 0}, {1, 32, 32, -28, 32}, {0, 32, 32, 4, 0}, {0, 32, 32, -28, 32}, {0, 64, 32, 8, 0}, {0, 96, 32, 4, 0},
      {0, 96, 32, -28, 32}, {0, 100, 32, 4, 0}, {0, 100, 32, -28, 32}, {0, 100, 32, -92, 96}, {0, 100, 32, 140, 90}, {0, 100, 32, -200, 0}};
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (auto& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t gdim[5] = {(cuuint64_t)W, (cuuint64_t)Hh, 1, 1, 1}, gstr[4] = {(cuuint64_t)W * 4, (cuuint64_t)W * Hh * 4, (cuuint64_t)W * Hh * 4, (cuuint64_t)W * Hh * 4};
    cuuint32_t box[5] = {(cuuint32_t)c.box0, (cuuint32_t)c.box1, 1, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     c.sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("sw=%d box=%dx%d: encode failed (%d)\n", c.sw, c.box0, c.box1, (int)r); continue; }
    probe<<<1, 128, 100 * 1024>>>(tm, c.c0, c.c1, c.box0, c.box1, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sw=%d box=%dx%d at (%d,%d): %s\n", c.sw, c.box0, c.box1, c.c0, c.c1, cudaGetErrorString(e)); return 1; }
    std::vector<float> g(c.box0 * c.box1);
    cudaMemcpy(g.data(), o, g.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    if (!c.sw) for (int r2 = 0; r2 < c.box1; ++r2) for (int cc = 0; cc < c.box0; ++cc) {
      const int gr = c.c1 + r2, gc = c.c0 + cc;
      const float want = (gr >= 0 && gr < Hh && gc >= 0 && gc < W) ? gr * 1000.f + gc : 0.f;
      if (g[r2 * c.box0 + cc] != want) ++bad;
    }
    printf("sw=%d box=%dx%d at (%d,%d): ok, first=%g mismatches(unswizzled layout check)=%d\n", c.sw, c.box0, c.box1, c.c0, c.c1, g[0], bad);
  }
  return 0;
}

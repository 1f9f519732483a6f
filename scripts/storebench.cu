// Micro-benchmark (not part of the library): store pattern of the GEMM epilogue without any MMA.
// Each CTA (320 threads, 8 storing warps) writes tiles of 128 rows x 256 fp32 columns (row pitch `pitch` floats) exactly like
// gemm_tf32x3_2cta_kernel's epilogue: warp (q, half) -> rows q*32.., columns half*128..; per access 4 rows x 128 B (float4 per lane).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__global__ void __launch_bounds__(320) store_tiles(float* out, long long pitch, int tiles_per_cta, int tiles_n, int mode) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 2) return;
  const int e = warp - 2, q = warp & 3, half = e >> 2;
  const int rsub = lane >> 3, c4 = (lane & 7) * 4;
  for (int t = 0; t < tiles_per_cta; ++t) {
    const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
    const long long mt = tile / tiles_n, nt = tile % tiles_n;
    float* base = out + (mt * 128 + q * 32 + rsub) * pitch + nt * 256 + half * 128 + c4;
    for (int j = 0; j < 4; ++j)
      for (int it = 0; it < 8; ++it) {
        float4 v = make_float4(t, j, it, lane);
        float* p = base + (long long)(it * 4) * pitch + j * 32;
        if (mode == 0) *reinterpret_cast<float4*>(p) = v;
        else asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
  }
}
int main() {
  const long long rows = 128LL * 148 * 16, tiles_n = 7;   // 7 n-tiles of 256 columns
  for (long long pitch : {1792LL, 1888LL, 2048LL}) {
    for (int mode = 0; mode < 2; ++mode) {
      float* d; size_t bytes = (size_t)rows * pitch * 4;
      cudaMalloc(&d, bytes);
      const int tiles = (int)(rows / 128 * tiles_n), per = tiles / 148;
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      store_tiles<<<148, 320>>>(d, pitch, per, (int)tiles_n, mode);
      cudaEventRecord(a);
      for (int r = 0; r < 3; ++r) store_tiles<<<148, 320>>>(d, pitch, per, (int)tiles_n, mode);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      double gb = 3.0 * per * 148 * 128.0 * 256 * 4 / 1e9;
      printf("pitch %lld mode %d: %.1f GB in %.3f ms -> %.1f GB/s (%s)\n", pitch, mode, gb, ms, gb / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
      cudaFree(d);
    }
  }
  return 0;
}

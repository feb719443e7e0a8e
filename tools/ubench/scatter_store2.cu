// Second micro-benchmark behind the fan epilogue: a warp store instruction (32 lanes x 16 B) covers R rows of a [rows][8208 B] array
// with 512 / R contiguous bytes in each; all chunks of a row are written by consecutive instructions of the same warp.
//   nvcc -arch=sm_100a -O3 -o scatter_store2 scatter_store2.cu && ./scatter_store2
#include <cstdio>
#include <cuda_runtime.h>

template <int R>
__global__ void k(float4 *out, long long rows, long long pitch16, int bytes_per_row) {
  constexpr int LPR = 32 / R;                   // lanes per row, 16 B each
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int steps = bytes_per_row / (LPR * 16);
  for (long long rg = warp; rg < rows / R; rg += nwarps) {
    float4 *p = out + (rg * R + lane / LPR) * pitch16 + (lane % LPR);
    for (int s = 0; s < steps; ++s) p[(long long)s * LPR] = make_float4(1.f, 2.f, 3.f, (float)s);
  }
}
template <int R> void run(float4 *d, long long rows, long long pitch16) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 2; ++it) { cudaEventRecord(a); k<R><<<148 * 8, 256>>>(d, rows, pitch16, 8192); cudaEventRecord(b); cudaEventSynchronize(b); }
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%2d rows x %3d B per instruction: %7.3f ms  %7.1f GB/s\n", R, 512 / R, ms, (double)rows * 8192 / ms / 1e6);
}
int main() {
  const long long rows = 512LL * 181, pitch16 = 8208 / 16;
  float4 *d; cudaMalloc(&d, rows * pitch16 * 16);
  run<1>(d, rows, pitch16); run<2>(d, rows, pitch16); run<4>(d, rows, pitch16); run<8>(d, rows, pitch16); run<16>(d, rows, pitch16); run<32>(d, rows, pitch16);
  return 0;
}

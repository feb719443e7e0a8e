// Micro-benchmark behind the delay-and-sum fan epilogue (fan_tc.cu): how fast can an SM write a [rows][8208 B] array when a warp store
// instruction covers 16 rows x 32 B (two lanes per row), as a function of how many consecutive 32-byte chunks of a row the SAME warp
// writes back to back (1 = one sector per 128-byte line now, the other three much later; 4 = whole lines; 16 = 512 contiguous bytes).
//   nvcc -arch=sm_100a -O3 -o scatter_store scatter_store.cu && ./scatter_store
#include <cstdio>
#include <cuda_runtime.h>

template <int RUN>   // consecutive 32-byte chunks per row written by one warp before it moves to other rows
__global__ void k(float4 *out, long long rows, int chunks_per_row, long long pitch16) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long row_groups = rows / 16;
  const int runs = chunks_per_row / RUN;
  // work unit = (run index, row group): RUN chunks of 16 rows; run index slowest so that the other runs of a row come much later
  for (long long u = warp; u < row_groups * runs; u += nwarps) {
    const long long rg = u % row_groups; const int run = (int)(u / row_groups);
    float4 *p = out + (rg * 16 + (lane >> 1)) * pitch16 + (long long)run * RUN * 2 + (lane & 1);
#pragma unroll
    for (int c = 0; c < RUN; ++c) p[2 * c] = make_float4(1.f, 2.f, 3.f, (float)c);
  }
}

template <int RUN> void run(float4 *d, long long rows, int cpr, long long pitch16, const char *what) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(a);
    k<RUN><<<148 * 8, 256>>>(d, rows, cpr, pitch16);
    cudaEventRecord(b); cudaEventSynchronize(b);
  }
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)rows * cpr * 32;
  printf("%-28s %7.3f ms  %7.1f GB/s\n", what, ms, bytes / ms / 1e6);
}

int main() {
  const long long rows = 512LL * 181, pitch16 = 8208 / 16; const int cpr = 256;   // cfg3: 92672 beam rows of 1026 complex bins
  float4 *d; cudaMalloc(&d, rows * pitch16 * 16);
  run<1>(d, rows, cpr, pitch16, "1 sector per visit");
  run<2>(d, rows, cpr, pitch16, "2 sectors (64 B)");
  run<4>(d, rows, cpr, pitch16, "4 sectors (128 B)");
  run<8>(d, rows, cpr, pitch16, "8 sectors (256 B)");
  run<16>(d, rows, cpr, pitch16, "16 sectors (512 B)");
  return 0;
}

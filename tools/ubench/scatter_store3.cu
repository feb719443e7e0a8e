// Third micro-benchmark behind the fan epilogue: 256-bit stores (st.global.v8.f32, sm_100).  A warp store instruction (32 lanes x 32 B)
// covers R rows of a [rows][8208 B] array with 1024 / R contiguous bytes in each; R = 32 is one full 32-byte sector per lane and row,
// the pattern of four bins per (frame, direction) written by ONE lane instead of two 16-byte halves by a lane pair (scatter_store2: 1.8 TB/s).
//   nvcc -arch=sm_100a -O3 -o scatter_store3 scatter_store3.cu && ./scatter_store3
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(void *p, float a, float b, float c, float d) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int R>
__global__ void k(char *out, long long rows, long long pitch, int bytes_per_row) {
  constexpr int LPR = 32 / R;                   // lanes per row, 32 B each
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int steps = bytes_per_row / (LPR * 32);
  for (long long rg = warp; rg < rows / R; rg += nwarps) {
    char *p = out + (rg * R + lane / LPR) * pitch + (lane % LPR) * 32;
    for (int s = 0; s < steps; ++s) st256(p + (long long)s * LPR * 32, 1.f, 2.f, 3.f, (float)s);
  }
}
template <int R> void run(char *d, long long rows, long long pitch) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 2; ++it) { cudaEventRecord(a); k<R><<<148 * 8, 256>>>(d, rows, pitch, 8192); cudaEventRecord(b); cudaEventSynchronize(b); }
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("256-bit stores, %2d rows x %4d B per instruction: %7.3f ms  %7.1f GB/s\n", R, 1024 / R, ms, (double)rows * 8192 / ms / 1e6);
}
int main() {
  const long long rows = 512LL * 181, pitch = 8224;   // 32-byte aligned rows
  char *d; cudaMalloc(&d, rows * pitch);
  run<1>(d, rows, pitch); run<2>(d, rows, pitch); run<4>(d, rows, pitch); run<8>(d, rows, pitch); run<16>(d, rows, pitch); run<32>(d, rows, pitch);
  return 0;
}

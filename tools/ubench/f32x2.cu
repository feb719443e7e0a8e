// Micro-benchmark: issue rate of packed fp32x2 (FADD2 / FFMA2 / FMUL2) against scalar FADD / FFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE> __global__ void kern(float *out, int iters, float s) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  float2 b = make_float2(s, s * 0.5f), c = make_float2(s * 0.25f, -s);
  if (MODE == 0) {          // scalar FADD: 16 per iteration
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i].x += b.x; a[i].y += b.y; }
  } else if (MODE == 1) {   // packed FADD2: 8 per iteration (same flops)
    u64 p[8], pb = pk(b.x, b.y);
    for (int i = 0; i < 8; ++i) p[i] = pk(a[i].x, a[i].y);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(p[i], pb);
    for (int i = 0; i < 8; ++i) a[i] = upk(p[i]);
  } else if (MODE == 2) {   // scalar FFMA: 16 per iteration
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
  } else if (MODE == 3) {   // packed FFMA2: 8 per iteration
    u64 p[8], pb = pk(b.x, b.y), pc = pk(c.x, c.y);
    for (int i = 0; i < 8; ++i) p[i] = pk(a[i].x, a[i].y);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pb, pc);
    for (int i = 0; i < 8; ++i) a[i] = upk(p[i]);
  } else if (MODE == 4) {   // complex multiply, scalar: 8 cmul per iteration (2 FMUL + 2 FFMA each)
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = make_float2(a[i].x * b.x - a[i].y * b.y, a[i].x * b.y + a[i].y * b.x);
  } else if (MODE == 5) {   // complex multiply, packed: (x,x)*(bx,by) then (y,y)*(-by,bx) + prev
    u64 pb = pk(b.x, b.y), pbs = pk(-b.y, b.x);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        u64 xx = pk(a[i].x, a[i].x), yy = pk(a[i].y, a[i].y), t;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(xx), "l"(pb));
        a[i] = upk(fma2(yy, pbs, t));
      }
  }
  float acc = 0.f;
  for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> void run(const char *name, float *d, int iters, double flops_per_iter_thread) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8, block = 256;
  kern<MODE><<<grid, block>>>(d, 16, 1.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<MODE><<<grid, block>>>(d, iters, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double tf = flops_per_iter_thread * iters * grid * block / (ms * 1e-3) / 1e12;
  printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, ms, tf);
}
int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int it = 1 << 16;
  run<0>("FADD scalar (16/iter)", d, it, 16);
  run<1>("FADD2 packed (8/iter)", d, it, 16);
  run<2>("FFMA scalar (16/iter)", d, it, 32);
  run<3>("FFMA2 packed (8/iter)", d, it, 32);
  run<4>("cmul scalar (8/iter)", d, it, 48);
  run<5>("cmul packed (8/iter)", d, it, 48);
  return 0;
}

// Shared device helpers for the mcarray_b200 CUDA path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MCAG_CHECK_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return mcag_set_cuda_error(e__);  \
  } while (0)

int mcag_set_cuda_error(cudaError_t e);          // capi.cu: records cudaGetErrorString, returns MCAG_ERR_CUDA
int mcag_set_error(int code, const char *msg);   // capi.cu

namespace mcag {

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// spectrum row pitch in complex bins: K = N/2+1 padded to an even count so rows are 16-byte aligned
__host__ __device__ constexpr int spec_pitch(int N) { return N / 2 + 2; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a*conj(b)
#ifndef MCAG_NO_F32X2
// packed fp32 (FADD2, sm_100): one issue slot for both components (tools/ubench/f32x2.cu: same lane rate, half the instructions).
// The FFT kernels are issue / shared-memory bound, so this is worth 2.5 % on the fused STFT -> GCC-PHAT kernel (tools/variants.sh).
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long pa, pb, pr; float2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pr));
  return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long pa, pb, pr; float2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pr));
  return r;
}
__device__ __forceinline__ float2 cscale(float2 a, float s_) {
  unsigned long long pa, ps, pr; float2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ps) : "f"(s_));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pa), "l"(ps));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pr));
  return r;
}
#else
__device__ __forceinline__ float2 cscale(float2 a, float s_) { return make_float2(a.x * s_, a.y * s_); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// PHAT whitening of one bin: X/|X|, 0 where |X| = 0 (oracle/CONVENTIONS.md C5)
__device__ __forceinline__ float2 whiten(float2 x) {
  float m2 = x.x * x.x + x.y * x.y;
  float inv = m2 > 0.f ? rsqrtf(m2) : 0.f;
  // one Newton step brings rsqrtf (2 ulp) to ~1 ulp; the correlations are sums of unit phasors
  inv = m2 > 0.f ? inv * (1.5f - 0.5f * m2 * inv * inv) : 0.f;
  return make_float2(x.x * inv, x.y * inv);
}

// Phase ramps: `fx` is a per-bin phase increment in turns as unsigned 0.64 fixed point, so k*fx wraps mod one
// turn exactly in integer arithmetic (no large-argument sincos, SURVEY.md §7 "hard parts").  Returns
// (cos, sin)(2*pi*k*turns).
__device__ __forceinline__ float2 phase_ramp(uint64_t fx, int k) {
  uint64_t ph = fx * (uint64_t)(uint32_t)k;
  int32_t top = (int32_t)(uint32_t)(ph >> 32);            // signed turns in [-0.5, 0.5) * 2^32
  float s, c;
  sincospif((float)top * 4.656612873077393e-10f, &s, &c);  // 2^-31: argument in units of pi
  return make_float2(c, s);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// first-maximum argmax (lowest index wins ties), as wipp::maxidx / ippsMaxIndx
__device__ __forceinline__ void warp_argmax(float &v, int &i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// order-preserving map float -> uint32 (a < b  <=>  key(a) < key(b) for non-NaN a, b with -0 folded into +0 by the caller)
__device__ __forceinline__ unsigned float_order_key(float v) {
  const unsigned u = __float_as_uint(v);
  return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(unsigned k) {
  return __uint_as_float(k ^ ((k & 0x80000000u) ? 0x80000000u : 0xffffffffu));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// one 256-bit store (sm_100: STG.E.256); the address must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(void *p, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6), "f"(a7)
               : "memory");
}

}  // namespace mcag

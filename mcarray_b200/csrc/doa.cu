// K5 doa_select: SteeringBeamforming::computeEnergyInDOA + selectDOA
// (SteeringBeamforming.cpp:132-195) and the correlation smoothing + argmax of FreqGCCBinauralLocalisation
// (BinauralLocalisation.cpp:444-459), as pair-sum / first-order scan / stencil + top-S kernels.
#include "common.cuh"
#include "kernels.h"

namespace mcag {

// esum[bt][d] = sum_p scale * corr[bt][p][d], pairs added in index order (SteeringBeamforming.cpp:137-141)
__global__ void pair_sum_kernel(const float *__restrict__ corr, long long BT, int P, int D, float scale, float *__restrict__ esum) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BT * D) return;
  const long long bt = i / D;
  const int d = (int)(i - bt * D);
  const float *c = corr + bt * P * D + d;
  float acc = 0.f;
  for (int p = 0; p < P; ++p) acc += scale * c[(size_t)p * D];
  esum[i] = acc;
}

// E_t = a * E_{t-1} + esum_t on active frames, E_t = E_{t-1} otherwise; one thread per (stream, direction).  The inputs of 8 frames are
// fetched before the recurrence touches them (esum and energy may be the same buffer, so the compiler cannot hoist the loads itself):
// the scan is a chain of dependent FMAs, not a chain of dependent memory round trips.
__global__ void energy_scan_kernel(const float *esum, int B, int T, int D, float a, const unsigned char *__restrict__ active,
                                   float *__restrict__ state, float *energy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  float e = state[i];
  constexpr int CH = 8;
  for (int t0 = 0; t0 < T; t0 += CH) {
    float in[CH]; bool on[CH];
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const int t = t0 + u;
      in[u] = (t < T) ? esum[((long long)b * T + t) * D + d] : 0.f;
      on[u] = (t < T) && (!active || active[(long long)b * T + t]);
    }
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const int t = t0 + u;
      if (on[u]) e = fmaf(a, e, in[u]);
      if (t < T) energy[((long long)b * T + t) * D + d] = e;
    }
  }
  state[i] = e;
}

// selectDOA: one warp per frame.  s[i] (i < D-2) only needs energy[i-1 .. i+3]:
//   e = (E - m)/(-2m), m = -15 P;  f[i] = (e[i+1]-e[i] < 0) ? 1 : 0;  ff = median3(f) with replicated borders;
//   s[i] = (ff[i+1] - ff[i]) * e[i+1].   Then S rounds of first-maximum argmax, zeroing each winner.
__global__ void select_doa_kernel(const float *__restrict__ energy, long long BT, int D, int n_pairs, int S, int32_t *__restrict__ idx,
                                  float *__restrict__ prob) {
  extern __shared__ float s_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long bt = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (bt >= BT) return;
  float *s = s_all + (size_t)warp * D;
  const float *E = energy + bt * D;
  const float m = -15.0f * (float)n_pairs, inv = 1.0f / (-2.0f * m);
  const int nf = D - 1, ns = D - 2;
  auto en = [&](int d) { return (E[d] - m) * inv; };
  auto fsign = [&](int i) { i = min(max(i, 0), nf - 1); return (en(i + 1) - en(i) < 0.f) ? 1.f : 0.f; };
  auto med = [&](int i) { float a = fsign(i - 1), b = fsign(i), c = fsign(i + 1); return (a + b + c >= 2.f) ? 1.f : 0.f; };
  for (int i = lane; i < ns; i += 32) s[i] = (med(i + 1) - med(i)) * en(i + 1);
  __syncwarp();
  for (int r = 0; r < S; ++r) {
    float bv = -3.0e38f; int bi = 0x7fffffff;
    for (int i = lane; i < ns; i += 32) { float v = s[i]; if (v > bv) { bv = v; bi = i; } }
    warp_argmax(bv, bi);
    if (lane == 0) { s[bi] = 0.f; idx[bt * S + r] = bi + 1; prob[bt * S + r] = bv; }
    __syncwarp();
  }
}

// FreqGCC smoothing: c_t = (1-alpha) corr_t + alpha c_{t-1}, alpha = 0 up to and including the first active frame,
// `mem` afterwards (BinauralLocalisation.cpp:444-448,523); one thread per (stream, delay).
__global__ void curve_scan_kernel(const float *__restrict__ corr, int B, int T, int D, float mem, const unsigned char *__restrict__ active,
                                  float *__restrict__ state, const unsigned char *__restrict__ started_in, float *__restrict__ curves) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  float prev = state[i];
  bool started = started_in[b] != 0;
  // only `prev` and `started` are carried: the loads are unconditional and the loop is unrolled so that eight frames of inputs are in
  // flight per thread (one frame at a time the kernel was bound by the latency of its dependent loads)
#pragma unroll 8
  for (int t = 0; t < T; ++t) {
    const long long o = ((long long)b * T + t) * D + d;
    const float c = corr[o];
    const bool on = !active || active[(long long)b * T + t];
    if (on) {
      const float alpha = started ? mem : 0.f;
      prev = (1.f - alpha) * c + alpha * prev;
      started = true;
    }
    curves[o] = prev;
  }
  state[i] = prev;
}
__global__ void started_update_kernel(int B, int T, const unsigned char *__restrict__ active, unsigned char *__restrict__ started) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || started[b]) return;
  for (int t = 0; t < T; ++t) if (!active || active[(long long)b * T + t]) { started[b] = 1; return; }
}
__global__ void row_argmax_kernel(const float *__restrict__ x, long long rows, int D, int32_t *__restrict__ idx) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float bv = -3.0e38f; int bi = 0x7fffffff;
  for (int i = lane; i < D; i += 32) { float v = x[row * D + i]; if (v > bv) { bv = v; bi = i; } }
  warp_argmax(bv, bi);
  if (lane == 0) idx[row] = bi;
}

// Per-frame first-maximum argmax of a (slice of a) direction map, packed so that an integer MAX reduction over slices (one
// NCCL all-reduce when the grid is sharded over GPUs) yields the global maximum with ties going to the lowest direction index:
//   packed = ordered(E) << 31 | (0x7FFFFFFF - d_global),   ordered() = the monotone float -> uint32 map; always >= 0 as int64.
__global__ void argmax_pack_kernel(const float *__restrict__ x, long long rows, int D, int d_offset, long long *__restrict__ packed) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float bv = -3.0e38f; int bi = 0x7fffffff;
  for (int i = lane; i < D; i += 32) { float v = x[row * D + i]; if (v > bv) { bv = v; bi = i; } }
  warp_argmax(bv, bi);
  if (lane == 0) {
    uint32_t u = __float_as_uint(bv);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    packed[row] = (long long)(((unsigned long long)u << 31) | (unsigned long long)(0x7FFFFFFF - (d_offset + bi)));
  }
}
int k_argmax_pack(const float *x, long long rows, int D, int d_offset, long long *packed, cudaStream_t st) {
  if (rows <= 0) return 0;
  argmax_pack_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, D, d_offset, packed);
  MCAG_CHECK_LAUNCH();
  return 0;
}

int k_pair_sum(const float *corr, long long BT, int P, int D, float scale, float *esum, cudaStream_t st) {
  const long long n = BT * D;
  if (n <= 0) return 0;
  pair_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(corr, BT, P, D, scale, esum);
  MCAG_CHECK_LAUNCH();
  return 0;
}
int k_energy_scan(const float *esum, int B, int T, int D, float a, const unsigned char *active, float *state, float *energy, cudaStream_t st) {
  if (B * D <= 0 || T <= 0) return 0;
  energy_scan_kernel<<<(B * D + 127) / 128, 128, 0, st>>>(esum, B, T, D, a, active, state, energy);
  MCAG_CHECK_LAUNCH();
  return 0;
}
int k_select_doa(const float *energy, long long BT, int D, int n_pairs, int S, int32_t *idx, float *prob, cudaStream_t st) {
  if (BT <= 0) return 0;
  if (D < 3) return mcag_set_error(1, "select_doa: need at least 3 directions");
  int wpb = 4;
  while (wpb > 1 && (size_t)wpb * D * sizeof(float) > 160 * 1024) wpb >>= 1;
  size_t smem = (size_t)wpb * D * sizeof(float);
  if (smem > 200 * 1024) return mcag_set_error(1, "select_doa: direction grid too large");
  cudaFuncSetAttribute(select_doa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  select_doa_kernel<<<(unsigned)((BT + wpb - 1) / wpb), wpb * 32, smem, st>>>(energy, BT, D, n_pairs, S, idx, prob);
  MCAG_CHECK_LAUNCH();
  return 0;
}
int k_curve_scan_argmax(const float *corr, int B, int T, int D, float, float mem, const unsigned char *active, float *state,
                        unsigned char *started, float *curves, int32_t *idx, cudaStream_t st) {
  if (B * D <= 0 || T <= 0) return 0;
  curve_scan_kernel<<<(B * D + 127) / 128, 128, 0, st>>>(corr, B, T, D, mem, active, state, started, curves);
  MCAG_CHECK_LAUNCH();
  started_update_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, T, active, started);
  MCAG_CHECK_LAUNCH();
  const long long rows = (long long)B * T;
  row_argmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(curves, rows, D, idx);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

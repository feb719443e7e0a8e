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
    if (bi >= ns) { bi = 0; bv = s[0]; }   // every candidate NaN (a non-finite input sample): index 0 like wipp::maxidx, never out of bounds
    if (lane == 0) { s[bi] = 0.f; idx[bt * S + r] = bi + 1; prob[bt * S + r] = bv; }
    __syncwarp();
  }
}

// FreqGCC smoothing: c_t = (1-alpha) corr_t + alpha c_{t-1} on frames the gate lets through (BinauralLocalisation.cpp:444-448).  alpha is
// the reference's _corrMemoryFactor state machine (:323,523,528-561): 0 at construction, `mem` after every voiced frame; on a gated
// frame with the noise floor estimated it is `mem` while fewer than windowsToDecay = 3 fs/(N/2) silent frames have passed and 0
// afterwards (so the first voiced frame after a long pause REPLACES the curve), and the silent-frame counter advances.  One thread per
// (stream, delay); every thread of a stream replays the same (alpha, silence) sequence from the carried FgState, fg_track_kernel
// writes the state back.  `est` [B][T]: noise floor estimated when the frame was gated (gate_kernel); NULL = always.
__global__ void curve_scan_kernel(const float *__restrict__ corr, int B, int T, int D, float mem, int windows_to_decay,
                                  const unsigned char *__restrict__ active, const unsigned char *__restrict__ est, float *__restrict__ state,
                                  const FgState *__restrict__ fg, float *__restrict__ curves) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  float prev = state[i];
  float alpha = fg[b].alpha;
  int silence = fg[b].silence;
  // the loads are unconditional and the loop is unrolled so that eight frames of inputs are in flight per thread (one frame at a
  // time the kernel was bound by the latency of its dependent loads)
#pragma unroll 8
  for (int t = 0; t < T; ++t) {
    const long long o = ((long long)b * T + t) * D + d;
    const float c = corr[o];
    const bool on = !active || active[(long long)b * T + t];
    const bool e = !est || est[(long long)b * T + t];
    if (on) {
      prev = (1.f - alpha) * c + alpha * prev;
      alpha = mem; silence = 0;
    } else if (e) {
      alpha = silence < windows_to_decay ? mem : 0.f;
      ++silence;
    }
    curves[o] = prev;
  }
  state[i] = prev;
}

// Per-stream state of FreqGCCBinauralLocalisation after the frames of a call: the memory-factor state machine above, plus (track != 0)
// the deterministic DOA tracker = the `#else` branch of USE_PARTICLE_FILTER (BinauralLocalisation.cpp:501-504):
//   setProbability(_currentDOA, _prob, 1) on this frame's smoothed curve with the PREVIOUS _currentDOA (:454,569-631),
//   _currentDOA = _doaMemoryFactor * _currentDOA + (1 - _doaMemoryFactor) * doaIdx2angle(argmax)  (float factor, double DOA),
//   _doaMemoryFactor: 0 -> dmem (0.6f) after a voiced frame (:524), 0 again after windowsToDecay silent frames (:555).
// One warp per stream, frames in order; lanes share the 61-cell min / sum reductions, lane 0 carries the doubles.
__global__ void __launch_bounds__(128) fg_track_kernel(const float *__restrict__ curves, const int32_t *__restrict__ idx, int B, int T, int D, float mem,
                                                      float dmem, int windows_to_decay, float doa_step, int track,
                                                      const unsigned char *__restrict__ active, const unsigned char *__restrict__ est,
                                                      FgState *__restrict__ fg, double *__restrict__ track_doa, float *__restrict__ track_prob) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  FgState s = fg[b];
  auto cell_angle = [&](int i) { return (double)(float)((double)((float)i * doa_step) - 1.57079632679489661923); };   // doaIdx2angle (float-typed)
  for (int t = 0; t < T; ++t) {
    const long long bt = (long long)b * T + t;
    const bool on = !active || active[bt];
    const bool e = !est || est[bt];
    if (on) {
      if (track) {
        const float *c = curves + bt * D;
        float mn = 3.0e38f; double sm = 0.0;
        for (int i = lane; i < D; i += 32) { const float v = c[i]; mn = fminf(mn, v); sm += (double)v; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); sm += __shfl_xor_sync(0xffffffffu, sm, o); }
        if (lane == 0) {
          sm = __dsub_rn(sm, __dmul_rn((double)mn, (double)D));
          // angle2DOAidx (microhponeArrayHelpers.cpp:110-115) of the previous DOA, as a float
          float ang = (float)s.doa;
          ang = (float)fmax((double)ang, -1.57079632679489661923);
          ang = (float)fmin((double)ang, 1.57079632679489661923);
          const int ci = (int)(((double)ang + 1.57079632679489661923) / (double)doa_step);
          const double angle = cell_angle(ci);
          double p;
          if (0 < ci && ci < D - 1) {
            double pc, pd, nc, nd;
            if (angle > s.doa) { pc = c[ci - 1]; pd = cell_angle(ci - 1); nc = c[ci]; nd = angle; }
            else { pc = c[ci]; pd = angle; nc = c[ci + 1]; nd = cell_angle(ci + 1); }
            const double slope = __ddiv_rn(__dsub_rn(nc, pc), __dsub_rn(nd, pd));
            p = __dadd_rn(__dmul_rn(slope, __dsub_rn(s.doa, pd)), pc);
          } else {
            p = c[min(max(ci, 0), D - 1)];
          }
          double pr = 0.0;
          if (sm > 0.0) pr = __ddiv_rn(__dsub_rn(p, (double)mn), sm);
          s.prob = (pr < 0.01) ? 0.f : (float)pr;
          const double doa = cell_angle(idx[bt]);
          s.doa = __dadd_rn(__dmul_rn((double)s.dalpha, s.doa), __dmul_rn((double)(1.f - s.dalpha), doa));
        }
      }
      s.alpha = mem; s.dalpha = dmem; s.silence = 0;
    } else if (e) {
      const bool decay = s.silence < windows_to_decay;
      s.alpha = decay ? mem : 0.f; s.dalpha = decay ? dmem : 0.f;
      ++s.silence;
    }
    if (track && lane == 0) { track_doa[bt] = s.doa; track_prob[bt] = s.prob; }
  }
  if (lane == 0) fg[b] = s;
}
__global__ void row_argmax_kernel(const float *__restrict__ x, long long rows, int D, int32_t *__restrict__ idx) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float bv = -3.0e38f; int bi = 0x7fffffff;
  for (int i = lane; i < D; i += 32) { float v = x[row * D + i]; if (v > bv) { bv = v; bi = i; } }
  warp_argmax(bv, bi);
  if (bi >= D) bi = 0;   // all-NaN row: index 0 like wipp::maxidx
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
  if (bi >= D) bi = 0;   // all-NaN row: index 0 like wipp::maxidx
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
int k_curve_scan_argmax(const float *corr, int B, int T, int D, float mem, float dmem, int windows_to_decay, float doa_step, int track,
                        const unsigned char *active, const unsigned char *est, float *state, FgState *fg, float *curves, int32_t *idx,
                        double *track_doa, float *track_prob, cudaStream_t st) {
  if (B * D <= 0 || T <= 0) return 0;
  curve_scan_kernel<<<(B * D + 127) / 128, 128, 0, st>>>(corr, B, T, D, mem, windows_to_decay, active, est, state, fg, curves);
  MCAG_CHECK_LAUNCH();
  const long long rows = (long long)B * T;
  row_argmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(curves, rows, D, idx);
  MCAG_CHECK_LAUNCH();
  fg_track_kernel<<<(B + 3) / 4, 128, 0, st>>>(curves, idx, B, T, D, mem, dmem, windows_to_decay, doa_step, track, active, est, fg, track_doa, track_prob);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

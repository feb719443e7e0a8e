// K2 gcc_phat: replaces dsp::GeneralisedCrossCorrelation as driven by
// SteeringBeamforming::computeCorrelations (SteeringBeamforming.cpp:104-130) and
// FreqGCCBinauralLocalisation::processParametrisation (BinauralLocalisation.cpp:438-444).
//   k_tdoa_lags : integer-lag mode (BASELINE config 2) — PHAT cross-spectrum, inverse real FFT in shared memory,
//                 lag-window extraction and first-maximum argmax fused in one kernel.
//   k_gcc_tau   : fractional-delay (tau grid) mode, the reference's own semantics, as a register-tiled contraction
//                 of PHAT cross-spectra against steering phasors generated on the fly from fixed-point phase ramps.
#include "fft.cuh"
#include "kernels.h"

namespace mcag {

// ---------------------------------------------------------------------------------------------------
// integer-lag GCC-PHAT.  One CTA per (frame, stream): the M whitened spectra are staged once in shared
// memory, then each group of N/16 threads runs one pair at a time:
//   Z[k] = E[k] + i O[k] from G = U_i conj(U_j)  ->  N/2-point inverse complex FFT  ->  r[l]/2 in packed form
//   S[l] = r[l]/2 + (Re G[0] + (-1)^l Re G[N/2])/2   (one-sided sum of oracle/CONVENTIONS.md C5)
// ---------------------------------------------------------------------------------------------------
template <int N, int G>
__global__ void __launch_bounds__(G *(N / 16)) tdoa_kernel(const float2 *__restrict__ spec, int T, int M, int max_lag,
                                                            const float2 *__restrict__ tw_g, float *__restrict__ curves,
                                                            int32_t *__restrict__ lags, float *__restrict__ peaks) {
  constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), NT = G * TPF;
  const int P = M * (M - 1) / 2, L = 2 * max_lag + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *s_U = reinterpret_cast<float2 *>(smem_raw);                 // M * KP
  float2 *s_tw = s_U + (size_t)M * KP;                                // NC
  float2 *s_buf = s_tw + NC;                                          // G * fft_buf_len(NC)
  float *s_curve = reinterpret_cast<float *>(s_buf + G * fft_buf_len(NC));   // G * L
  unsigned char *s_pair = reinterpret_cast<unsigned char *>(s_curve + G * L);   // 2 * P

  const int tid = threadIdx.x, t = blockIdx.x, b = blockIdx.y;
  const float2 *src = spec + ((long long)b * T + t) * M * KP;
  for (int i = tid; i < M * KP; i += NT) {
    const int k = i % KP;
    s_U[i] = (k <= NC) ? whiten(src[i]) : make_float2(0.f, 0.f);
  }
  for (int i = tid; i < NC; i += NT) s_tw[i] = tw_g[i];
  for (int p = tid; p < P; p += NT) {   // pair p -> (i, j), i < j lexicographic (SteeringBeamforming.cpp:63-65)
    int i = 0, rem = p;
    while (rem >= M - 1 - i) { rem -= M - 1 - i; ++i; }
    s_pair[2 * p] = (unsigned char)i;
    s_pair[2 * p + 1] = (unsigned char)(i + 1 + rem);
  }
  __syncthreads();

  const int g = tid / TPF, j = tid % TPF;
  float2 *buf = s_buf + g * fft_buf_len(NC);
  float *curve = s_curve + g * L;
  const int rounds = (P + G - 1) / G;
  for (int it = 0; it < rounds; ++it) {   // uniform trip count: every thread runs every round, stores are predicated
    const int p = it * G + g;
    const bool live = p < P;
    const int pc = live ? p : P - 1;
    const float2 *Ui = s_U + (size_t)s_pair[2 * pc] * KP, *Uj = s_U + (size_t)s_pair[2 * pc + 1] * KP;
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int k = j + r * (NC / 8);
      float2 gk = cmulc(Ui[k], Uj[k]), gn = cmulc(Ui[NC - k], Uj[NC - k]);
      if (k == 0) { gk.y = 0.f; gn.y = 0.f; }
      float2 e = make_float2(0.5f * (gk.x + gn.x), 0.5f * (gk.y - gn.y));
      float2 d = make_float2(0.5f * (gk.x - gn.x), 0.5f * (gk.y + gn.y));
      float2 o = cmul(d, tw_lookup<true>(s_tw, k, NC));
      v[r] = make_float2(e.x - o.y, e.y + o.x);
    }
    fft_run<NC, true>(v, buf, s_tw, j, g);
    const float g0 = Ui[0].x * Uj[0].x, gny = Ui[NC].x * Uj[NC].x;   // both spectra are real at DC / Nyquist
    float best = -3.0e38f; int besti = 0x7fffffff;
    for (int c = j; c < L; c += TPF) {
      const int l = c - max_lag;
      const int li = (l + N) & (N - 1);
      const float2 z = buf[fft_pad(li >> 1)];
      const float s = ((li & 1) ? z.y : z.x) + 0.5f * (g0 + ((l & 1) ? -gny : gny));
      curve[c] = s;
      if (s > best) { best = s; besti = c; }   // ascending c per thread: first maximum kept
    }
    // first-maximum argmax across the group
    if constexpr (TPF >= 32) {
      warp_argmax(best, besti);
      constexpr int WPF = TPF / 32;
      __shared__ float s_bv[G * (WPF > 0 ? WPF : 1)];
      __shared__ int s_bi[G * (WPF > 0 ? WPF : 1)];
      if ((tid & 31) == 0) { s_bv[g * WPF + (j >> 5)] = best; s_bi[g * WPF + (j >> 5)] = besti; }
      group_sync<TPF>(g);
      if (j == 0 && live) {
        float bv = s_bv[g * WPF]; int bi = s_bi[g * WPF];
        for (int w = 1; w < WPF; ++w) { float ov = s_bv[g * WPF + w]; int oi = s_bi[g * WPF + w]; if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; } }
        lags[((long long)b * T + t) * P + p] = bi - max_lag;
        if (peaks) peaks[((long long)b * T + t) * P + p] = bv;
      }
    } else {
#pragma unroll
      for (int o = TPF / 2; o > 0; o >>= 1) {   // stays inside the aligned TPF-lane segment
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (j == 0 && live) {
        lags[((long long)b * T + t) * P + p] = besti - max_lag;
        if (peaks) peaks[((long long)b * T + t) * P + p] = best;
      }
      group_sync<TPF>(g);
    }
    if (curves && live) {
      float *dst = curves + (((long long)b * T + t) * P + p) * L;
      for (int c = j; c < L; c += TPF) dst[c] = curve[c];
    }
    group_sync<TPF>(g);
  }
}

template <int N> static int launch_tdoa(const float2 *spec, int B, int T, int M, int max_lag, const float2 *tw, float *curves, int32_t *lags,
                                        float *peaks, cudaStream_t st) {
  constexpr int NC = N / 2, TPF = NC / 8;
  constexpr int G = (TPF >= 128) ? 2 : (256 / TPF);
  const int P = M * (M - 1) / 2, L = 2 * max_lag + 1;
  size_t smem = sizeof(float2) * ((size_t)M * spec_pitch(N) + NC + (size_t)G * fft_buf_len(NC)) + sizeof(float) * G * L + 2 * P + 16;
  if (smem > 220 * 1024) return mcag_set_error(1, "tdoa: M*N too large for the shared-memory staged kernel");
  auto kern = tdoa_kernel<N, G>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<dim3(T, B), G * TPF, smem, st>>>(spec, T, M, max_lag, tw, curves, lags, peaks);
  MCAG_CHECK_LAUNCH();
  return 0;
}

int k_tdoa_lags(const float2 *spec, int B, int T, int M, int N, int max_lag, const float2 *tw, float *curves, int32_t *lags, float *peaks,
                cudaStream_t st) {
  if (T <= 0 || B <= 0) return 0;
  if (M < 2 || M > 255) return mcag_set_error(1, "tdoa: need 2..255 channels");
  if (max_lag < 0 || max_lag > N / 2 - 1) return mcag_set_error(1, "tdoa: max_lag out of range");
  switch (N) {
    case 256: return launch_tdoa<256>(spec, B, T, M, max_lag, tw, curves, lags, peaks, st);
    case 512: return launch_tdoa<512>(spec, B, T, M, max_lag, tw, curves, lags, peaks, st);
    case 1024: return launch_tdoa<1024>(spec, B, T, M, max_lag, tw, curves, lags, peaks, st);
    case 2048: return launch_tdoa<2048>(spec, B, T, M, max_lag, tw, curves, lags, peaks, st);
  }
  return mcag_set_error(1, "tdoa: frame size must be 256, 512, 1024 or 2048");
}

// ---------------------------------------------------------------------------------------------------
// tau-grid GCC-PHAT, pair form: corr[b][t][p][d] = Re sum_k G_p[t][k] exp(+j 2 pi k tau_pd / N).
// CTA = (64-frame tile, pair, stream); 128 threads = 16 frame-threads x 8 direction-threads, each owning a
// 4 x DPT register tile.  Per 32-bin chunk the PHAT cross-spectra G[f][k] and the phasors W[d][k] are staged in
// shared memory; W is generated from 0.64 fixed-point phase increments (no table traffic, exact range reduction).
// ---------------------------------------------------------------------------------------------------
constexpr int GT_TF = 64, GT_KC = 32, GT_DT = 8;

template <int DPT>
__global__ void __launch_bounds__(128) gcc_tau_kernel(const float2 *__restrict__ spec, int T, int M, int N, const uint64_t *__restrict__ pair_fx,
                                                       int D, int d_base, float *__restrict__ corr) {
  constexpr int DTOT = GT_DT * DPT, PITCH = GT_KC + 1;
  __shared__ float2 s_G[GT_TF * PITCH];
  __shared__ float2 s_W[DTOT * PITCH];
  const int KP = spec_pitch(N), K = N / 2 + 1, P = M * (M - 1) / 2;
  const int tid = threadIdx.x, t0 = blockIdx.x * GT_TF, p = blockIdx.y, b = blockIdx.z;
  int mi = 0, rem = p;
  while (rem >= M - 1 - mi) { rem -= M - 1 - mi; ++mi; }
  const int mj = mi + 1 + rem;
  const int ft = tid / GT_DT, dt = tid % GT_DT;
  float acc[4][DPT];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < DPT; ++c) acc[a][c] = 0.f;

  for (int k0 = 0; k0 < K; k0 += GT_KC) {
    for (int idx = tid; idx < GT_TF * GT_KC; idx += 128) {
      const int f = idx / GT_KC, kk = idx % GT_KC, k = k0 + kk, t = t0 + f;
      float2 gph = make_float2(0.f, 0.f);
      if (k < K && t < T) {
        const float2 *row = spec + ((long long)b * T + t) * M * KP;
        gph = whiten(cmulc(row[(size_t)mi * KP + k], row[(size_t)mj * KP + k]));
      }
      s_G[f * PITCH + kk] = gph;
    }
    for (int idx = tid; idx < DTOT * GT_KC; idx += 128) {
      const int dl = idx / GT_KC, kk = idx % GT_KC, d = d_base + dl;
      float2 w = make_float2(0.f, 0.f);
      if (d < D) w = phase_ramp(pair_fx[(size_t)p * D + d], k0 + kk);
      s_W[dl * PITCH + kk] = w;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < GT_KC; ++kk) {
      float2 gv[4], wv[DPT];
#pragma unroll
      for (int a = 0; a < 4; ++a) gv[a] = s_G[(ft + 16 * a) * PITCH + kk];
#pragma unroll
      for (int c = 0; c < DPT; ++c) wv[c] = s_W[(dt + GT_DT * c) * PITCH + kk];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < DPT; ++c) acc[a][c] = fmaf(gv[a].x, wv[c].x, fmaf(-gv[a].y, wv[c].y, acc[a][c]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int t = t0 + ft + 16 * a;
    if (t >= T) continue;
    float *dst = corr + (((long long)b * T + t) * P + p) * D;
#pragma unroll
    for (int c = 0; c < DPT; ++c) {
      const int d = d_base + dt + GT_DT * c;
      if (d < D) dst[d] = acc[a][c];
    }
  }
}

int k_gcc_tau(const float2 *spec, int B, int T, int M, int N, const uint64_t *pair_fx, int D, float *corr, cudaStream_t st) {
  if (T <= 0 || B <= 0) return 0;
  const int P = M * (M - 1) / 2;
  dim3 grid((T + GT_TF - 1) / GT_TF, P, B);
  if (P > 65535 || B > 65535) return mcag_set_error(1, "gcc_tau: too many pairs or streams for one launch");
  if (D <= GT_DT * 5) {
    gcc_tau_kernel<5><<<grid, 128, 0, st>>>(spec, T, M, N, pair_fx, D, 0, corr);
    MCAG_CHECK_LAUNCH();
  } else {
    for (int d0 = 0; d0 < D; d0 += GT_DT * 8) {
      gcc_tau_kernel<8><<<grid, 128, 0, st>>>(spec, T, M, N, pair_fx, D, d0, corr);
      MCAG_CHECK_LAUNCH();
    }
  }
  return 0;
}

}  // namespace mcag

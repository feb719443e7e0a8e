// K6 binaural_mask: FastBinauralMasking::processParametrisation and helpers
// (FastBinauralMasking.cpp:126-538, constants FastBinauralMasking.h:111-128), split into
//   stats  : per (frame, band) reductions of the band-filtered spectra (H_b is real, so |H_b X|^2 = H_b^2 |X|^2)
//   scan   : the per-band first-order power tracker over frames plus every mask decision / gain (sequential in t)
//   apply  : out[k] = X[k] * sum_b gain_b * H_b[k]  (gain only on bins 0..N/2-1, the Nyquist bin is summed unmasked)
#include "common.cuh"
#include "kernels.h"
#include "mask_common.cuh"

namespace mcag {

// The mel bands are narrow (a triangular filter covers a few per cent of the bins), so every kernel first finds the non-zero bin
// range of each band (and the band range of each bin) from the coefficient table; the frame loops then touch only those.
//
// Persistent CTAs, one warp per frame at a time.  stats[bt][b][0] = sum_{k<N/2} H2 |(L+R)/2|^2      (getFramePower :496-538)
//                               [1] = sum_{k<K}   H2 Re(R conj L)        (normaliseFFTCorrelation :437-441)
//                               [2],[3] = sum_{k<K} H2 |L|^2, |R|^2      (:446-452, maskFrameByScaling :262-264)
//                               [4],[5] = sum_{k<N/2} H2 |L|^2, |R|^2    (noisyFrame -> getPower :222,520-538)
// Each band is summed by one lane in ascending bin order (the reference's order; zero coefficients add exact zeros).
constexpr int MS_WARPS = 8;
__global__ void __launch_bounds__(32 * MS_WARPS) mask_stats_kernel(const float2 *__restrict__ spec, long long BT, int N, const float *__restrict__ H2,
                                                                   int nb, float *__restrict__ stats) {
  extern __shared__ float4 s_q_all[];   // [MS_WARPS][K] per bin: (|L+R|^2/4, Re(R L*), |L|^2, |R|^2), then the band ranges
  const int KP = spec_pitch(N), K = N / 2 + 1, NH = N / 2;
  int *s_lo = reinterpret_cast<int *>(s_q_all + (size_t)MS_WARPS * K), *s_hi = s_lo + nb;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const float *h = H2 + (size_t)b * KP;
    int lo = K, hi = 0;
    for (int k = 0; k < K; ++k)
      if (h[k] != 0.f) { lo = min(lo, k); hi = k + 1; }
    s_lo[b] = lo; s_hi[b] = hi;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 *s_q = s_q_all + (size_t)warp * K;
  for (long long bt = (long long)blockIdx.x * MS_WARPS + warp; bt < BT; bt += (long long)gridDim.x * MS_WARPS) {
    const float2 *L = spec + bt * 2 * KP, *R = L + KP;
    for (int k = lane; k < K; k += 32) {
      const float2 l = L[k], r = R[k];
      const float mx = 0.5f * l.x + 0.5f * r.x, my = 0.5f * l.y + 0.5f * r.y;
      s_q[k] = make_float4(mx * mx + my * my, r.x * l.x + r.y * l.y, l.x * l.x + l.y * l.y, r.x * r.x + r.y * r.y);
    }
    __syncwarp();
    for (int b = lane; b < nb; b += 32) {
      const float *h = H2 + (size_t)b * KP;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
      const int hi = s_hi[b];
      for (int k = s_lo[b]; k < hi; ++k) {
        const float w = __ldg(h + k);
        const float4 q = s_q[k];
        a1 = fmaf(w, q.y, a1); a2 = fmaf(w, q.z, a2); a3 = fmaf(w, q.w, a3);
        if (k < NH) { a0 = fmaf(w, q.x, a0); a4 = fmaf(w, q.z, a4); a5 = fmaf(w, q.w, a5); }
      }
      float *o = stats + (bt * nb + b) * MS_NSTAT;
      o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3; o[4] = a4; o[5] = a5;
    }
    __syncwarp();
  }
}

// one thread per (stream, band), sequential over frames.  method / alg enums: ArrayModules.h:81,89.
__global__ void mask_scan_kernel(const float *__restrict__ stats, int B, int T, int N, int nb, int method, int alg,
                                 const float *__restrict__ thresholds, float *__restrict__ Qs, float *__restrict__ noise_s,
                                 int first_call, float *__restrict__ gains, unsigned char *__restrict__ decisions,
                                 float *__restrict__ q_trace) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nb) return;
  const int s = i / nb, b = i - s * nb;
  float Q = Qs[i], noise = noise_s[i];
  const float thr = thresholds[b];
  for (int t = 0; t < T; ++t) {
    const long long bt = (long long)s * T + t;
    float gl, gr;
    const int dec = mask_band_step(stats + (bt * nb + b) * MS_NSTAT, N, method, alg, thr, first_call, Q, noise, gl, gr);
    gains[(bt * nb + b) * 2] = gl;
    gains[(bt * nb + b) * 2 + 1] = gr;
    if (decisions) decisions[bt * nb + b] = (unsigned char)dec;
    if (q_trace) q_trace[bt * nb + b] = Q;
    ++first_call;                                                   // :193-197
  }
  Qs[i] = Q; noise_s[i] = noise;   // the frame counter (_firstCall) is common to all streams and lives on the host
}

// persistent CTAs, one warp per frame at a time, in place; each bin sums only the bands that cover it (in band order).
// Per frame the warp first has its 16-byte loads of both channels in flight (two bins per load), builds the per-bin weights of
// the frame in shared memory meanwhile, then scales and stores: the dependent chain band range -> coefficient -> gain no longer
// sits in front of every HBM access (the one-bin-at-a-time version ran at 47 % of the HBM peak).
constexpr int MA_UNROLL = 5;   // 32 lanes x 5 loads x 2 bins >= N/2+2 bins up to N = 512; larger frames take more rounds
__global__ void __launch_bounds__(32 * MS_WARPS) mask_apply_kernel(float2 *__restrict__ spec, long long BT, int N, const float *__restrict__ H, int nb,
                                                                   const float *__restrict__ gains) {
  extern __shared__ __align__(16) float s_ga[];   // [MS_WARPS][nb][2] gains, [MS_WARPS][KP] float2 weights, then per bin the band range [blo, bhi)
  const int KP = spec_pitch(N), K = N / 2 + 1, NH = N / 2, KP2 = KP / 2;
  float2 *s_w_all = reinterpret_cast<float2 *>(s_ga + (((size_t)MS_WARPS * nb * 2 + 3) & ~(size_t)3));
  int *s_blo = reinterpret_cast<int *>(s_w_all + (size_t)MS_WARPS * KP), *s_bhi = s_blo + K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int lo = nb, hi = 0;
    for (int b = 0; b < nb; ++b)
      if (H[(size_t)b * KP + k] != 0.f) { lo = min(lo, b); hi = b + 1; }
    s_blo[k] = lo; s_bhi[k] = hi;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *s_g = s_ga + (size_t)warp * nb * 2;
  float2 *s_w = s_w_all + (size_t)warp * KP;
  for (long long bt = (long long)blockIdx.x * MS_WARPS + warp; bt < BT; bt += (long long)gridDim.x * MS_WARPS) {
    float4 *L4 = reinterpret_cast<float4 *>(spec + bt * 2 * KP), *R4 = reinterpret_cast<float4 *>(spec + bt * 2 * KP + KP);
    for (int i0 = 0; i0 < KP2; i0 += 32 * MA_UNROLL) {
      float4 l[MA_UNROLL], r[MA_UNROLL];
#pragma unroll
      for (int u = 0; u < MA_UNROLL; ++u) {
        const int i = i0 + lane + 32 * u;
        if (i < KP2) { l[u] = L4[i]; r[u] = R4[i]; }
      }
      if (i0 == 0) {   // the frame's weights, once
        for (int i = lane; i < nb * 2; i += 32) s_g[i] = gains[bt * nb * 2 + i];
        __syncwarp();
        for (int k = lane; k < KP; k += 32) {
          float wl = 0.f, wr = 0.f;
          if (k < K) {
            const int bhi = s_bhi[k];
            for (int b = s_blo[k]; b < bhi; ++b) {
              const float h = __ldg(H + (size_t)b * KP + k);
              wl = fmaf(h, (k < NH) ? s_g[2 * b] : 1.f, wl);
              wr = fmaf(h, (k < NH) ? s_g[2 * b + 1] : 1.f, wr);
            }
          }
          s_w[k] = make_float2(wl, wr);   // the pad bin gets weight 0 (it holds 0)
        }
        __syncwarp();
      }
#pragma unroll
      for (int u = 0; u < MA_UNROLL; ++u) {
        const int i = i0 + lane + 32 * u;
        if (i < KP2) {
          const float4 w = *reinterpret_cast<const float4 *>(s_w + 2 * i);   // (wl, wr) of bins 2i and 2i+1
          L4[i] = make_float4(l[u].x * w.x, l[u].y * w.x, l[u].z * w.z, l[u].w * w.z);
          R4[i] = make_float4(r[u].x * w.y, r[u].y * w.y, r[u].z * w.w, r[u].w * w.w);
        }
      }
    }
    __syncwarp();
  }
}

int k_mask_stats(const float2 *spec, long long BT, int N, const float *H2, int nb, float *stats, cudaStream_t st) {
  if (BT <= 0) return 0;
  size_t smem = sizeof(float4) * MS_WARPS * (N / 2 + 1) + sizeof(int) * 2 * nb;
  cudaFuncSetAttribute(mask_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mask_stats_kernel, 32 * MS_WARPS, smem);
  if (per_sm < 1) per_sm = 1;
  const long long want = (BT + MS_WARPS - 1) / MS_WARPS, cap = (long long)sms * per_sm;
  mask_stats_kernel<<<(unsigned)(want < cap ? want : cap), 32 * MS_WARPS, smem, st>>>(spec, BT, N, H2, nb, stats);
  MCAG_CHECK_LAUNCH();
  return 0;
}
int k_mask_scan(const float *stats, int B, int T, int N, int nb, int method, int alg, const float *thresholds, float *Q, float *noise,
                int first_call, float *gains, unsigned char *decisions, float *q_trace, cudaStream_t st) {
  if (B * nb <= 0 || T <= 0) return 0;
  mask_scan_kernel<<<(B * nb + 63) / 64, 64, 0, st>>>(stats, B, T, N, nb, method, alg, thresholds, Q, noise, first_call, gains, decisions, q_trace);
  MCAG_CHECK_LAUNCH();
  return 0;
}
int k_mask_apply(float2 *spec, long long BT, int N, const float *H, int nb, const float *gains, cudaStream_t st) {
  if (BT <= 0) return 0;
  size_t smem = sizeof(float) * (((size_t)MS_WARPS * nb * 2 + 3) & ~(size_t)3) + sizeof(float2) * MS_WARPS * spec_pitch(N) + sizeof(int) * 2 * (N / 2 + 1);
  if (smem > 48 * 1024) cudaFuncSetAttribute(mask_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // persistent grid = exactly the resident CTAs (a fixed 4 per SM left a half-empty second wave when 3 fit)
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mask_apply_kernel, 32 * MS_WARPS, smem);
  if (per_sm < 1) per_sm = 1;
  const long long want = (BT + MS_WARPS - 1) / MS_WARPS, cap = (long long)sms * per_sm;
  mask_apply_kernel<<<(unsigned)(want < cap ? want : cap), 32 * MS_WARPS, smem, st>>>(spec, BT, N, H, nb, gains);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

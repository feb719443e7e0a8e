// N2 multiband binaural localisation: MultibandBinarualLocalisation::processOneSubband / processSumamry
// (MultibandBinarualLocalisation.cpp:145-259).  Per linear sub-band a GCC-PHAT curve on the DOA grid with 0.4 memory and
// its arg-max, then an energy-weighted histogram of the band arg-maxima whose own arg-max is the published DOA.
//   band    : per frame, the raw band curves sum_{k in band} Re(G[k] W[d][k]), the band energies and the floor power
//   scan    : c = (1-m) raw + m prev over frames, per (stream, band, delay)                                   (:180-183)
//   summary : band arg-maxima, histogram in band order, its sum / arg-max / prob                                (:184-233)
//   gate    : the power-floor state machine and the hold of the published cell on silent frames              (:127-143,214-255)
// PHAT whitening makes the band response drop out of G wherever it is non-zero, so the band curves share one whitened
// cross-spectrum per frame and differ only in the bins they add.
#include "common.cuh"
#include "kernels.h"

namespace mcag {

constexpr int MB_WARPS = 8;

__global__ void __launch_bounds__(32 * MB_WARPS) mb_band_kernel(const float2 *__restrict__ spec, long long BT, int N, const float *__restrict__ H,
                                                                int nb, const float2 *__restrict__ W, int D, float *__restrict__ band_raw,
                                                                float *__restrict__ band_energy, float *__restrict__ floor_pow) {
  extern __shared__ float2 s_all[];   // [MB_WARPS][K] whitened cross-spectrum, then [MB_WARPS][K] weighted bin power, then band ranges
  const int KP = spec_pitch(N), K = N / 2 + 1;
  float *s_pw_all = reinterpret_cast<float *>(s_all + (size_t)MB_WARPS * K);
  int *s_lo = reinterpret_cast<int *>(s_pw_all + (size_t)MB_WARPS * K), *s_hi = s_lo + nb;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const float *h = H + (size_t)b * KP;
    int lo = K, hi = 0;
    for (int k = 0; k < K; ++k)
      if (h[k] != 0.f) { lo = min(lo, k); hi = k + 1; }
    s_lo[b] = lo; s_hi[b] = hi;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 *s_G = s_all + (size_t)warp * K;
  float *s_pw = s_pw_all + (size_t)warp * K;
  const int KF = K / 2;                                         // FFTPower(frames, K): the first K/2 bins of the CCS buffer (:130)
  const float nF = (float)(K - 2);
  for (long long bt = (long long)blockIdx.x * MB_WARPS + warp; bt < BT; bt += (long long)gridDim.x * MB_WARPS) {
    const float2 *L = spec + bt * 2 * KP, *R = L + KP;
    float fl = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float2 l = L[k], r = R[k];
      s_G[k] = whiten(cmulc(l, r));
      const float p = l.x * l.x + l.y * l.y + r.x * r.x + r.y * r.y;
      s_pw[k] = ((k == 0 || k == K - 1) ? 1.f : 2.f) * p;
      if (k < KF) fl += ((k == 0 || k == KF - 1) ? 1.f : 2.f) * p;
    }
    fl = warp_sum(fl);
    if (lane == 0) floor_pow[bt] = 0.5f * fl / (nF * nF);       // mean over the two channels
    __syncwarp();
    for (int i = lane; i < nb * D; i += 32) {
      const int b = i / D, d = i - b * D;
      const float *h = H + (size_t)b * KP;
      const float2 *w = W + (size_t)d * KP;
      float acc = 0.f;
      const int hi = s_hi[b];
      for (int k = s_lo[b]; k < hi; ++k) {
        if (__ldg(h + k) == 0.f) continue;
        const float2 g = s_G[k], a = __ldg(w + k);
        acc = fmaf(g.x, a.x, fmaf(-g.y, a.y, acc));
      }
      band_raw[(bt * nb + b) * D + d] = acc;
    }
    for (int b = lane; b < nb; b += 32) {                       // FFTPower of the band frame (:188), mean over the two channels
      const float *h = H + (size_t)b * KP;
      float acc = 0.f;
      const int hi = s_hi[b];
      for (int k = s_lo[b]; k < hi; ++k) { const float hv = __ldg(h + k); acc = fmaf(hv * hv, s_pw[k], acc); }
      band_energy[bt * nb + b] = 0.5f * acc / ((float)N * (float)N);
    }
    __syncwarp();
  }
}

// one thread per (stream, band, delay); arithmetic in the reference's order: c *= (1-m); prev *= m; c += prev; prev = c
__global__ void mb_scan_kernel(const float *__restrict__ raw, int B, int T, int nbD, float mem, float *__restrict__ state, float *__restrict__ curves) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nbD) return;
  const int s = i / nbD, e = i - s * nbD;
  const float keep = 1.0f - mem;
  float prev = state[i];
  for (int t = 0; t < T; ++t) {
    const long long o = ((long long)s * T + t) * nbD + e;
    const float c = __fadd_rn(__fmul_rn(raw[o], keep), __fmul_rn(prev, mem));
    curves[o] = c;
    prev = c;
  }
  state[i] = prev;
}

// one warp per frame
__global__ void __launch_bounds__(32 * MB_WARPS) mb_summary_kernel(const float *__restrict__ curves, const float *__restrict__ band_energy,
                                                                   long long BT, int nb, int D, float *__restrict__ hist_out,
                                                                   int32_t *__restrict__ band_cells, int32_t *__restrict__ raw_cell,
                                                                   float *__restrict__ raw_prob) {
  extern __shared__ float s_hist_all[];   // [MB_WARPS][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long bt = (long long)blockIdx.x * MB_WARPS + warp;
  if (bt >= BT) return;
  float *hist = s_hist_all + (size_t)warp * D;
  for (int d = lane; d < D; d += 32) hist[d] = 0.f;
  __syncwarp();
  for (int b = 0; b < nb; ++b) {
    const float *c = curves + (bt * nb + b) * D;
    float bv = -3.0e38f; int bi = 0x7fffffff;
    for (int d = lane; d < D; d += 32) { const float v = c[d]; if (v > bv) { bv = v; bi = d; } }
    warp_argmax(bv, bi);                                         // wipp::maxidx: first maximum (:184)
    if (lane == 0) {
      hist[bi] += band_energy[bt * nb + b];                      // _energyInDOA[idx] += _energies[bin], bands in order (:190)
      band_cells[bt * nb + b] = bi;
    }
    __syncwarp();
  }
  if (lane == 0) {
    float sum = 0.f, mx = hist[0]; int mi = 0;
    for (int d = 0; d < D; ++d) { const float v = hist[d]; sum += v; if (v > mx) { mx = v; mi = d; } }   // wipp::sum / maxidx (:222-223)
    raw_cell[bt] = mi;
    raw_prob[bt] = (sum != 0.f) ? hist[mi] / sum : sum;          // :225-228
  }
  __syncwarp();
  for (int d = lane; d < D; d += 32) hist_out[bt * D + d] = hist[d];
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused band / scan / summary for the shapes the reference produces (D <= 64 delays, bands of at most 16 bins): one CTA per
// stream walks the frames of the call in chunks of MF_TF.  The band kernels above stay as the general path (any D, any H).
//   A  stage   warp per frame, lanes over bins (coalesced): whitened cross-spectrum and weighted bin power of the bins the
//              bands touch go to shared memory, the floor power (bins < K/2) to HBM
//   B  energy  thread per (frame, band)
//   C  curves  warp per band, lane = delay (two delays per lane when D > 32): the band's phasors live in registers, the warp runs
//              over the chunk's frames in time order with the 0.4-memory recurrence in registers (no band_raw round trip
//              through HBM), stores the curve and takes the band arg-max with two REDUX instructions
//   D  summary thread per frame: energy histogram in band order, its sum / arg-max / prob (arithmetic order of mb_summary_kernel)
constexpr int MF_TF = 64, MF_WARPS = 8;

template <int BW>
__global__ void __launch_bounds__(32 * MF_WARPS) mb_fused_kernel(const float2 *__restrict__ spec, int B, int T, int N, const float *__restrict__ H,
                                                                 const int *__restrict__ band_lohi, int nb, int kmin, int kmax,
                                                                 const float2 *__restrict__ W, int D, float mem, float *__restrict__ state,
                                                                 float *__restrict__ curves, float *__restrict__ band_energy,
                                                                 float *__restrict__ floor_pow, float *__restrict__ hist_out,
                                                                 int32_t *__restrict__ band_cells, int32_t *__restrict__ raw_cell,
                                                                 float *__restrict__ raw_prob) {
  extern __shared__ __align__(16) unsigned char mf_smem[];
  const int KP = spec_pitch(N), K = N / 2 + 1, KB = kmax - kmin, GP = KB + BW;   // rows padded so a band's BW reads stay inside the row
  float2 *s_G = reinterpret_cast<float2 *>(mf_smem);                 // [MF_TF][GP]
  float *s_pw = reinterpret_cast<float *>(s_G + (size_t)MF_TF * GP);  // [MF_TF][KB]
  float *s_e = s_pw + (size_t)MF_TF * KB;                             // [MF_TF][nb]
  int *s_cell = reinterpret_cast<int *>(s_e + (size_t)MF_TF * nb);    // [MF_TF][nb]
  float *s_hist = reinterpret_cast<float *>(s_cell + (size_t)MF_TF * nb);   // [MF_TF][D]
  int *s_lo = reinterpret_cast<int *>(s_hist + (size_t)MF_TF * D), *s_hi = s_lo + nb;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int b = tid; b < nb; b += blockDim.x) { s_lo[b] = band_lohi[2 * b]; s_hi[b] = band_lohi[2 * b + 1]; }
  for (int i = tid; i < MF_TF * BW; i += blockDim.x) s_G[(size_t)(i / BW) * GP + KB + (i % BW)] = make_float2(0.f, 0.f);   // the row padding
  __syncthreads();
  const int KF = K / 2;                                         // FFTPower(frames, K): the first K/2 bins of the CCS buffer (:130)
  const float nF = (float)(K - 2), keep = 1.0f - mem;
  const int kend = max(KF, kmax);
  for (int s = blockIdx.x; s < B; s += gridDim.x) {
    for (int t0 = 0; t0 < T; t0 += MF_TF) {
      const int nt = min(MF_TF, T - t0);
      const long long bt0 = (long long)s * T + t0;
      // ---- A
      for (int f = warp; f < nt; f += MF_WARPS) {
        const float2 *L = spec + (bt0 + f) * 2 * KP, *R = L + KP;
        float fl = 0.f;
        for (int k0 = lane; k0 < kend; k0 += 128) {   // four bins of both channels in flight per lane before the first use
          float2 l[4], r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = k0 + 32 * u;
            l[u] = k < kend ? L[k] : make_float2(0.f, 0.f);
            r[u] = k < kend ? R[k] : make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = k0 + 32 * u;
            const float pq = l[u].x * l[u].x + l[u].y * l[u].y + r[u].x * r[u].x + r[u].y * r[u].y;
            if (k < KF) fl += ((k == 0 || k == KF - 1) ? 1.f : 2.f) * pq;
            if (k >= kmin && k < kmax) {
              s_G[(size_t)f * GP + k - kmin] = whiten(cmulc(l[u], r[u]));
              s_pw[(size_t)f * KB + k - kmin] = ((k == 0 || k == K - 1) ? 1.f : 2.f) * pq;
            }
          }
        }
        fl = warp_sum(fl);
        if (lane == 0) floor_pow[bt0 + f] = 0.5f * fl / (nF * nF);       // mean over the two channels
      }
      __syncthreads();
      // ---- B: FFTPower of the band frame (:188), mean over the two channels
      for (int i = tid; i < nt * nb; i += blockDim.x) {
        const int f = i / nb, b = i - f * nb;
        const float *h = H + (size_t)b * KP;
        float acc = 0.f;
        for (int k = s_lo[b]; k < s_hi[b]; ++k) { const float hv = __ldg(h + k); acc = fmaf(hv * hv, s_pw[(size_t)f * KB + k - kmin], acc); }
        const float e = 0.5f * acc / ((float)N * (float)N);
        s_e[f * nb + b] = e;
        band_energy[(bt0 + f) * nb + b] = e;
      }
      // ---- C
      for (int b = warp; b < nb; b += MF_WARPS) {
        const int lo = s_lo[b], d0 = lane, d1 = lane + 32;
        const bool on0 = d0 < D, on1 = d1 < D;
        float2 w0[BW], w1[BW];
#pragma unroll
        for (int i = 0; i < BW; ++i) {
          const int k = lo + i;
          const bool use = k < s_hi[b] && __ldg(H + (size_t)b * KP + k) != 0.f;   // bins outside the band's support add exactly 0
          w0[i] = (use && on0) ? __ldg(W + (size_t)d0 * KP + k) : make_float2(0.f, 0.f);
          w1[i] = (use && on1) ? __ldg(W + (size_t)d1 * KP + k) : make_float2(0.f, 0.f);
        }
        float *st0 = state + ((size_t)s * nb + b) * D;
        float prev0 = on0 ? st0[d0] : 0.f, prev1 = on1 ? st0[d1] : 0.f;
        const float2 *g = s_G + (lo - kmin);
        for (int f = 0; f < nt; ++f) {
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int i = 0; i < BW; ++i) {
            const float2 gv = g[(size_t)f * GP + i];
            a0 = fmaf(gv.x, w0[i].x, fmaf(-gv.y, w0[i].y, a0));
            a1 = fmaf(gv.x, w1[i].x, fmaf(-gv.y, w1[i].y, a1));
          }
          // c *= (1-m); prev *= m; c += prev; prev = c   (:180-183)
          const float c0 = __fadd_rn(__fmul_rn(a0, keep), __fmul_rn(prev0, mem)), c1 = __fadd_rn(__fmul_rn(a1, keep), __fmul_rn(prev1, mem));
          prev0 = c0; prev1 = c1;
          float *dst = curves + ((bt0 + f) * nb + b) * D;
          if (on0) dst[d0] = c0;
          if (on1) dst[d1] = c1;
          // wipp::maxidx: first maximum (:184)
          float bv = on0 ? c0 : -3.0e38f; int bi = on0 ? d0 : 0x7fffffff;
          if (on1 && c1 > bv) { bv = c1; bi = d1; }
          const unsigned key = float_order_key(bv + 0.f);
          const unsigned kbest = __reduce_max_sync(0xffffffffu, key);
          const unsigned ibest = __reduce_min_sync(0xffffffffu, key == kbest ? (unsigned)bi : 0x7fffffffu);
          if (lane == 0) { s_cell[f * nb + b] = (int)ibest; band_cells[(bt0 + f) * nb + b] = (int)ibest; }
        }
        if (on0) st0[d0] = prev0;
        if (on1) st0[d1] = prev1;
      }
      __syncthreads();
      // ---- D
      for (int f = tid; f < nt; f += blockDim.x) {
        float *hist = s_hist + (size_t)f * D;
        for (int d = 0; d < D; ++d) hist[d] = 0.f;
        for (int b = 0; b < nb; ++b) hist[s_cell[f * nb + b]] += s_e[f * nb + b];   // _energyInDOA[idx] += _energies[bin], bands in order (:190)
        float sum = 0.f, mx = hist[0]; int mi = 0;
        for (int d = 0; d < D; ++d) { const float v = hist[d]; sum += v; if (v > mx) { mx = v; mi = d; } }   // wipp::sum / maxidx (:222-223)
        raw_cell[bt0 + f] = mi;
        raw_prob[bt0 + f] = (sum != 0.f) ? hist[mi] / sum : sum;          // :225-228
      }
      __syncthreads();
      for (int i = tid; i < nt * D; i += blockDim.x) hist_out[bt0 * D + i] = s_hist[i];
      __syncthreads();
    }
  }
}

bool k_mb_fused_supported(int D, int max_band_width, int kmin, int kmax, int nb) {
  const int BW = max_band_width <= 8 ? 8 : 16;
  const size_t smem = (size_t)MF_TF * ((size_t)(kmax - kmin + BW) * 8 + (size_t)(kmax - kmin) * 4 + (size_t)nb * 8 + (size_t)D * 4) + (size_t)nb * 8;
  return D >= 1 && D <= 64 && max_band_width >= 1 && max_band_width <= 16 && kmax > kmin && smem <= 200 * 1024;
}
int k_mb_fused(const float2 *spec, int B, int T, int N, const float *H, const int *band_lohi, int nb, int max_band_width, int kmin, int kmax,
               const float2 *W, int D, float mem, float *state, float *curves, float *band_energy, float *floor_pow, float *hist,
               int32_t *band_cells, int32_t *raw_cell, float *raw_prob, cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  const int BW = max_band_width <= 8 ? 8 : 16;
  const size_t smem = (size_t)MF_TF * ((size_t)(kmax - kmin + BW) * 8 + (size_t)(kmax - kmin) * 4 + (size_t)nb * 8 + (size_t)D * 4) + (size_t)nb * 8;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = B < sms * 64 ? B : sms * 64;   // one CTA per stream: the hardware scheduler balances the tail
  if (BW == 8) {
    cudaFuncSetAttribute(mb_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mb_fused_kernel<8><<<grid, 32 * MF_WARPS, smem, st>>>(spec, B, T, N, H, band_lohi, nb, kmin, kmax, W, D, mem, state, curves, band_energy, floor_pow,
                                                          hist, band_cells, raw_cell, raw_prob);
  } else {
    cudaFuncSetAttribute(mb_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mb_fused_kernel<16><<<grid, 32 * MF_WARPS, smem, st>>>(spec, B, T, N, H, band_lohi, nb, kmin, kmax, W, D, mem, state, curves, band_energy, floor_pow,
                                                           hist, band_cells, raw_cell, raw_prob);
  }
  MCAG_CHECK_LAUNCH();
  return 0;
}

struct MbGateState { double acc; double floor; int samples; int estimated; };   // same layout as the processors' GateState

// one warp per stream: the lanes fetch the inputs of 32 frames at once (coalesced), then every lane replays the sequential state
// machine over those frames from shuffled values (uniform, redundant) and lane i stores the results of frame i.  One thread per
// stream walking the frames was latency-bound on its dependent loads (0.14 ms for 2048 streams x 125 frames).
__global__ void __launch_bounds__(128) mb_gate_kernel(const float *__restrict__ floor_pow, const float *__restrict__ chan_pow,
                                                      const int32_t *__restrict__ raw_cell, const float *__restrict__ raw_prob, int B, int T, int N,
                                                      int use_floor, float margin_db, int needed, MbGateState *__restrict__ gs,
                                                      int32_t *__restrict__ cell_state, float *__restrict__ power_out,
                                                      unsigned char *__restrict__ active, int32_t *__restrict__ cells, float *__restrict__ prob) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= B) return;
  MbGateState g = gs[s];
  int32_t cur = cell_state[s];
  const int K = N / 2 + 1;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int nt = min(32, T - t0);
    const long long bt = (long long)s * T + t0 + lane;
    float fp = 0.f, cp0 = 0.f, cp1 = 0.f, rp = 0.f; int32_t rc = 0;
    if (lane < nt) { fp = floor_pow[bt]; cp0 = chan_pow[bt * 2]; cp1 = chan_pow[bt * 2 + 1]; rc = raw_cell[bt]; rp = raw_prob[bt]; }
    double my_power = 0.0; bool my_on = false; int32_t my_cell = 0;
    for (int i = 0; i < nt; ++i) {
      const float fpi = __shfl_sync(0xffffffffu, fp, i), c0 = __shfl_sync(0xffffffffu, cp0, i), c1 = __shfl_sync(0xffffffffu, cp1, i);
      const int32_t rci = __shfl_sync(0xffffffffu, rc, i);
      double power;
      if (!g.estimated) {                                          // setPowerFloor (:127-143)
        g.acc += (double)fpi * (double)(2 * K - 2);
        g.samples += 2 * K - 2;
        if (g.samples >= needed) {
          g.estimated = 1;
          g.acc /= (double)g.samples;
          g.acc = 10.0 * log10(g.acc) + (double)margin_db;
        }
        g.floor = g.acc;
        power = g.floor;
      } else {
        power = 0.5 * ((double)c0 + (double)c1);                   // FFTPower(frames, N+2) (:221)
      }
      const bool on = (power > g.floor) || !use_floor;             // :225
      if (on) cur = rci;
      if (lane == i) { my_power = power; my_on = on; my_cell = cur; }
    }
    if (lane < nt) {
      cells[bt] = my_cell;
      prob[bt] = my_on ? rp : -100000.f;                           // :253
      power_out[bt] = (float)my_power;
      active[bt] = my_on ? 1 : 0;
    }
  }
  if (lane == 0) { gs[s] = g; cell_state[s] = cur; }
}

static int mb_grid(long long BT) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = (BT + MB_WARPS - 1) / MB_WARPS, cap = (long long)sms * 4;
  return (int)(want < cap ? want : cap);
}

int k_mb_band(const float2 *spec, long long BT, int N, const float *H, int nb, const float2 *W, int D, float *band_raw, float *band_energy,
              float *floor_pow, cudaStream_t st) {
  if (BT <= 0) return 0;
  const int K = N / 2 + 1;
  size_t smem = (sizeof(float2) + sizeof(float)) * MB_WARPS * K + sizeof(int) * 2 * nb;
  cudaFuncSetAttribute(mb_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mb_band_kernel<<<mb_grid(BT), 32 * MB_WARPS, smem, st>>>(spec, BT, N, H, nb, W, D, band_raw, band_energy, floor_pow);
  MCAG_CHECK_LAUNCH();
  return 0;
}
int k_mb_scan(const float *raw, int B, int T, int nb, int D, float mem, float *state, float *curves, cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  const int n = B * nb * D;
  mb_scan_kernel<<<(n + 127) / 128, 128, 0, st>>>(raw, B, T, nb * D, mem, state, curves);
  MCAG_CHECK_LAUNCH();
  return 0;
}
int k_mb_summary(const float *curves, const float *band_energy, long long BT, int nb, int D, float *hist, int32_t *band_cells, int32_t *raw_cell,
                 float *raw_prob, cudaStream_t st) {
  if (BT <= 0) return 0;
  mb_summary_kernel<<<(unsigned)((BT + MB_WARPS - 1) / MB_WARPS), 32 * MB_WARPS, sizeof(float) * MB_WARPS * D, st>>>(curves, band_energy, BT, nb, D, hist,
                                                                                                                    band_cells, raw_cell, raw_prob);
  MCAG_CHECK_LAUNCH();
  return 0;
}
int k_mb_gate(const float *floor_pow, const float *chan_pow, const int32_t *raw_cell, const float *raw_prob, int B, int T, int N, int use_floor,
              float margin_db, int needed, void *gate_state, int32_t *cell_state, float *power_out, unsigned char *active, int32_t *cells, float *prob,
              cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  mb_gate_kernel<<<(B + 3) / 4, 128, 0, st>>>(floor_pow, chan_pow, raw_cell, raw_prob, B, T, N, use_floor, margin_db, needed,
                                               static_cast<MbGateState *>(gate_state), cell_state, power_out, active, cells, prob);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

// K2 gcc_phat: replaces dsp::GeneralisedCrossCorrelation as driven by
// SteeringBeamforming::computeCorrelations (SteeringBeamforming.cpp:104-130) and
// FreqGCCBinauralLocalisation::processParametrisation (BinauralLocalisation.cpp:438-444).
//   k_tdoa_lags : integer-lag mode (BASELINE config 2) — PHAT cross-spectrum, inverse real FFT in shared memory,
//                 lag-window extraction and first-maximum argmax fused in one kernel.
//   k_gcc_tau   : fractional-delay (tau grid) mode, the reference's own semantics, as a register-tiled contraction
//                 of PHAT cross-spectra against steering phasors generated on the fly from fixed-point phase ramps.
#include "fft.cuh"
#include "kernels.h"
#include "tdoa_warp.cuh"

#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace mcag {

// ---------------------------------------------------------------------------------------------------
// integer-lag GCC-PHAT.  The M whitened spectra of one frame sit in shared memory (s_U); each group of N/16
// threads runs one pair at a time:
//   Z[k] = E[k] + i O[k] from G = U_i conj(U_j)  ->  N/2-point inverse complex FFT  ->  r[l]/2 in packed form
//   S[l] = r[l]/2 + (Re G[0] + (-1)^l Re G[N/2])/2   (one-sided sum of oracle/CONVENTIONS.md C5)
// followed by the lag-window extraction and a first-maximum argmax (warp shuffles).
// Two kernels share this pair phase:
//   tdoa_kernel       spectra come from HBM (kernel-level entry mcag_k_tdoa_lags)
//   stft_tdoa_kernel  the fused STFT -> GCC-PHAT pipeline of the TDOA processor: frames are read straight from the
//                     sample rows, windowed and transformed in the same CTA; spectra only leave the SM when asked for.
// ---------------------------------------------------------------------------------------------------
template <int N, int G> struct TdoaSmem {
  static constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), WPF = (TPF + 31) / 32;
};

template <int N, int G>
__device__ __forceinline__ void pair_table(unsigned char *s_pair, int M, int P, int tid, int NT) {
  for (int p = tid; p < P; p += NT) {   // pair p -> (i, j), i < j lexicographic (SteeringBeamforming.cpp:63-65)
    int i = 0, rem = p;
    while (rem >= M - 1 - i) { rem -= M - 1 - i; ++i; }
    s_pair[2 * p] = (unsigned char)i;
    s_pair[2 * p + 1] = (unsigned char)(i + 1 + rem);
  }
}

// E/O twiddles i * conj(W_N^k) of the packed inverse real transform for the 4 bins k = j + r*TPF (r < 4) this thread builds: the
// same for every pair and frame, so they live in registers
template <int N> __device__ __forceinline__ void load_eo_twiddles(const float2 *s_tw, int j, float2 (&wk)[4]) {
  constexpr int NC = N / 2, TPF = NC / 8;
#pragma unroll
  for (int r = 0; r < 4; ++r) {   // stored times i: the build then forms E + iO and conj(E - iO) with packed adds only
    const float2 w = tw_lookup<true>(s_tw, j + r * TPF, NC);
    wk[r] = make_float2(-w.y, w.x);
  }
}

// all P pairs of the frame whose whitened spectra are in s_U; every thread of the CTA calls it (uniform trip count).
// The lag window only needs the lowest and highest NC/R_last outputs of the inverse transform, so its last pass is
// output-pruned (fft_inv_last_pruned): nothing is stored, the window values go from registers into the arg-max.
template <int N, int G>
__device__ __forceinline__ void tdoa_pairs(const float2 *s_U, const float2 (&wk)[4], const float2 *s_twp, fft_buf_t buf, const unsigned char *s_pair,
                                           float *s_bv, int *s_bi, int P, int max_lag, int g, int j, float *__restrict__ curves_ft,
                                           int32_t *__restrict__ lags_ft, float *__restrict__ peaks_ft, const int *s_pout = nullptr) {
  constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N);
  using PL = FftPlan<NC>;
  constexpr int RL = PL::R[PL::NP - 1], NBL = 8 / RL, NSL = NC / RL;
  const int L = 2 * max_lag + 1;
  const int hl = max_lag >> 1, hh = (max_lag + 1) >> 1;   // complex outputs 0..hl hold lags >= 0, NC-hh..NC-1 the negative ones
  const bool pruned = hl < NSL && hh <= NSL;
  const int rounds = (P + G - 1) / G;
  for (int it = 0; it < rounds; ++it) {   // every thread runs every round, stores are predicated
    const int p = it * G + g;
    const bool live = p < P;
    const int pc = live ? p : P - 1;
    const float2 *Ui = s_U + (size_t)s_pair[2 * pc] * KP, *Uj = s_U + (size_t)s_pair[2 * pc + 1] * KP;
    // Mirrored build of the transform input: the bins k and NC-k need the same two cross-spectrum values, Z[k] = E + iO and
    // Z[NC-k] = conj(E) + i conj(O), so a thread forms each bin pair ONCE (k = j + r*TPF, r < 4: half the loads and half the
    // arithmetic of building its 8 first-pass points itself).  Z[k] is its own first-pass point r; Z[NC-k] is point 7-r of the
    // partner thread TPF-j, handed over through the partner's private pass-0 region of the transform buffer (slots 8p..8p+7,
    // which only p itself overwrites afterwards).  Thread TPF/2 is its own partner; thread 0 pairs (r*TPF, (8-r)*TPF) inside
    // itself and also forms the self-paired bin NC/2.
    float2 v[8];
    const int pj = (j == 0) ? 0 : TPF - j;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = j + r * TPF;
      float2 gk = cmulc(Ui[k], Uj[k]), gc = cmulc(Uj[NC - k], Ui[NC - k]);   // G[k] and conj(G[NC-k])
      if (k == 0) { gk.y = 0.f; gc.y = 0.f; }
      const float2 e = cscale(cadd(gk, gc), 0.5f), d = cscale(csub(gk, gc), 0.5f);
      const float2 o = cmul(d, wk[r]);               // i * O[k]
      v[r] = cadd(e, o);                             // Z[k] = E + iO
      const float2 zn = cconj(csub(e, o));           // Z[NC-k] = conj(E - iO)
      if (j != 0) sts64(buf ^ (8u * (uint32_t)fft_pad(8 * pj + 7 - r)), zn);
      else if (r != 0) sts64(buf ^ (8u * (uint32_t)fft_pad(8 - r)), zn);
    }
    if (j == 0) sts64(buf ^ (8u * (uint32_t)fft_pad(4)), cmulc(Uj[NC / 2], Ui[NC / 2]));   // Z[NC/2] = conj(G[NC/2])
    group_sync<TPF>(g);
#pragma unroll
    for (int r = 4; r < 8; ++r) v[r] = lds64(buf ^ (8u * (uint32_t)fft_pad(8 * j + r)));
    const float g0 = Ui[0].x * Uj[0].x, gny = Ui[NC].x * Uj[NC].x;   // both spectra are real at DC / Nyquist
    const float b_even = 0.5f * (g0 + gny), b_odd = 0.5f * (g0 - gny);
    const int po = s_pout ? s_pout[pc] : p;   // where the pair's results go (channel-tiled launches: the global pair index)
    float *cdst = (curves_ft && live) ? curves_ft + (size_t)po * L : nullptr;
    float best = -3.0e38f; int besti = 0x7fffffff;
    auto cand = [&](int l, float val) {   // first maximum: the lowest window index wins ties
      const float sv = val + ((l & 1) ? b_odd : b_even);
      const int c = l + max_lag;
      if (cdst) cdst[c] = sv;
      if (sv > best || (sv == best && c < besti)) { best = sv; besti = c; }
    };
    if (pruned) {
      fft_run_head<NC, true>(v, buf, s_twp, j, g);
      auto block = [&](auto bc) {
        constexpr int B_ = decltype(bc)::value;
        const int jj = j + B_ * (NC / 8);
        const bool lo = jj <= hl, hi = jj >= NSL - hh;
        if (lo || hi) {
          float2 y0, y1;
          fft_inv_last_pruned<NC, B_>(buf, s_twp, j, lo, hi, y0, y1);
          if (lo) {
            cand(2 * jj, y0.x);
            if (2 * jj + 1 <= max_lag) cand(2 * jj + 1, y0.y);
          }
          if (hi) {
            const int l = 2 * (jj + (RL - 1) * NSL) - N;   // <= -2
            if (l >= -max_lag) cand(l, y1.x);
            cand(l + 1, y1.y);
          }
        }
      };
      block(std::integral_constant<int, 0>{});
      if constexpr (NBL > 1) block(std::integral_constant<int, 1>{});
      if constexpr (NBL > 2) { block(std::integral_constant<int, 2>{}); block(std::integral_constant<int, 3>{}); }
    } else {
      fft_run<NC, true>(v, buf, s_twp, j, g);
      for (int c = j; c < L; c += TPF) {
        const int l = c - max_lag;
        const int li = (l + N) & (N - 1);
        const float2 z = fft_buf_get(buf, li >> 1);
        cand(l, (li & 1) ? z.y : z.x);
      }
    }
    // first-maximum argmax across the group
    if constexpr (TPF >= 32) {
      // two REDUX instructions per warp instead of a 5-step shuffle tree: maximum of the order-preserving integer image of
      // the value (x + 0 folds -0 into +0 so that equal floats have equal keys), then the lowest window index that attains it
      constexpr int WPF = TPF / 32;
      const unsigned key = float_order_key(best + 0.f);
      const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
      const unsigned imin = __reduce_min_sync(0xffffffffu, key == kmax ? (unsigned)besti : 0x7fffffffu);
      unsigned *s_bk = reinterpret_cast<unsigned *>(s_bv) + (it & 1) * (G * WPF);   // double-buffered by round: one sync per round
      int *s_bx = s_bi + (it & 1) * (G * WPF);
      if ((threadIdx.x & 31) == 0) { s_bk[g * WPF + (j >> 5)] = kmax; s_bx[g * WPF + (j >> 5)] = (int)imin; }
      group_sync<TPF>(g);   // also: every thread of the group is done reading buf
      if (j == 0 && live) {
        unsigned bk = s_bk[g * WPF]; int bi = s_bx[g * WPF];
        for (int w = 1; w < WPF; ++w) { unsigned ok = s_bk[g * WPF + w]; int oi = s_bx[g * WPF + w]; if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; } }
        lags_ft[po] = bi - max_lag;
        if (peaks_ft) peaks_ft[po] = float_from_order_key(bk);
      }
    } else {
      const unsigned gmask = ((1u << (TPF & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(TPF - 1));
#pragma unroll
      for (int o = TPF / 2; o > 0; o >>= 1) {   // stays inside the aligned TPF-lane segment
        float ov = __shfl_xor_sync(gmask, best, o);
        int oi = __shfl_xor_sync(gmask, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (j == 0 && live) {
        lags_ft[po] = besti - max_lag;
        if (peaks_ft) peaks_ft[po] = best;
      }
      group_sync<TPF>(g);
    }
  }
}

// Channel tiles: the kernel stages the whitened spectra of channels [ci0, ci0 + cin) and, when cjn > 0, [cj0, cj0 + cjn) and runs the
// pairs (i < j) inside the first tile (cjn = 0) or every (i in the first, j in the second) pair; results land at the pair's global
// index i (2M - i - 1)/2 + j - i - 1.  One launch with cin = M covers arrays whose spectra fit in shared memory; larger arrays (64
// microphones at N = 1024 need 263 KB) are covered tile pair by tile pair.
template <int N, int G>
__global__ void __launch_bounds__(G *(N / 16)) tdoa_kernel(const float2 *__restrict__ spec, int T, int M, int max_lag,
                                                            const float2 *__restrict__ tw_g, float *__restrict__ curves,
                                                            int32_t *__restrict__ lags, float *__restrict__ peaks, int ci0, int cin, int cj0, int cjn) {
  constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), NT = G * TPF, WPF = (TPF + 31) / 32;
  const int P = M * (M - 1) / 2, L = 2 * max_lag + 1;
  const int CS = cin + cjn, PL = cjn ? cin * cjn : cin * (cin - 1) / 2;   // staged channels, pairs of this launch
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = fft_align_smem(smem_raw, 8 * NC);
  float2 *s_buf = reinterpret_cast<float2 *>(smem);                   // G * fft_buf_len(NC), each buffer aligned to its size
  float2 *s_U = s_buf + G * fft_buf_len(NC);                          // CS * KP
  float2 *s_tw = s_U + (size_t)CS * KP;                               // fft_table_len(N): tw[NC] then twp
  float2 *s_twp = s_tw + NC;
  float *s_bv = reinterpret_cast<float *>(s_tw + fft_table_len(N));   // 2 * G * WPF (double-buffered by pair round)
  int *s_bi = reinterpret_cast<int *>(s_bv + 2 * G * WPF);            // 2 * G * WPF
  int *s_pout = s_bi + 2 * G * WPF;                                   // PL global pair indices
  unsigned char *s_pair = reinterpret_cast<unsigned char *>(s_pout + PL);   // 2 * PL staged-slot pairs

  const int tid = threadIdx.x, t = blockIdx.x, b = blockIdx.y;
  const float2 *src = spec + ((long long)b * T + t) * M * KP;
  for (int i = tid; i < CS * KP; i += NT) {
    const int slot = i / KP, k = i - slot * KP;
    const int ch = slot < cin ? ci0 + slot : cj0 + (slot - cin);
    s_U[i] = (k <= NC) ? whiten(src[(size_t)ch * KP + k]) : make_float2(0.f, 0.f);
  }
  fft_load_tables<N>(s_tw, tw_g, tid, NT);
  for (int q = tid; q < PL; q += NT) {   // local pair q -> staged slots and global pair index, lexicographic like SteeringBeamforming.cpp:63-65
    int a, c;
    if (cjn) { a = q / cjn; c = cin + (q - a * cjn); }
    else { a = 0; int rem = q; while (rem >= cin - 1 - a) { rem -= cin - 1 - a; ++a; } c = a + 1 + rem; }
    const int gi = ci0 + a, gj = c < cin ? ci0 + c : cj0 + (c - cin);
    s_pair[2 * q] = (unsigned char)a; s_pair[2 * q + 1] = (unsigned char)c;
    s_pout[q] = gi * (2 * M - gi - 1) / 2 + (gj - gi - 1);
  }
  __syncthreads();
  const int g = tid / TPF, j = tid % TPF;
  const long long ft = (long long)b * T + t;
  float2 wk[4];
  load_eo_twiddles<N>(s_tw, j, wk);
  tdoa_pairs<N, G>(s_U, wk, s_twp, smem_u32(s_buf + g * fft_buf_len(NC)), s_pair, s_bv, s_bi, PL, max_lag, g, j,
                   curves ? curves + ft * P * L : nullptr, lags + ft * P, peaks ? peaks + ft * P : nullptr, s_pout);
}

// Analysis phase of the fused kernels: the M windows of frame (b, t) are read straight from the sample rows (coalesced 8-byte loads, the
// overlapping half is an L2 hit of the neighbouring frame), windowed while packing into the N/2-point complex FFT, post-processed to the
// one-sided spectrum (optionally written, always its Parseval power) and whitened into s_U.  Channels m = g, g + G, ...; every thread
// of the CTA calls it, the caller synchronises the CTA afterwards.
template <int N, int G>
__device__ __forceinline__ void analysis_phase(const float *__restrict__ x, long long row_pitch, int M, int hop, bool vec_ok, long long ft, int b, int t,
                                               const float2 *s_w, const float2 *s_tw, const float2 *s_twp, fft_buf_t buf, float2 *s_U, float *s_bv,
                                               float2 *__restrict__ spec, float *__restrict__ chan_pow, int g, int j, float2 *s_Uf = nullptr) {
  constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), WPF = (TPF + 31) / 32;
  const int tid = threadIdx.x;
  // the samples of the NEXT channel of this group are requested before the current one is transformed: one global-memory round trip per
  // frame is exposed instead of one per channel (ncu r2c_cfg2_warp: 8 % of the samples sat on the first use of these loads)
  auto load_channel = [&](int m, float2 (&dst)[8]) {
    const float *src = x + ((long long)b * M + m) * row_pitch + (long long)t * hop;
    if (vec_ok) {
      const float2 *s2 = reinterpret_cast<const float2 *>(src);
#pragma unroll
      for (int r = 0; r < 8; ++r) dst[r] = __ldg(s2 + j + r * TPF);
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) { const int n = j + r * TPF; dst[r] = make_float2(src[2 * n], src[2 * n + 1]); }
    }
  };
  float2 nxt[8];
  if (g < M) load_channel(g, nxt);
  for (int m = g; m < M; m += G) {
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = nxt[r];
    if (m + G < M) load_channel(m + G, nxt);
#pragma unroll
    for (int r = 0; r < 8; ++r) { const float2 w = s_w[j + r * TPF]; v[r].x *= w.x; v[r].y *= w.y; }
    fft_run<NC, false>(v, buf, s_twp, j, g);
    // real post-processing: X[k] = E + W^k O, X[NC-k] = conj(E - W^k O)
    float2 *U = s_U + (size_t)m * KP;
    float2 *out = spec ? spec + ((ft * M + m) * KP) : nullptr;
    float pw = 0.f;
    for (int k = j; k <= NC / 2; k += TPF) {
      float2 zk = fft_buf_get(buf, k), zn = fft_buf_get(buf, (NC - k) & (NC - 1));
      float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));   // -i/2 (zk - conj zn)
      float2 w = s_tw[k];   // k <= NC/2: first half of the table, W_N^k itself
      float2 wo = cmul(w, o);
      float2 xk = cadd(e, wo), xn = cconj(csub(e, wo));
      if (k == 0) { xk.y = 0.f; xn.y = 0.f; }
      if (out) { out[k] = xk; out[NC - k] = xn; }
      const float wk = (k == 0) ? 1.f : 2.f;
      pw += wk * (xk.x * xk.x + xk.y * xk.y);
      if (k != NC - k) pw += wk * (xn.x * xn.x + xn.y * xn.y);
      const float2 uk = whiten(xk), un = whiten(xn);
      U[k] = uk;
      U[NC - k] = un;
      if (s_Uf) {   // N = 1024 warp lag phase: compact copy of the bins 32 q + 16 (k and NC - k fall into this class together)
        if ((k & 31) == 16) { s_Uf[m * 16 + (k >> 5)] = uk; s_Uf[m * 16 + ((NC - k) >> 5)] = un; }
      }
    }
    if (j == 0 && out) out[NC + 1] = make_float2(0.f, 0.f);   // pad bin
    if (chan_pow) {   // Parseval power of the windowed frame, fixed reduction order
      if constexpr (TPF >= 32) {
        pw = warp_sum(pw);
        if ((tid & 31) == 0) s_bv[g * WPF + (j >> 5)] = pw;
        group_sync<TPF>(g);
        if (j == 0) {
          float sacc = 0.f;
          for (int i = 0; i < WPF; ++i) sacc += s_bv[g * WPF + i];
          chan_pow[ft * M + m] = sacc / ((float)N * (float)N);
        }
      } else {
        // several transforms share a warp and a neighbour group may be idle this round: shuffle only among this group's lanes
        const unsigned gmask = (TPF >= 32) ? 0xffffffffu : (((1u << (TPF & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(TPF - 1)));
#pragma unroll
        for (int o2 = TPF / 2; o2 > 0; o2 >>= 1) pw += __shfl_xor_sync(gmask, pw, o2);
        if (j == 0) chan_pow[ft * M + m] = pw / ((float)N * (float)N);
      }
    }
    group_sync<TPF>(g);
  }
}

// Fused STFT -> GCC-PHAT -> lag argmax.  Persistent CTAs walk the (stream, frame) list; per frame a CTA
//   1. reads the M windows of N samples straight from the sample rows (coalesced 8-byte loads, the overlapping half is an
//      L2 hit of the neighbouring frame), applies the analysis window while packing into the N/2-point complex FFT,
//   2. post-processes to the one-sided spectrum, optionally writes it (and always its Parseval power), whitens it into s_U,
//   3. runs the pair phase above.
// Compulsory HBM traffic per frame: 4*M*hop bytes in, 4*P bytes out (+ 8*M*(N/2+2) when spectra are requested).
template <int N, int G>
__global__ void __launch_bounds__(G *(N / 16), 768 / (G * (N / 16))) stft_tdoa_kernel(const float *__restrict__ x, long long row_pitch, int B, int T, int M, int hop,
                                                                 int max_lag, const float *__restrict__ win, const float2 *__restrict__ tw_g,
                                                                 float2 *__restrict__ spec, float *__restrict__ chan_pow,
                                                                 float *__restrict__ curves, int32_t *__restrict__ lags) {
  constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), NT = G * TPF, WPF = (TPF + 31) / 32;
  const int P = M * (M - 1) / 2, L = 2 * max_lag + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = fft_align_smem(smem_raw, 8 * NC);
  float2 *s_buf = reinterpret_cast<float2 *>(smem);                   // G * NC, each buffer aligned to its size
  float2 *s_U = s_buf + G * fft_buf_len(NC);                          // M * KP
  float2 *s_tw = s_U + (size_t)M * KP;                                // fft_table_len(N)
  float2 *s_twp = s_tw + NC;
  float2 *s_w = s_tw + fft_table_len(N);                              // NC (window, pairs of samples)
  float *s_bv = reinterpret_cast<float *>(s_w + NC);                  // 2 * G * WPF (double-buffered by pair round)
  int *s_bi = reinterpret_cast<int *>(s_bv + 2 * G * WPF);            // 2 * G * WPF
  unsigned char *s_pair = reinterpret_cast<unsigned char *>(s_bi + 2 * G * WPF);   // 2 * P

  const int tid = threadIdx.x;
  fft_load_tables<N>(s_tw, tw_g, tid, NT);
  for (int i = tid; i < NC; i += NT) s_w[i] = make_float2(win[2 * i], win[2 * i + 1]);
  pair_table<N, G>(s_pair, M, P, tid, NT);
  __syncthreads();
  const int g = tid / TPF, j = tid % TPF;
  const fft_buf_t buf = smem_u32(s_buf + g * fft_buf_len(NC));
  float2 wk[4];
  load_eo_twiddles<N>(s_tw, j, wk);
  const bool vec_ok = ((row_pitch & 1) == 0) && ((hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
  const long long nframes = (long long)B * T;

  for (long long ft = blockIdx.x; ft < nframes; ft += gridDim.x) {
    const int b = (int)(ft / T), t = (int)(ft - (long long)b * T);
    analysis_phase<N, G>(x, row_pitch, M, hop, vec_ok, ft, b, t, s_w, s_tw, s_twp, buf, s_U, s_bv, spec, chan_pow, g, j);
    __syncthreads();
    tdoa_pairs<N, G>(s_U, wk, s_twp, buf, s_pair, s_bv, s_bi, P, max_lag, g, j,
                     curves ? curves + ft * P * L : nullptr, lags + ft * P, nullptr);
    __syncthreads();
  }
}

// The same fused pipeline for N = 1024 with the warp-synchronous lag phase of tdoa_warp.cuh: after the analysis phase every warp takes
// two pairs at a time (one per half-warp) and runs their decimated inverse transforms in registers.  256 threads: 4 groups of 64 for
// the forward transforms, 8 warps for the pairs; 2 CTAs per SM (the 32-point register transform needs ~120 registers).
template <int LCAP>
__global__ void __launch_bounds__(256, 2) stft_tdoa_warp_kernel(const float *__restrict__ x, long long row_pitch, int B, int T, int M, int hop, int max_lag,
                                                                const float *__restrict__ win, const float2 *__restrict__ tw_g,
                                                                const float2 *__restrict__ lag_tab, float2 *__restrict__ spec,
                                                                float *__restrict__ chan_pow, float *__restrict__ curves, int32_t *__restrict__ lags) {
  constexpr int N = 1024, G = 4, NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), NT = G * TPF, WPF = TPF / 32, NW = NT / 32;
  const int P = M * (M - 1) / 2, L = 2 * max_lag + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = fft_align_smem(smem_raw, 8 * NC);
  float2 *s_buf = reinterpret_cast<float2 *>(smem);                   // G * NC, each buffer aligned to its size
  float *s_scr = reinterpret_cast<float *>(s_buf + G * fft_buf_len(NC));   // NW warps x 2 pairs x 16 rows x 32 floats (transpose scratch)
  float2 *s_U = reinterpret_cast<float2 *>(s_scr + NW * 1024);        // M * KP
  float2 *s_tw = s_U + (size_t)M * KP;                                // fft_table_len(N)
  float2 *s_twp = s_tw + NC;
  float2 *s_w = s_tw + fft_table_len(N);                              // NC (window, pairs of samples)
  float2 *s_lag = s_w + NC;                                           // kLagTabLen
  float2 *s_Uf = s_lag + kLagTabLen;                                  // M * 16: bins 32 q + 16 of every channel
  float *s_bv = reinterpret_cast<float *>(s_Uf + (size_t)M * 16);     // G * WPF
  unsigned char *s_pair = reinterpret_cast<unsigned char *>(s_bv + G * WPF);   // 2 * P

  const int tid = threadIdx.x;
  fft_load_tables<N>(s_tw, tw_g, tid, NT);
  for (int i = tid; i < NC; i += NT) s_w[i] = make_float2(win[2 * i], win[2 * i + 1]);
  for (int i = tid; i < kLagTabLen; i += NT) s_lag[i] = lag_tab[i];
  pair_table<N, G>(s_pair, M, P, tid, NT);
  __syncthreads();
  const int g = tid / TPF, j = tid % TPF, warp = tid >> 5, half = (tid >> 4) & 1;
  const fft_buf_t buf = smem_u32(s_buf + g * fft_buf_len(NC));
  float *scratch = s_scr + warp * 1024 + half * 512;
  const bool vec_ok = ((row_pitch & 1) == 0) && ((hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
  const long long nframes = (long long)B * T;
  const int tasks = (P + 1) / 2;

  for (long long ft = blockIdx.x; ft < nframes; ft += gridDim.x) {
    const int b = (int)(ft / T), t = (int)(ft - (long long)b * T);
    analysis_phase<N, G>(x, row_pitch, M, hop, vec_ok, ft, b, t, s_w, s_tw, s_twp, buf, s_U, s_bv, spec, chan_pow, g, j, s_Uf);
    __syncthreads();
    for (int task = warp; task < tasks; task += NW) {
      const int p = 2 * task + half;
      const bool live = p < P;
      const int pc = live ? p : P - 1;
      tdoa_pair_halfwarp<LCAP>(s_U, s_Uf, s_lag, scratch, s_pair[2 * pc], s_pair[2 * pc + 1], live, pc, max_lag,
                               curves ? curves + ft * P * L : nullptr, lags + ft * P, nullptr);
    }
    __syncthreads();
  }
}

// lag-phase table of tdoa_warp.cuh, one copy per device (double precision on the host)
static const float2 *lag_table_device() {
  static float2 *tabs[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!tabs[dev]) {
    std::vector<float2> h(kLagTabLen);
    for (int i = 0; i < 32; ++i)
      for (int s = 0; s < 16; ++s) {
        // weight 1/2 on sub-sequence 0 (it is its own Hermitian partner); row 1 is read as the per-lag ROTATION of the recurrence: no weight
        const double a1 = 2.0 * M_PI * (double)(s * i) / 1024.0, w = (s == 0 && i != 1) ? 0.5 : 1.0;
        const double a2 = 2.0 * M_PI * (double)((32 * s + 16) * i) / 1024.0;
        h[i * 16 + s] = make_float2((float)(w * std::cos(a1)), (float)(w * std::sin(a1)));
        h[512 + i * 16 + s] = make_float2((float)std::cos(a2), (float)std::sin(a2));
      }
    float2 *d = nullptr;
    if (cudaMalloc(&d, sizeof(float2) * kLagTabLen) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), sizeof(float2) * kLagTabLen, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    tabs[dev] = d;
  }
  return tabs[dev];
}

template <int LCAP>
static int launch_stft_tdoa_warp(const float *x, long long row_pitch, int B, int T, int M, int hop, int max_lag, const float *win, const float2 *tw,
                                 float2 *spec, float *chan_pow, float *curves, int32_t *lags, cudaStream_t st) {
  constexpr int N = 1024, NC = 512, G = 4, NW = 8;
  const int P = M * (M - 1) / 2;
  const float2 *lag_tab = lag_table_device();
  if (!lag_tab) return mcag_set_error(2, "tdoa: could not create the lag table");
  size_t smem = sizeof(float2) * ((size_t)M * spec_pitch(N) + fft_table_len(N) + (size_t)G * fft_buf_len(NC) + NC + kLagTabLen + (size_t)M * 16) + 4 * NW * 1024 +
                4 * G * 2 + 2 * P + 16 + 8 * NC /* buffer alignment slack */;
  if (smem > 110 * 1024) return -1;   // more microphones than two CTAs per SM can stage: the caller falls back to the general kernel
  auto kern = stft_tdoa_warp_kernel<LCAP>;
  static int sm_count = 0, dev_cached = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) { cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); dev_cached = dev; }
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
  if (per_sm < 1) per_sm = 1;
  const long long nframes = (long long)B * T;
  const long long grid = nframes < (long long)sm_count * per_sm ? nframes : (long long)sm_count * per_sm;
  kern<<<(unsigned)grid, 256, smem, st>>>(x, row_pitch, B, T, M, hop, max_lag, win, tw, lag_tab, spec, chan_pow, curves, lags);
  MCAG_CHECK_LAUNCH();
  return 0;
}

template <int N> static int launch_tdoa(const float2 *spec, int B, int T, int M, int max_lag, const float2 *tw, float *curves, int32_t *lags,
                                        float *peaks, cudaStream_t st) {
  constexpr int NC = N / 2, TPF = NC / 8;
  constexpr int G = (TPF >= 128) ? 2 : (256 / TPF);
  auto smem_for = [&](int cs, int pl) {
    return sizeof(float2) * ((size_t)cs * spec_pitch(N) + fft_table_len(N) + (size_t)G * fft_buf_len(NC)) + 16 * G * ((TPF + 31) / 32) + 6 * (size_t)pl + 16 +
           8 * NC /* buffer alignment slack */;
  };
  auto kern = tdoa_kernel<N, G>;
  const size_t cap = 220 * 1024;
  if (smem_for(M, M * (M - 1) / 2) <= cap) {
    const size_t smem = smem_for(M, M * (M - 1) / 2);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<dim3(T, B), G * TPF, smem, st>>>(spec, T, M, max_lag, tw, curves, lags, peaks, 0, M, 0, 0);
    MCAG_CHECK_LAUNCH();
    return 0;
  }
  // channel tiles of C channels: 2 C spectra staged per launch
  int C = M;
  while (C > 1 && smem_for(2 * C, C * C) > cap) --C;
  if (smem_for(2 * C, C * C) > cap) return mcag_set_error(1, "tdoa: frame size too large for the shared-memory staged kernel");
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(2 * C, C * C));
  for (int i0 = 0; i0 < M; i0 += C) {
    const int ni = (M - i0 < C) ? M - i0 : C;
    if (ni > 1) {
      kern<<<dim3(T, B), G * TPF, smem_for(ni, ni * (ni - 1) / 2), st>>>(spec, T, M, max_lag, tw, curves, lags, peaks, i0, ni, 0, 0);
      MCAG_CHECK_LAUNCH();
    }
    for (int j0 = i0 + C; j0 < M; j0 += C) {
      const int nj = (M - j0 < C) ? M - j0 : C;
      kern<<<dim3(T, B), G * TPF, smem_for(ni + nj, ni * nj), st>>>(spec, T, M, max_lag, tw, curves, lags, peaks, i0, ni, j0, nj);
      MCAG_CHECK_LAUNCH();
    }
  }
  return 0;
}

template <int N> static int launch_stft_tdoa(const float *x, long long row_pitch, int B, int T, int M, int hop, int max_lag, const float *win,
                                             const float2 *tw, float2 *spec, float *chan_pow, float *curves, int32_t *lags, cudaStream_t st) {
  constexpr int NC = N / 2, TPF = NC / 8;
  constexpr int G = (TPF >= 128) ? 4 : (256 / TPF);
  const int P = M * (M - 1) / 2, L = 2 * max_lag + 1;
  size_t smem = sizeof(float2) * ((size_t)M * spec_pitch(N) + fft_table_len(N) + (size_t)G * fft_buf_len(NC) + NC) +
                16 * G * ((TPF + 31) / 32) + 2 * P + 16 + 8 * NC /* buffer alignment slack */;
  if (smem > 220 * 1024) return -1;   // more spectra than one CTA can stage: the caller runs stft + the channel-tiled lag kernel
  auto kern = stft_tdoa_kernel<N, G>;
  static int sm_count = 0, dev_cached = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) { cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); dev_cached = dev; }
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G * TPF, smem);
  if (per_sm < 1) per_sm = 1;
  const long long nframes = (long long)B * T;
  const long long grid = nframes < (long long)sm_count * per_sm ? nframes : (long long)sm_count * per_sm;
  kern<<<(unsigned)grid, G * TPF, smem, st>>>(x, row_pitch, B, T, M, hop, max_lag, win, tw, spec, chan_pow, curves, lags);
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

bool k_stft_tdoa_fits(int M, int N) {   // the fused kernel stages all M spectra of a frame in one CTA
  const int NC = N / 2, TPF = NC / 8, G = (TPF >= 128) ? 4 : (256 / TPF);
  return sizeof(float2) * ((size_t)M * spec_pitch(N) + fft_table_len(N) + (size_t)G * fft_buf_len(NC) + NC) + 16 * G * ((TPF + 31) / 32) + (size_t)M * (M - 1) + 16 +
             8 * NC <= 220 * 1024;
}

int k_stft_tdoa(const float *x, long long row_pitch, int B, int T, int M, int N, int hop, int max_lag, const float *win, const float2 *tw,
                float2 *spec, float *chan_pow, float *curves, int32_t *lags, cudaStream_t st) {
  if (T <= 0 || B <= 0) return 0;
  if (!k_stft_tdoa_fits(M, N)) {   // large arrays: spectra through HBM (the caller provides `spec`), then the channel-tiled lag kernel
    if (!spec) return mcag_set_error(1, "tdoa: this array needs a spectrum buffer (more channels than the fused kernel stages)");
    const int rc = k_stft(x, row_pitch, B * M, M, T, N, hop, win, tw, spec, chan_pow, nullptr, st);
    if (rc) return rc;
    return k_tdoa_lags(spec, B, T, M, N, max_lag, tw, curves, lags, nullptr, st);
  }
  if (M < 2 || M > 255) return mcag_set_error(1, "tdoa: need 2..255 channels");
  if (max_lag < 0 || max_lag > N / 2 - 1) return mcag_set_error(1, "tdoa: max_lag out of range");
  if (hop <= 0 || hop > N) return mcag_set_error(1, "tdoa: bad hop");
  if (N == 1024 && max_lag <= 31 && !getenv("MCAG_TDOA_GENERAL")) {   // warp-synchronous decimated lag phase (tdoa_warp.cuh)
    const int rc = max_lag <= 28 ? launch_stft_tdoa_warp<28>(x, row_pitch, B, T, M, hop, max_lag, win, tw, spec, chan_pow, curves, lags, st)
                                 : launch_stft_tdoa_warp<31>(x, row_pitch, B, T, M, hop, max_lag, win, tw, spec, chan_pow, curves, lags, st);
    if (rc >= 0) return rc;
  }
  switch (N) {
    case 256: return launch_stft_tdoa<256>(x, row_pitch, B, T, M, hop, max_lag, win, tw, spec, chan_pow, curves, lags, st);
    case 512: return launch_stft_tdoa<512>(x, row_pitch, B, T, M, hop, max_lag, win, tw, spec, chan_pow, curves, lags, st);
    case 1024: return launch_stft_tdoa<1024>(x, row_pitch, B, T, M, hop, max_lag, win, tw, spec, chan_pow, curves, lags, st);
    case 2048: return launch_stft_tdoa<2048>(x, row_pitch, B, T, M, hop, max_lag, win, tw, spec, chan_pow, curves, lags, st);
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

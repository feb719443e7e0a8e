// K1 stft_r2c and K7 istft_ola.  Replace DSPONE's STFT::frameAnalysis / frameSynthesis and the
// framing + overlap-add of ShortTimeProcess::process that every reference processor runs
// (SourceSeparationAndLocalisation.cpp:51-52, SourceLocalisation.cpp:51-52, BinauralLocalisation.cpp:320-322,
// FastBinauralMasking.cpp:57; consumer mcabeamf.cpp:112-119).
#include <cstdlib>
#include "fft.cuh"
#include "fft16.cuh"
#include "kernels.h"

namespace mcag {

int k_frame_power_raw(const float2 *spec, long long rows, int N, float *raw, cudaStream_t st);   // capi.cu

// ---------------------------------------------------------------------------------------------------
// K1: persistent CTAs walk work items = F consecutive frames of one (stream, channel) row.  The window and the FFT
// tables are loaded once per CTA; the (F-1)*hop + N samples the frames of an item share are staged into shared memory
// by a 1-D bulk async copy (TMA engine), double-buffered so the copy of the next item runs under the transforms of the
// current one.  Each frame is windowed while it is packed into the N/2-point complex FFT, and the one-sided spectrum is
// written with its Parseval power.  HBM traffic per frame: hop*4 B in (+ overlap from L2), (N/2+2)*8 B out.
// ---------------------------------------------------------------------------------------------------
template <int N, int F, int G>
__global__ void __launch_bounds__(G *(N / 16)) stft_kernel(const float *__restrict__ x, long long row_pitch, int M, int T, int hop,
                                                            const float *__restrict__ win, const float2 *__restrict__ tw_g,
                                                            float2 *__restrict__ spec, float *__restrict__ chan_pow, float *__restrict__ chan_raw,
                                                            int tiles_per_row, long long n_items) {
  constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), NT = G * TPF, XLEN = (F - 1) * N + N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = fft_align_smem(smem_raw, 8 * NC);
  float2 *s_buf = reinterpret_cast<float2 *>(smem);                   // G * fft_buf_len(NC), each buffer aligned to its size
  float *s_xb = reinterpret_cast<float *>(s_buf + G * fft_buf_len(NC));  // 2 x ((F-1)*hop_max + N) floats, hop <= N
  float *s_w = s_xb + 2 * XLEN;                                       // N
  float2 *s_tw = reinterpret_cast<float2 *>(s_w + N);                 // fft_table_len(N): tw[NC] then twp
  float2 *s_twp = s_tw + NC;                                          // per-thread inter-pass twiddles
  float *s_red = reinterpret_cast<float *>(s_tw + fft_table_len(N));  // 2 x G * (TPF/32 or 1): Parseval power, plain sum of |X|^2
  __shared__ __align__(8) uint64_t s_bar[2];

  const int tid = threadIdx.x;
  // every item starts at a 16-byte aligned sample and has a multiple of 4 samples when this holds
  const bool bulk = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((row_pitch & 3) == 0) && ((hop & 3) == 0);
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
  for (int i = tid; i < N; i += NT) s_w[i] = win[i];
  fft_load_tables<N>(s_tw, tw_g, tid, NT);
  __syncthreads();

  auto item_src = [&](long long item, int &row, int &t0, int &nf) {
    row = (int)(item / tiles_per_row);
    t0 = (int)(item - (long long)row * tiles_per_row) * F;
    nf = min(F, T - t0);
    return x + (long long)row * row_pitch + (long long)t0 * hop;
  };
  auto issue = [&](long long item, int slot) {   // thread 0 only
    int row, t0, nf;
    const float *src = item_src(item, row, t0, nf);
    const uint32_t bytes = (uint32_t)((nf - 1) * hop + N) * 4u;
    mbar_expect_tx(&s_bar[slot], bytes);
    bulk_g2s(s_xb + slot * XLEN, src, bytes, &s_bar[slot]);
  };

  const int g = tid / TPF, j = tid % TPF;
  const fft_buf_t buf = smem_u32(s_buf + g * fft_buf_len(NC));
  if (bulk && tid == 0 && (long long)blockIdx.x < n_items) issue(blockIdx.x, 0);
  int it = 0;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    const int slot = it & 1;
    int row, t0, nf;
    const float *src = item_src(item, row, t0, nf);
    float *s_x = s_xb + slot * XLEN;
    if (bulk) {
      // the other slot was last read before the __syncthreads that ended the previous iteration
      if (tid == 0 && item + gridDim.x < n_items) issue(item + gridDim.x, slot ^ 1);
      mbar_wait(&s_bar[slot], (uint32_t)(it >> 1) & 1u);
    } else {
      const int nsamp = (nf - 1) * hop + N;
      for (int i = tid; i < nsamp; i += NT) s_x[i] = src[i];
      __syncthreads();
    }
    const int b = row / M, m = row - b * M;

    for (int f = g; f < F; f += G) {   // uniform trip count per group; inactive frames are skipped as a group
      if (f < nf) {
        const float2 *xs = reinterpret_cast<const float2 *>(s_x + f * hop);   // hop is even -> 8-byte aligned
        const float2 *ws = reinterpret_cast<const float2 *>(s_w);
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int n = j + r * (NC / 8);
          float2 a = xs[n], w = ws[n];
          v[r] = make_float2(a.x * w.x, a.y * w.y);
        }
        fft_run<NC, false>(v, buf, s_twp, j, g);
        // real post-processing: X[k] = E + W^k O, X[NC-k] = conj(E - W^k O); k = j + i*TPF (i < 4), and k = NC/2 for j = 0
        float2 *out = spec + (((long long)b * T + (t0 + f)) * M + m) * KP;
        float pw = 0.f, pr = 0.f;
        auto post = [&](int k) {
          float2 zk = fft_buf_get(buf, k), zn = fft_buf_get(buf, (NC - k) & (NC - 1));
          float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
          float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));   // -i/2 (zk - conj zn)
          float2 wo = cmul(s_tw[k], o);                                          // k <= NC/2 < N/2: straight from the table
          float2 xk = cadd(e, wo), xn = cconj(csub(e, wo));
          if (k == 0) { xk.y = 0.f; xn.y = 0.f; }
          out[k] = xk;
          out[NC - k] = xn;
          const float wk = (k == 0) ? 1.f : 2.f;
          const float mk = xk.x * xk.x + xk.y * xk.y, mn = (k != NC - k) ? xn.x * xn.x + xn.y * xn.y : 0.f;
          pw += wk * (mk + mn);
          pr += mk + mn;   // plain sum over the K bins: mean square of the CCS buffer (BinauralLocalisation.cpp:390-391)
        };
#pragma unroll
        for (int i = 0; i < 4; ++i) post(j + i * TPF);
        if (j == 0) { post(NC / 2); out[NC + 1] = make_float2(0.f, 0.f); }   // middle bin and the pad bin
        // Parseval power of the windowed frame, reduced in a fixed order
        constexpr int WPF = (TPF + 31) / 32;
        if constexpr (TPF >= 32) {   // N = 256 packs two transforms per warp: its power comes from frame_power_kernel
          pw = warp_sum(pw);
          if (chan_raw) pr = warp_sum(pr);
          if (chan_pow) {
            if ((tid & 31) == 0) { s_red[g * WPF + (j >> 5)] = pw; s_red[(G + g) * WPF + (j >> 5)] = pr; }
            group_sync<TPF>(g);
            if (j == 0) {
              float sacc = 0.f, racc = 0.f;
              for (int i = 0; i < WPF; ++i) { sacc += s_red[g * WPF + i]; racc += s_red[(G + g) * WPF + i]; }
              chan_pow[((long long)b * T + (t0 + f)) * M + m] = sacc / ((float)N * (float)N);
              if (chan_raw) chan_raw[((long long)b * T + (t0 + f)) * M + m] = racc;
            }
          }
        }
        group_sync<TPF>(g);
      }
    }
    __syncthreads();   // all frames of the item are done with s_x[slot] before a later copy lands in it
  }
}

// ---------------------------------------------------------------------------------------------------
// K1 for N = 512 and N = 1024 on the register-resident engine of fft16.cuh: the same persistent work items and bulk-copy staging as
// stft_kernel, one HALF-WARP per frame (HW frames of a row per item), N/32 sample pairs per lane.  Per N = 512 frame the shared memory
// sees the staged samples once, the window, the inter-pass twiddles and one 2 KB transpose: 108 wavefronts (shuffles included) against 197.
// ---------------------------------------------------------------------------------------------------
#ifndef MCAG_STFT512_HW
#define MCAG_STFT512_HW 8
#endif
#ifndef MCAG_STFT512_MINB
#define MCAG_STFT512_MINB 6
#endif
#ifndef MCAG_STFT1024_HW
#define MCAG_STFT1024_HW 8
#endif
#ifndef MCAG_STFT1024_MINB
#define MCAG_STFT1024_MINB 1
#endif
template <int N, int HW, int MINB>
__global__ void __launch_bounds__(HW * 16, MINB) stft_hw_kernel(const float *__restrict__ x, long long row_pitch, int M, int T, int hop,
                                                                 const float *__restrict__ win, const float2 *__restrict__ tw_g,
                                                                 float2 *__restrict__ spec, float *__restrict__ chan_pow, float *__restrict__ chan_raw,
                                                                 int tiles_per_row, long long n_items, int xlen) {
  constexpr int R = N / 32, KP = spec_pitch(N), F = HW, NT = HW * 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *s_xb = reinterpret_cast<float2 *>(smem_raw);          // HW transpose buffers
  float2 *s_t1 = s_xb + HW * fft16_buf_len<R>();
  float2 *s_w = s_t1 + fft16_tab_len<R>();                      // N/2 window pairs
  float *s_xs = reinterpret_cast<float *>(s_w + N / 2);         // 2 x xlen staged samples
  __shared__ __align__(8) uint64_t s_bar[2];

  const int tid = threadIdx.x, l16 = tid & 15, f = tid >> 4;
  const bool bulk = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((row_pitch & 3) == 0) && ((hop & 3) == 0);
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
  fft16_load_table<R>(s_t1, tw_g, tid, NT);
  for (int i = tid; i < N / 2; i += NT) s_w[i] = make_float2(win[2 * i], win[2 * i + 1]);
  const float2 wl = tw_g[l16];
  __syncthreads();

  auto item_src = [&](long long item, int &row, int &t0, int &nf) {   // n_items < 2^31 (checked by the launcher): 32-bit division
    row = (int)((unsigned)item / (unsigned)tiles_per_row);
    t0 = ((int)item - row * tiles_per_row) * F;
    nf = min(F, T - t0);
    return x + (long long)row * row_pitch + (long long)t0 * hop;
  };
  auto issue = [&](long long item, int slot) {   // thread 0 only
    int row, t0, nf;
    const float *src = item_src(item, row, t0, nf);
    const uint32_t bytes = (uint32_t)((nf - 1) * hop + N) * 4u;
    mbar_expect_tx(&s_bar[slot], bytes);
    bulk_g2s(s_xs + slot * xlen, src, bytes, &s_bar[slot]);
  };

  float2 *xbuf = s_xb + f * fft16_buf_len<R>();
  if (bulk && tid == 0 && (long long)blockIdx.x < n_items) issue(blockIdx.x, 0);
  int it = 0;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    const int slot = it & 1;
    int row, t0, nf;
    const float *src = item_src(item, row, t0, nf);
    float *s_x = s_xs + slot * xlen;
    if (bulk) {
      if (tid == 0 && item + gridDim.x < n_items) issue(item + gridDim.x, slot ^ 1);
      mbar_wait(&s_bar[slot], (uint32_t)(it >> 1) & 1u);
    } else {
      const int nsamp = (nf - 1) * hop + N;
      for (int i = tid; i < nsamp; i += NT) s_x[i] = src[i];
      __syncthreads();
    }
    const int b = (int)((unsigned)row / (unsigned)M), m = row - b * M;
    // a half-warp past the end of the row repeats the last frame (its partner half-warp needs it for the warp-wide shuffles) and stores nothing
    const bool live = f < nf;
    const int fr = live ? f : nf - 1;
    const float2 *xs = reinterpret_cast<const float2 *>(s_x + fr * hop);   // hop is even -> 8-byte aligned
    float2 v[R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
      const float2 s = xs[16 * a + l16], w = s_w[16 * a + l16];
      v[a] = make_float2(s.x * w.x, s.y * w.y);
    }
    fft_hw<R, false>(v, xbuf, s_t1, l16);
    const long long orow = ((long long)b * T + (t0 + f)) * M + m;
    float2 *out = spec + orow * KP;
    float nyq, pw = 0.f, x0 = 0.f;
    fft_hw_real_post<R>(v, wl, l16, nyq, [&](int e, float2 X) {
      if (live) out[l16 + 16 * e] = X;
      if (e == 0) x0 = X.x;
      pw += X.x * X.x + X.y * X.y;
    });
    // Parseval weights: 1 for k = 0 and the Nyquist bin, 2 for the others; the plain sum over the K bins beside it
    float pr = pw + (l16 == 0 ? nyq * nyq : 0.f);
    pw = 2.f * pw - (l16 == 0 ? x0 * x0 - nyq * nyq : 0.f);
    if (chan_pow) pw = hw_sum(pw);
    if (chan_raw) pr = hw_sum(pr);
    if (live && l16 == 0) {
      *reinterpret_cast<float4 *>(out + N / 2) = make_float4(nyq, 0.f, 0.f, 0.f);   // the Nyquist bin and the pad bin
      if (chan_pow) chan_pow[orow] = pw / ((float)N * (float)N);
      if (chan_raw) chan_raw[orow] = pr;
    }
    __syncthreads();   // all frames of the item are done with s_x[slot] before a later copy lands in it
  }
}

// power for the small-TPF case is handled by a dedicated reduction kernel (frame_power_kernel below), which is
// also the generic path: spec [rows] -> pow[rows], rows = B*T*M.
__global__ void frame_power_kernel(const float2 *__restrict__ spec, long long rows, int N, float *__restrict__ pow) {
  const int KP = spec_pitch(N), K = N / 2 + 1;
  const long long row = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float2 *s = spec + row * KP;
  float acc = 0.f;
  for (int k = threadIdx.x & 31; k < K; k += 32) {
    float2 v = s[k];
    acc += ((k == 0 || k == K - 1) ? 1.f : 2.f) * (v.x * v.x + v.y * v.y);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) pow[row] = acc / ((float)N * (float)N);
}

// ---------------------------------------------------------------------------------------------------
// K7: one CTA = F consecutive hop-segments of one (stream, channel) row.  It inverse-transforms the
// F + R - 1 frames that overlap them (R = N/hop), applies the synthesis window and sums the overlaps in frame
// order (oldest first, as the reference's overlap-add does).  Segment indices >= T belong to the carried tail.
// F is picked by the launcher so that F + R - 1 is a multiple of the G transform groups (every round of transforms is full).
// ---------------------------------------------------------------------------------------------------
template <int N, int G>
__global__ void __launch_bounds__(G *(N / 16)) istft_kernel(const float2 *__restrict__ spec, int C_in, int C_out, int T, int hop, int F,
                                                             const float *__restrict__ win, const float2 *__restrict__ tw_g,
                                                             const float *__restrict__ tail_in, float *__restrict__ tail_out,
                                                             float *__restrict__ out, long long out_pitch, int out_rows) {
  constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), NT = G * TPF;
  const int R = N / hop;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = fft_align_smem(smem_raw, 8 * NC);
  float2 *s_buf = reinterpret_cast<float2 *>(smem);                   // G * fft_buf_len(NC), each buffer aligned to its size
  float *s_y = reinterpret_cast<float *>(s_buf + G * fft_buf_len(NC));   // (F + R - 1) * N
  float *s_w = s_y + (F + R - 1) * N;                                 // N
  float2 *s_tw = reinterpret_cast<float2 *>(s_w + N);                 // fft_table_len(N): tw[NC] then twp
  float2 *s_twp = s_tw + NC;                                          // per-thread inter-pass twiddles
  float2 *s_in = s_tw + fft_table_len(N);                             // G * KP  (staged spectrum rows)

  const int tid = threadIdx.x, row = blockIdx.y, seg0 = blockIdx.x * F;
  const int b = row / C_out, c = row % C_out;
  const int nfr = F + R - 1;            // frames seg0-R+1 .. seg0+F-1
  for (int i = tid; i < N; i += NT) s_w[i] = win[i];
  fft_load_tables<N>(s_tw, tw_g, tid, NT);
  __syncthreads();

  const int g = tid / TPF, j = tid % TPF;
  const fft_buf_t buf = smem_u32(s_buf + g * fft_buf_len(NC));
  float2 *xin = s_in + g * KP;
  for (int fi = g; fi < nfr; fi += G) {
    const int t = seg0 - (R - 1) + fi;
    float *y = s_y + fi * N;
    if (t < 0 || t >= T) {
      for (int i = j; i < N / 4; i += TPF) reinterpret_cast<float4 *>(y)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float2 *src = spec + (((long long)b * T + t) * C_in + c) * KP;
      for (int k = j; k < KP / 2; k += TPF)   // KP is even and rows are 16-byte aligned: two bins per load
        reinterpret_cast<float4 *>(xin)[k] = reinterpret_cast<const float4 *>(src)[k];
      group_sync<TPF>(g);
      float2 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int k = j + r * (NC / 8);
        float2 xk = xin[k], xn = xin[NC - k];
        if (k == 0) { xk.y = 0.f; xn.y = 0.f; }   // c2r ignores the imaginary parts of DC and Nyquist
        float2 e = make_float2(0.5f * (xk.x + xn.x), 0.5f * (xk.y - xn.y));
        float2 d = make_float2(0.5f * (xk.x - xn.x), 0.5f * (xk.y + xn.y));   // (xk - conj xn)/2
        float2 o = cmul(d, tw_lookup<true>(s_tw, k, NC));                      // * conj(W^k)
        v[r] = make_float2(e.x - o.y, e.y + o.x);                              // E + iO
      }
      fft_run<NC, true>(v, buf, s_twp, j, g);
      const float sc = 1.0f / (float)NC;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int n = j + r * TPF;
        const float2 z = fft_buf_get(buf, n), w = reinterpret_cast<const float2 *>(s_w)[n];
        reinterpret_cast<float2 *>(y)[n] = make_float2(z.x * sc * w.x, z.y * sc * w.y);
      }
      group_sync<TPF>(g);
    }
  }
  __syncthreads();
  // overlap-add, oldest frame first; the carried tail (older still) goes in first of all.  Four samples per thread; hop is a
  // power of two (N / hop is 1, 2 or 4), so segment and offset are a shift and a mask.
  const int ov = N - hop, lh = 31 - __clz(hop), q = hop >> 2;
  float *orow = out + ((long long)b * out_rows + c) * out_pitch;
  const bool vec = ((out_pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int i4 = tid; i4 < F * q; i4 += NT) {
    const int sl = i4 >> (lh - 2), n = (i4 & (q - 1)) << 2, seg = seg0 + sl;
    if (seg >= T + R - 1) break;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tail_in && seg < R - 1) acc = *reinterpret_cast<const float4 *>(tail_in + (long long)row * ov + seg * hop + n);
    for (int r = R - 1; r >= 0; --r) {
      const float4 y4 = *reinterpret_cast<const float4 *>(s_y + (sl + (R - 1) - r) * N + n + r * hop);
      acc.x += y4.x; acc.y += y4.y; acc.z += y4.z; acc.w += y4.w;
    }
    if (seg < T) {
      float *dst = orow + (long long)seg * hop + n;
      if (vec) *reinterpret_cast<float4 *>(dst) = acc;
      else { dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w; }
    } else if (tail_out) {
      *reinterpret_cast<float4 *>(tail_out + (long long)row * ov + (seg - T) * hop + n) = acc;
    }
  }
}

static int device_sm_count() {
  static int sm_count = 0, dev_cached = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) { cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); dev_cached = dev; }
  return sm_count;
}

// N = 512 / 1024 on the half-warp engine (MCAG_STFT_STOCKHAM=1 keeps the shared-memory engine for A/B runs)
template <int N, int HW, int MINB>
static int launch_stft_hw(const float *x, long long row_pitch, int rows, int M, int T, int hop, const float *win, const float2 *tw, float2 *spec,
                          float *chan_pow, float *chan_raw, cudaStream_t st) {
  constexpr int R = N / 32;
  const int xlen = (HW - 1) * hop + N;   // hop is even; the bulk path needs hop % 4 == 0, which keeps both slots 16-byte aligned
  const size_t smem = sizeof(float2) * (HW * fft16_buf_len<R>() + fft16_tab_len<R>() + N / 2) + sizeof(float) * 2 * (size_t)xlen;
  auto kern = stft_hw_kernel<N, HW, MINB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, HW * 16, smem);
  if (per_sm < 1) per_sm = 1;
  const int tiles_per_row = (T + HW - 1) / HW;
  const long long n_items = (long long)tiles_per_row * rows, cap = (long long)device_sm_count() * per_sm;
  if (n_items >= (1ll << 31)) return mcag_set_error(1, "stft: too many frames in one call");
  kern<<<(unsigned)(n_items < cap ? n_items : cap), HW * 16, smem, st>>>(x, row_pitch, M, T, hop, win, tw, spec, chan_pow, chan_pow ? chan_raw : nullptr,
                                                                         tiles_per_row, n_items, xlen);
  MCAG_CHECK_LAUNCH();
  return 0;
}

template <int N> static int launch_stft(const float *x, long long row_pitch, int rows, int M, int T, int hop, const float *win,
                                        const float2 *tw, float2 *spec, float *chan_pow, float *chan_raw, cudaStream_t st) {
  constexpr int NC = N / 2, TPF = NC / 8;
  constexpr int G = (TPF >= 128) ? 2 : (256 / TPF);
  constexpr int F = (N >= 2048) ? 4 : (G > 8 ? G : 8);
  size_t smem = sizeof(float) * 2 * ((F - 1) * N + N) + sizeof(float) * N + sizeof(float2) * fft_table_len(N) + sizeof(float2) * G * fft_buf_len(NC) +
                sizeof(float) * G * 8 + 8 * NC /* buffer alignment slack */;
  auto kern = stft_kernel<N, F, G>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static int sm_count = 0, dev_cached = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) { cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); dev_cached = dev; }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G * TPF, smem);
  if (per_sm < 1) per_sm = 1;
  const int tiles_per_row = (T + F - 1) / F;
  const long long n_items = (long long)tiles_per_row * rows, cap = (long long)sm_count * per_sm;
  float *pow_in_kernel = (TPF >= 32) ? chan_pow : nullptr;
  kern<<<(unsigned)(n_items < cap ? n_items : cap), G * TPF, smem, st>>>(x, row_pitch, M, T, hop, win, tw, spec, pow_in_kernel,
                                                                         pow_in_kernel ? chan_raw : nullptr, tiles_per_row, n_items);
  MCAG_CHECK_LAUNCH();
  if (chan_pow && TPF < 32) {
    long long nrows = (long long)(rows / M) * T * M;
    frame_power_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, st>>>(spec, nrows, N, chan_pow);
    MCAG_CHECK_LAUNCH();
    if (chan_raw) return k_frame_power_raw(spec, nrows, N, chan_raw, st);
  }
  return 0;
}

template <int N> static int launch_istft(const float2 *spec, int B, int T, int C_in, int C_out, int hop, const float *win, const float2 *tw,
                                         const float *tail_in, float *tail_out, float *out, long long out_pitch, int out_rows, cudaStream_t st) {
  constexpr int NC = N / 2, TPF = NC / 8;
  constexpr int G = (TPF >= 128) ? 2 : (256 / TPF);
  const int R = N / hop;
  // frames transformed per CTA: a multiple of G, about 16 but at most 64 KB of staged time-domain frames
  int nfr = 65536 / (4 * N);
  if (nfr > 16) nfr = 16;
  nfr = nfr / G * G;
  if (nfr < G) nfr = G;
  while (nfr - (R - 1) < 1) nfr += G;
  const int F = nfr - (R - 1);
  size_t smem = sizeof(float) * (size_t)nfr * N + sizeof(float) * N + sizeof(float2) * fft_table_len(N) + sizeof(float2) * G * fft_buf_len(NC) +
                sizeof(float2) * G * spec_pitch(N) + 8 * NC /* buffer alignment slack */;
  auto kern = istft_kernel<N, G>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((T + R - 1 + F - 1) / F, B * C_out);
  kern<<<grid, G * TPF, smem, st>>>(spec, C_in, C_out, T, hop, F, win, tw, tail_in, tail_out, out, out_pitch, out_rows);
  MCAG_CHECK_LAUNCH();
  return 0;
}

int k_stft(const float *x, long long row_pitch, int rows, int M, int T, int N, int hop, const float *win, const float2 *tw, float2 *spec,
           float *chan_pow, float *chan_raw, cudaStream_t st) {
  if (T <= 0) return 0;
  if (hop <= 0 || hop > N || (hop & 1) || (N % hop)) return mcag_set_error(1, "stft: hop must be even and divide N");
  switch (N) {
    case 256: return launch_stft<256>(x, row_pitch, rows, M, T, hop, win, tw, spec, chan_pow, chan_raw, st);
    case 512: {
      static const bool stockham = getenv("MCAG_STFT_STOCKHAM") != nullptr;
      if (!stockham) return launch_stft_hw<512, MCAG_STFT512_HW, MCAG_STFT512_MINB>(x, row_pitch, rows, M, T, hop, win, tw, spec, chan_pow, chan_raw, st);
      return launch_stft<512>(x, row_pitch, rows, M, T, hop, win, tw, spec, chan_pow, chan_raw, st);
    }
    case 1024: {
      static const bool stockham = getenv("MCAG_STFT_STOCKHAM") != nullptr;
      if (!stockham) return launch_stft_hw<1024, MCAG_STFT1024_HW, MCAG_STFT1024_MINB>(x, row_pitch, rows, M, T, hop, win, tw, spec, chan_pow, chan_raw, st);
      return launch_stft<1024>(x, row_pitch, rows, M, T, hop, win, tw, spec, chan_pow, chan_raw, st);
    }
    case 2048: return launch_stft<2048>(x, row_pitch, rows, M, T, hop, win, tw, spec, chan_pow, chan_raw, st);
  }
  return mcag_set_error(1, "stft: frame size must be 256, 512, 1024 or 2048");
}

int k_istft(const float2 *spec, int B, int T, int C_in, int C_out, int N, int hop, const float *win, const float2 *tw, const float *tail_in,
            float *tail_out, float *out, long long out_pitch, int out_rows, cudaStream_t st) {
  if (T <= 0) return 0;
  if (out_rows < C_out) out_rows = C_out;
  if (hop <= 0 || hop > N || (N % hop) || N / hop > 4) return mcag_set_error(1, "istft: hop must divide N with N/hop <= 4");
  switch (N) {
    case 256: return launch_istft<256>(spec, B, T, C_in, C_out, hop, win, tw, tail_in, tail_out, out, out_pitch, out_rows, st);
    case 512: return launch_istft<512>(spec, B, T, C_in, C_out, hop, win, tw, tail_in, tail_out, out, out_pitch, out_rows, st);
    case 1024: return launch_istft<1024>(spec, B, T, C_in, C_out, hop, win, tw, tail_in, tail_out, out, out_pitch, out_rows, st);
    case 2048: return launch_istft<2048>(spec, B, T, C_in, C_out, hop, win, tw, tail_in, tail_out, out, out_pitch, out_rows, st);
  }
  return mcag_set_error(1, "istft: frame size must be 256, 512, 1024 or 2048");
}

int k_frame_power(const float2 *spec, long long rows, int N, float *pow, cudaStream_t st) {
  if (rows <= 0) return 0;
  frame_power_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(spec, rows, N, pow);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

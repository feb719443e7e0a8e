// K6 fused: FastBinauralMasking::processParametrisation (FastBinauralMasking.cpp:126-538) together with the DSPONE analysis and
// synthesis around it, as ONE kernel per process call: windowed load -> FFT -> 45-band statistics -> per-band power tracker and mask
// decision -> gains -> inverse FFT -> synthesis window -> overlap-add.  A group of threads owns one stereo stream and walks its frames
// in time order, so the spectra, the band statistics and the gains never leave the SM (the staged path of mask.cu materialises the
// spectra and re-reads them three times: 6.57 GB of DRAM traffic per cfg1m step against 1.05 GB of samples in + samples out,
// profiles/traffic.json) and the recurrences (Q_b, the overlap tail) live in registers / shared memory between frames.
//
// Every per-frame expression is the one of the staged kernels (mask_stats_kernel / mask_band_step / mask_apply_kernel, stft_kernel,
// istft_kernel), in the same order, so the two paths agree to the last bit on decisions and within rounding of the FFT schedule on
// audio; tests compare them.  The staged path remains for hop != N/2, N = 256, more bands than threads of a stream, the NOTHING method
// and MCAG_EMIT_SPECTRA.
#include "fft.cuh"
#include "kernels.h"
#include "mask_common.cuh"

namespace mcag {

struct MfParams {
  const float *x;            // [B*2][row_pitch] samples: frame t of a row starts at t * hop
  long long row_pitch;
  int B, T, hop;
  const float *win;          // [N]
  const float2 *tw;          // fft tables of N
  // compact filter-bank tables (built by mcag_create, copied to shared memory once per CTA), 4-byte words:
  //   binfo[nb][4] = {lo, hi, off, -}: band b covers bins [lo, hi), its squared magnitudes are h2c[off + k - lo]
  //   kinfo[K]     = lo | hi << 8 | off << 16: bin k is covered by bands [lo, hi), their magnitudes at k are hc[off + b - lo]
  //   h2c[n_h2c], hc[n_hc]
  const int *tab;
  int n_h2c, n_hc;
  int nb, method, alg, first_call;
  const float *thresholds;   // [nb]
  float *Q, *noise;          // [B][nb] carried state
  const float *tail_in;      // [B*2][N - hop]
  float *tail_out;
  float *out;                // [B*out_rows][out_pitch]
  long long out_pitch;
  int out_rows;
  float *chan_pow;           // [B][T][2] Parseval power of the windowed frames (the gate's input)
  unsigned char *decisions;  // optional [B][T][nb]
  float *q_trace;            // optional [B][T][nb]
};

// SPC streams per CTA; a stream = 2 groups (left, right) of TPF = N/16 threads
template <int N, int SPC>
__global__ void __launch_bounds__(SPC * 2 * (N / 16), 896 / (SPC * 2 * (N / 16))) mask_fused_kernel(const MfParams p) {
  constexpr int NC = N / 2, TPF = NC / 8, KP = spec_pitch(N), K = N / 2 + 1, NH = N / 2, NTS = 2 * TPF, WPF = (TPF + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = fft_align_smem(smem_raw, 8 * NC);
  float2 *s_buf = reinterpret_cast<float2 *>(smem);                       // SPC * 2 transform buffers, each aligned to its size
  float2 *s_tw = s_buf + SPC * 2 * fft_buf_len(NC);                       // fft_table_len(N): tw[NC] then twp (shared by the CTA)
  float2 *s_twp = s_tw + NC;
  float2 *s_w = s_tw + fft_table_len(N);                                  // NC window pairs (shared)
  float2 *s_X_all = s_w + NC;                                             // SPC * 2 * KP spectra
  float2 *s_g_all = s_X_all + (size_t)SPC * 2 * KP;                       // SPC * nb gains (gl, gr)
  float *s_red_all = reinterpret_cast<float *>(s_g_all + (size_t)SPC * p.nb);   // SPC * 2 * WPF Parseval partials
  int *s_tab = reinterpret_cast<int *>(s_red_all + ((SPC * 2 * WPF + 3) & ~3));   // binfo, kinfo, h2c, hc (16-byte aligned: nb * SPC is padded below)
  s_tab += (4 - ((SPC * p.nb * 2) & 3)) & 3;
  const int4 *s_binfo = reinterpret_cast<const int4 *>(s_tab);
  const int *s_kinfo = s_tab + 4 * p.nb;
  const float *s_h2c = reinterpret_cast<const float *>(s_kinfo + K), *s_hc = s_h2c + p.n_h2c;

  const int tid = threadIdx.x;
  fft_load_tables<N>(s_tw, p.tw, tid, blockDim.x);
  for (int i = tid; i < NC; i += blockDim.x) s_w[i] = make_float2(p.win[2 * i], p.win[2 * i + 1]);
  for (int i = tid; i < 4 * p.nb + K + p.n_h2c + p.n_hc; i += blockDim.x) s_tab[i] = p.tab[i];
  __syncthreads();

  const int sl = tid / NTS, ts = tid % NTS;          // stream slot of the CTA, thread of the stream
  const int c = ts / TPF, j = ts % TPF;              // channel (0 left, 1 right), thread of the transform
  const int g = sl * 2 + c;                          // transform group of the CTA (named barrier g + 1 when TPF > 32)
  const int b = blockIdx.x * SPC + sl;               // stream
  if (b >= p.B) return;                              // whole streams drop out together: no CTA-wide barrier below
  const fft_buf_t buf = smem_u32(s_buf + g * fft_buf_len(NC));
  float2 *s_X = s_X_all + (size_t)sl * 2 * KP, *X = s_X + c * KP;
  float2 *s_g = s_g_all + (size_t)sl * p.nb;
  float *s_red = s_red_all + sl * 2 * WPF;
  // named barrier of the stream with an IMMEDIATE id: with the id in a register ptxas reserves all 16 hardware barriers for the CTA and
  // the SM then fits 4 CTAs instead of 7 (ncu r2_cfg1m_fused: launch__occupancy_limit_barriers = 4)
  auto stream_sync = [&]() {
    if constexpr (TPF <= 32) {
      switch (sl) {
        case 0: asm volatile("bar.sync 1, %0;" ::"n"(NTS) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"n"(NTS) : "memory"); break;
        case 2: asm volatile("bar.sync 3, %0;" ::"n"(NTS) : "memory"); break;
        default: asm volatile("bar.sync 4, %0;" ::"n"(NTS) : "memory"); break;
      }
    } else {
      asm volatile("bar.sync %0, %1;" ::"r"(2 * SPC + 1 + sl), "n"(NTS) : "memory");   // group_sync already uses ids 1 .. 2 SPC here
    }
  };

  const float *src = p.x + ((long long)b * 2 + c) * p.row_pitch;
  float *orow = p.out + ((long long)b * p.out_rows + c) * p.out_pitch;
  const float sc = 1.0f / (float)NC;
  // overlap tail of this channel: the second half of the previous frame's synthesis, 4 pairs per thread (n = j + r TPF, r < 4)
  float2 tail[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) tail[r] = reinterpret_cast<const float2 *>(p.tail_in + ((long long)b * 2 + c) * (N - p.hop))[j + r * TPF];
  // band state of the stream: thread ts owns band ts (k_mask_fused_supported: nb <= threads of a stream)
  constexpr int MF_BPT = 1;
  float Qr[MF_BPT], noise_r[MF_BPT];
#pragma unroll
  for (int u = 0; u < MF_BPT; ++u) {
    const int bd = ts + u * NTS;
    Qr[u] = bd < p.nb ? p.Q[(long long)b * p.nb + bd] : 0.f;
    noise_r[u] = bd < p.nb ? p.noise[(long long)b * p.nb + bd] : 0.f;
  }
  int first_call = p.first_call;
  const int4 my_band = s_binfo[ts < p.nb ? ts : 0];   // {lo, hi, off} of the band this thread sums
  const float my_thr = p.thresholds[ts < p.nb ? ts : 0];

  // raw samples of the current frame, n = j + r TPF pairs: v[4..7] of frame t are v[0..3] of frame t + 1 (hop = N/2), so a frame costs
  // hop new samples per channel; they are requested one frame ahead (nxt) and arrive under the transforms of the current frame
  float2 v[8], nxt[4];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = __ldg(reinterpret_cast<const float2 *>(src) + j + r * TPF);

  for (int t = 0; t < p.T; ++t) {
    // ---- analysis: window, forward transform
    const int tn = (t + 1 < p.T) ? t + 1 : t;   // always a load (the last frame re-reads its own half): no divergent zero fill
#pragma unroll
    for (int r = 0; r < 4; ++r) nxt[r] = __ldg(reinterpret_cast<const float2 *>(src + (long long)tn * p.hop) + j + (r + 4) * TPF);
    float2 a[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { const float2 w = s_w[j + r * TPF]; a[r] = make_float2(v[r].x * w.x, v[r].y * w.y); }
    fft_run<NC, false>(a, buf, s_twp, j, g);
    // real post-processing (stft_kernel): X[k] = E + W^k O, X[NC-k] = conj(E - W^k O); k = j + i TPF (i < 4), and k = NC/2 for j = 0
    float pw = 0.f;
    auto post = [&](int k) {
      float2 zk = fft_buf_get(buf, k), zn = fft_buf_get(buf, (NC - k) & (NC - 1));
      float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
      float2 wo = cmul(s_tw[k], o);
      float2 xk = cadd(e, wo), xn = cconj(csub(e, wo));
      if (k == 0) { xk.y = 0.f; xn.y = 0.f; }
      X[k] = xk;
      X[NC - k] = xn;
      const float wk = (k == 0) ? 1.f : 2.f;
      const float mk = xk.x * xk.x + xk.y * xk.y, mn = (k != NC - k) ? xn.x * xn.x + xn.y * xn.y : 0.f;
      pw += wk * (mk + mn);
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) post(j + i * TPF);
    if (j == 0) { post(NC / 2); X[NC + 1] = make_float2(0.f, 0.f); }
    if (p.chan_pow) {   // Parseval power of the windowed frame, reduced in the fixed order of stft_kernel
      if constexpr (TPF >= 32) {
        pw = warp_sum(pw);
        if ((tid & 31) == 0) s_red[c * WPF + (j >> 5)] = pw;
      } else {
        const unsigned gmask = ((1u << TPF) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(TPF - 1));
#pragma unroll
        for (int o2 = TPF / 2; o2 > 0; o2 >>= 1) pw += __shfl_xor_sync(gmask, pw, o2);
        if (j == 0) s_red[c] = pw;
      }
    }
    stream_sync();   // both spectra of the frame are in s_X
    if (p.chan_pow && j == 0) {
      float sacc = 0.f;
      for (int i = 0; i < WPF; ++i) sacc += s_red[c * WPF + i];
      p.chan_pow[((long long)b * p.T + t) * 2 + c] = sacc / ((float)N * (float)N);
    }
    // ---- band statistics (mask_stats_kernel: one thread per band, ascending bins) and the tracker / decision / gains of the frame
#pragma unroll
    for (int u = 0; u < MF_BPT; ++u) {
      const int bd = ts + u * NTS;
      if (bd < p.nb) {
        const float *h = s_h2c + my_band.z - my_band.x;
        float st[MS_NSTAT] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int hi = my_band.y;
        for (int k = my_band.x; k < hi; ++k) {
          const float w = h[k];
          const float2 l = s_X[k], r = s_X[KP + k];
          const float mx = 0.5f * l.x + 0.5f * r.x, my = 0.5f * l.y + 0.5f * r.y;
          const float q0 = mx * mx + my * my, q1 = r.x * l.x + r.y * l.y, q2 = l.x * l.x + l.y * l.y, q3 = r.x * r.x + r.y * r.y;
          st[1] = fmaf(w, q1, st[1]); st[2] = fmaf(w, q2, st[2]); st[3] = fmaf(w, q3, st[3]);
          if (k < NH) { st[0] = fmaf(w, q0, st[0]); st[4] = fmaf(w, q2, st[4]); st[5] = fmaf(w, q3, st[5]); }
        }
        float gl, gr;
        const int dec = mask_band_step(st, N, p.method, p.alg, my_thr, first_call, Qr[u], noise_r[u], gl, gr);
        s_g[bd] = make_float2(gl, gr);
        const long long o = ((long long)b * p.T + t) * p.nb + bd;
        if (p.decisions) p.decisions[o] = (unsigned char)dec;
        if (p.q_trace) p.q_trace[o] = Qr[u];
      }
    }
    ++first_call;
    stream_sync();   // gains of the frame are in s_g
    // ---- apply (mask_apply_kernel): X[k] *= sum_b gain_b H_b[k] over the bands that cover bin k, in band order; the Nyquist bin
    //      is summed unmasked, the pad bin gets weight 0
    for (int k = j; k < KP; k += TPF) {
      float wgt = 0.f;
      if (k < K) {
        const int info = s_kinfo[k], blo = info & 255, bhi = (info >> 8) & 255;
        const float *hk = s_hc + (info >> 16) - blo;
        for (int bd = blo; bd < bhi; ++bd) {
          const float2 gg = s_g[bd];
          wgt = fmaf(hk[bd], (k < NH) ? (c ? gg.y : gg.x) : 1.f, wgt);
        }
      }
      const float2 xv = X[k];
      X[k] = make_float2(xv.x * wgt, xv.y * wgt);
    }
    group_sync<TPF>(g);   // this channel's masked spectrum is complete (only this group reads it below)
    // ---- synthesis (istft_kernel): E + iO build, inverse transform, window, overlap-add with the carried tail
    float2 z[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int k = j + r * (NC / 8);
      float2 xk = X[k], xn = X[NC - k];
      if (k == 0) { xk.y = 0.f; xn.y = 0.f; }
      float2 e = make_float2(0.5f * (xk.x + xn.x), 0.5f * (xk.y - xn.y));
      float2 d = make_float2(0.5f * (xk.x - xn.x), 0.5f * (xk.y + xn.y));
      float2 o = cmul(d, tw_lookup<true>(s_tw, k, NC));
      z[r] = make_float2(e.x - o.y, e.y + o.x);
    }
    // (no stream barrier here: the next frame writes X[c] after this build in program order, the other channel's group stopped
    //  reading X[c] before the barrier above, and the gains are rewritten only behind the next frame's first barrier)
    fft_run<NC, true>(z, buf, s_twp, j, g);
    float2 *dst = reinterpret_cast<float2 *>(orow + (long long)t * p.hop);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int n = j + r * TPF;
      const float2 y = fft_buf_get(buf, n), w = s_w[n];
      const float2 yw = make_float2(y.x * sc * w.x, y.y * sc * w.y);
      if (r < 4) dst[n] = make_float2(tail[r].x + yw.x, tail[r].y + yw.y);   // oldest first: carried tail, then this frame
      else tail[r - 4] = yw;
    }
    group_sync<TPF>(g);   // the transform buffer is free for the next frame
#pragma unroll
    for (int r = 0; r < 4; ++r) { v[r] = v[r + 4]; v[r + 4] = nxt[r]; }
  }
  // ---- carried state
#pragma unroll
  for (int r = 0; r < 4; ++r) reinterpret_cast<float2 *>(p.tail_out + ((long long)b * 2 + c) * (N - p.hop))[j + r * TPF] = tail[r];
#pragma unroll
  for (int u = 0; u < MF_BPT; ++u) {
    const int bd = ts + u * NTS;
    if (bd < p.nb) { p.Q[(long long)b * p.nb + bd] = Qr[u]; p.noise[(long long)b * p.nb + bd] = noise_r[u]; }
  }
}

template <int N> static int launch_mask_fused(const MfParams &p, cudaStream_t st) {
  constexpr int NC = N / 2, TPF = NC / 8, NTS = 2 * TPF;
  constexpr int SPC = (NTS >= 256) ? 1 : (NTS >= 128 ? 2 : (NTS >= 64 ? 2 : 4));
  const size_t smem = sizeof(float2) * ((size_t)SPC * 2 * fft_buf_len(NC) + fft_table_len(N) + NC + (size_t)SPC * 2 * spec_pitch(N) + (size_t)SPC * p.nb) +
                      sizeof(float) * SPC * 2 * ((TPF + 31) / 32) + 4 * (4 * (size_t)p.nb + N / 2 + 1 + p.n_h2c + p.n_hc) + 64 + 8 * NC /* buffer alignment slack */;
  auto kern = mask_fused_kernel<N, SPC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // all of the unified L1 / shared memory as shared memory: the driver's default carve-out fitted 4 CTAs of 27 KB per SM, the
  // registers allow 7 (28 warps; 2048 streams of cfg1m then run as a single wave)
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  kern<<<(unsigned)((p.B + SPC - 1) / SPC), SPC * NTS, smem, st>>>(p);
  MCAG_CHECK_LAUNCH();
  return 0;
}

bool k_mask_fused_supported(int N, int hop, int nb) {
  const int NTS = 2 * (N / 16);
  return hop * 2 == N && N >= 512 && nb >= 1 && nb <= NTS;   // N = 256 packs two transforms per warp: staged path
}

int k_mask_fused(const float *x, long long row_pitch, int B, int T, int N, int hop, const float *win, const float2 *tw, const int *tab, int n_h2c, int n_hc,
                 int nb, int method, int alg, const float *thresholds, float *Q, float *noise,
                 int first_call, const float *tail_in, float *tail_out, float *out, long long out_pitch, int out_rows, float *chan_pow,
                 unsigned char *decisions, float *q_trace, cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  if (!k_mask_fused_supported(N, hop, nb)) return mcag_set_error(1, "mask_fused: unsupported frame size / hop / band count");
  if ((row_pitch & 1) || (out_pitch & 1) || (reinterpret_cast<uintptr_t>(x) & 7) || (reinterpret_cast<uintptr_t>(out) & 7))
    return mcag_set_error(1, "mask_fused: sample rows must be 8-byte aligned");
  MfParams p;
  p.x = x; p.row_pitch = row_pitch; p.B = B; p.T = T; p.hop = hop; p.win = win; p.tw = tw; p.tab = tab; p.n_h2c = n_h2c;
  p.n_hc = n_hc; p.nb = nb; p.method = method; p.alg = alg; p.first_call = first_call; p.thresholds = thresholds; p.Q = Q; p.noise = noise;
  p.tail_in = tail_in; p.tail_out = tail_out; p.out = out; p.out_pitch = out_pitch; p.out_rows = out_rows; p.chan_pow = chan_pow;
  p.decisions = decisions; p.q_trace = q_trace;
  switch (N) {
    case 512: return launch_mask_fused<512>(p, st);
    case 1024: return launch_mask_fused<1024>(p, st);
    case 2048: return launch_mask_fused<2048>(p, st);
  }
  return mcag_set_error(1, "mask_fused: frame size must be 512, 1024 or 2048");
}

}  // namespace mcag

// Warp-synchronous GCC-PHAT lag phase for N = 1024 (BASELINE config 2): replaces the 28 shared-memory 512-point inverse transforms of
// tdoa_pairs by a decimated inverse that never leaves the registers.
//
// The lag window needs S[l] = sum_{k=0}^{N/2} Re(G[k] w^{kl}), w = exp(2 pi i / N), only for |l| <= L < 32 (oracle/CONVENTIONS.md C5;
// G = U_i conj(U_j), Hermitian-extended G_H[N-k] = conj G[k]).  With k = 32 q + s (32 sub-sequences of 32 bins):
//     sum_{k<N} G_H[k] w^{kl} = sum_s w^{sl} h_s[l mod 32],     h_s[m] = sum_{q<32} G_H[32 q + s] exp(2 pi i q m / 32)
// and sub-sequences s and 32 - s are complex conjugates of each other, so
//     S[l] = 1/2 Re h'_0[m] + sum_{s=1}^{15} Re(w^{sl} h_s[m]) + sum_{q<16} Re(G[32 q + 16] w^{(32 q + 16) l}),     m = l mod 32,
// where h'_0 is h_0 with its q = 0 and q = 16 inputs (G[0], G[N/2]) doubled, which absorbs the one-sided end-point terms.
// Sixteen lanes own one pair: lane s runs the 32-point inverse DFT of sub-sequence s entirely in registers (all twiddles are
// compile-time constants) and adds the contribution of ONE bin of the self-paired sub-sequence 16 (bin 32 s + 16) directly.  The per-lane
// lag terms are transposed and summed through a 2 KB swizzled scratch (fixed order: deterministic), each lane ends up with four lags, and
// two REDUX instructions per half-warp give the first-maximum arg-max.  A warp carries two pairs; there is no barrier wider than
// __syncwarp and no transform exchange through shared memory (the 512-point path moves every point through shared memory three times).
#pragma once
#include "fft.cuh"

namespace mcag {

// cos(2 pi n / 32), n mod 32; literals so that every twiddle of the unrolled transform is an immediate
__host__ __device__ constexpr float cos32(int n) {
  n &= 31;
  if (n > 16) n = 32 - n;
  const bool neg = n > 8;
  if (neg) n = 16 - n;
  float v = 0.f;
  switch (n) {
    case 0: v = 1.0f; break;
    case 1: v = 0.98078528040323044913f; break;
    case 2: v = 0.92387953251128675613f; break;
    case 3: v = 0.83146961230254523708f; break;
    case 4: v = 0.70710678118654752440f; break;
    case 5: v = 0.55557023301960222474f; break;
    case 6: v = 0.38268343236508977173f; break;
    case 7: v = 0.19509032201612826785f; break;
    default: v = 0.0f; break;
  }
  return neg ? -v : v;
}
__host__ __device__ constexpr float sin32(int n) { return cos32(n - 8); }

// h[m] = sum_q x[q] exp(+2 pi i q m / 32), natural order in and out, in place.  32 = 8 x 4: q = 4 q1 + q0, m = m0 + 8 m1.
__device__ __forceinline__ void idft32(float2 (&x)[32]) {
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) {
    float2 v[8];
#pragma unroll
    for (int q1 = 0; q1 < 8; ++q1) v[q1] = x[4 * q1 + q0];
    dft8<true>(v);
#pragma unroll
    for (int m0 = 0; m0 < 8; ++m0) {
      float2 y = v[m0];
      if (q0 * m0 != 0) {
        const float c = cos32(q0 * m0), s = sin32(q0 * m0);   // exp(+2 pi i q0 m0 / 32)
        y = make_float2(y.x * c - y.y * s, y.x * s + y.y * c);
      }
      x[4 * m0 + q0] = y;   // Y[q0][m0] parked at 4 m0 + q0
    }
  }
  float2 h[32];
#pragma unroll
  for (int m0 = 0; m0 < 8; ++m0) {
    float2 u0 = x[4 * m0], u1 = x[4 * m0 + 1], u2 = x[4 * m0 + 2], u3 = x[4 * m0 + 3];
    dft4<true>(u0, u1, u2, u3);
    h[m0] = u0; h[m0 + 8] = u1; h[m0 + 16] = u2; h[m0 + 24] = u3;
  }
#pragma unroll
  for (int m = 0; m < 32; ++m) x[m] = h[m];
}

// float2 entries of the lag-phase table appended for N = 1024: T1[i][s] = weight_s exp(2 pi i s i / N) and
// T2[i][s] = exp(2 pi i (32 s + 16) i / N), i = 0..31, s = 0..15 (weight 1/2 for s = 0, 1 otherwise).  The kernel only reads rows
// i = 0, 1 and 16: the other lags are reached by the recurrence t_{i+1} = t_i t_1 in registers (re-seeded at i = 16), because the
// table is per lane and reading it for every lag was 31 % of the kernel's shared-memory wavefronts (ncu r2b_cfg2_warp).
constexpr int kLagTabLen = 2 * 32 * 16;

// One half-warp = one pair.  s_U: whitened spectra [M][KP]; s_Uf: the bins k = 32 q + 16 of every channel, [M][16] (a compact copy: read
// straight from s_U they sit 256 bytes apart, a 16-way bank conflict); s_lag: the table above; scratch: 2 KB of this half-warp
// (16 rows x 128 B).  LCAP (28 or 31): compile-time bound of max_lag.
template <int LCAP>
__device__ __forceinline__ void tdoa_pair_halfwarp(const float2 *s_U, const float2 *s_Uf, const float2 *s_lag, float *scratch, int mi, int mj, bool live,
                                                   int p, int max_lag, float *__restrict__ curves_ft, int32_t *__restrict__ lags_ft,
                                                   float *__restrict__ peaks_ft) {
  constexpr int N = 1024, KP = spec_pitch(N);
  const int s = threadIdx.x & 15;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);
  const float2 *Ui = s_U + (size_t)mi * KP, *Uj = s_U + (size_t)mj * KP;
  // ---- sub-sequence s of the Hermitian-extended cross-spectrum
  float2 x[32];
#pragma unroll
  for (int q = 0; q < 16; ++q) x[q] = cmulc(Ui[32 * q + s], Uj[32 * q + s]);                 // k = 32 q + s <= 511
#pragma unroll
  for (int q = 16; q < 32; ++q) x[q] = cmulc(Uj[32 * (32 - q) - s], Ui[32 * (32 - q) - s]);  // k > 511: conj(G[N - k])
  if (s == 0) { x[0].x *= 2.f; x[0].y = 0.f; x[16].x *= 2.f; x[16].y = 0.f; }                 // G[0], G[N/2]: real, doubled (end-point terms)
  const float2 gf = cmulc(s_Uf[mi * 16 + s], s_Uf[mj * 16 + s]);                              // this lane's bin 32 s + 16 of sub-sequence 16
  idft32(x);
  // ---- lag terms of this lane, two passes of 16 |lags| (both signs per pass): pos[i] -> l = +i (m = i), neg[i] -> l = -i (m = 32 - i)
  const int L = 2 * max_lag + 1;
  const float2 r1 = s_lag[16 + s], r2 = s_lag[512 + 16 + s];   // per-lag rotations exp(2 pi i s / N), exp(2 pi i (32 s + 16) / N)
  float best = -3.0e38f; int besti = 0x7fffffff;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    float2 t1 = s_lag[pass * 256 + s], t2 = s_lag[512 + pass * 256 + s];   // exact at i = 0 and i = 16
    float v[32];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = 16 * pass + u;
      if (i > LCAP) { v[u] = 0.f; v[16 + u] = 0.f; continue; }
      const float a = t2.x * gf.x, b_ = t2.y * gf.y;
      v[u] = fmaf(t1.x, x[i].x, fmaf(-t1.y, x[i].y, a - b_));
      v[16 + u] = (i == 0) ? 0.f : fmaf(t1.x, x[(32 - i) & 31].x, fmaf(t1.y, x[(32 - i) & 31].y, a + b_));
      if (u < 15 && i < LCAP) { t1 = cmul(t1, r1); t2 = cmul(t2, r2); }
    }
    // transpose-and-sum through the scratch: row = lane (128 B), 16-byte chunk c at position c ^ (lane & 7)
    float4 *row = reinterpret_cast<float4 *>(scratch + s * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) row[c ^ (s & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp(hmask);
    float2 acc = make_float2(0.f, 0.f);   // elements 2 s and 2 s + 1 of this pass: lanes 0..7 positive lags, lanes 8..15 negative ones
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float2 a = *reinterpret_cast<const float2 *>(scratch + r * 32 + (((s >> 1) ^ (r & 7)) << 2) + ((s & 1) << 1));
      acc = cadd(acc, a);
    }
    __syncwarp(hmask);
    // ---- candidates of this lane (first maximum: the lowest window index c = l + max_lag wins ties)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 16 * pass + 2 * (s & 7) + e;
      const int l = (s < 8) ? i : -i;
      const bool ok = i <= max_lag && (s < 8 || i >= 1);
      const float val = e ? acc.y : acc.x;
      if (ok) {
        const int c = l + max_lag;
        if (curves_ft && live) curves_ft[(size_t)p * L + c] = val;
        if (val > best || (val == best && c < besti)) { best = val; besti = c; }
      }
    }
  }
  const unsigned key = float_order_key(best + 0.f);
  const unsigned kmax = __reduce_max_sync(hmask, key);
  const unsigned imin = __reduce_min_sync(hmask, key == kmax ? (unsigned)besti : 0x7fffffffu);
  if (s == 0 && live) {
    lags_ft[p] = (int)imin - max_lag;
    if (peaks_ft) peaks_ft[p] = float_from_order_key(kmax);
  }
}

}  // namespace mcag

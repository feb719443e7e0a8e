// Register-resident FFT engine for the N = 512 and N = 1024 real transforms of the STFT (BASELINE configs 1 and 3 - 5).
//
// The shared-memory Stockham engine of fft.cuh moves every point of a 256-point packed-complex transform through shared memory three
// times (three radix passes) and once more for the real post-processing: 197 shared-memory wavefronts per frame, and ncu shows the STFT
// and the fused masking kernel bound by exactly that (r2_cfg1l_kernels: l1tex data-pipe 76 % of peak at 39 % of the DRAM peak).  Here a
// transform belongs to SIXTEEN lanes holding sixteen points each (thirty-two for N = 1024): 256 = 16 x 16 (512 = 32 x 16), so both
// passes run in registers (all their twiddles are immediates), the only exchange is ONE swizzled transpose through shared memory, and
// the real post-processing pairs bin k with bin N/2 - k, which live in lane c and lane 16 - c of the same half-warp: a shuffle, not a
// round trip.  A warp carries two transforms (two frames of a row); nothing wider than __syncwarp is needed.
//
//   forward   lane b holds z[16 a + b] (a = register) -> dft16 / dft32 over a -> * W^(b c) -> transpose -> dft16 over b -> Z[c + 16 e] in lane c
//   inverse   the same schedule with conjugated twiddles takes Z[c + 16 e] in lane c back to z[16 a + b] in lane b (unscaled)
//
// Checked against numpy in tools/proto/fft16_halfwarp.py (index maps, lane pairing, bank-conflict freedom of the transpose, and the inverse
// pre-processing).  Used by stft_hw_kernel (stft.cu).  A one-warp-per-stream FastBinauralMasking kernel on this engine (left / right channel in
// the two half-warps, spectra held in registers from analysis to synthesis) was built and measured in round 2: 640 shared-memory
// wavefronts per stereo frame against 1090, but 1.56 ms per cfg1m step against 1.29 ms for mask_fused_kernel - one warp per stream is
// 14 warps per SM running a 4300-instruction loop body (68 KB of SASS, past the 32 KB L1.5 instruction cache: the no-instruction stall
// equalled the issue rate) and the per-frame dependency chain no longer hides behind other warps.  It was removed again.
#pragma once
#include "fft.cuh"
#include "tdoa_warp.cuh"   // cos32 / sin32 literals

namespace mcag {

// X[k] = sum_n x[n] exp(-+2 pi i n k / 16), natural order in and out, in place; 16 = 4 x 4: n = 4 n1 + n2, k = k1 + 4 k2
template <bool INV> __device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);   // v[4 k1 + n2] = A[n2][k1]
#pragma unroll
  for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
    for (int n2 = 1; n2 < 4; ++n2) {
      const int e = k1 * n2;   // W16^e
      float2 &y = v[4 * k1 + n2];
      if (e == 4) {
        y = rot90<INV>(y);
      } else {
        const float c = cos32(2 * e), s = INV ? sin32(2 * e) : -sin32(2 * e);
        y = make_float2(y.x * c - y.y * s, y.x * s + y.y * c);
      }
    }
  float2 o[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // v[4 k1 + k2] = X[k1 + 4 k2]
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) o[k1 + 4 * k2] = v[4 * k1 + k2];
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = o[k];
}

// cos / sin (2 pi n / 64) as literals (the 32-point set of tdoa_warp.cuh plus the odd multiples)
__host__ __device__ constexpr float cos64(int n) {
  n &= 63;
  if ((n & 1) == 0) return cos32(n >> 1);
  if (n > 32) n = 64 - n;
  const bool neg = n > 16;
  if (neg) n = 32 - n;
  float v = 0.f;
  switch (n) {
    case 1: v = 0.99518472667219688624f; break;
    case 3: v = 0.95694033573220886494f; break;
    case 5: v = 0.88192126434835502971f; break;
    case 7: v = 0.77301045336273696081f; break;
    case 9: v = 0.63439328416364549822f; break;
    case 11: v = 0.47139673682599764856f; break;
    case 13: v = 0.29028467725446236764f; break;
    default: v = 0.09801714032956060199f; break;   // 15
  }
  return neg ? -v : v;
}
__host__ __device__ constexpr float sin64(int n) { return cos64(n - 16); }

// X[k] = sum_n x[n] exp(-+2 pi i n k / 32), natural order in and out, in place; 32 = 8 x 4: n = 4 n1 + n0, k = k0 + 8 k1
template <bool INV> __device__ __forceinline__ void dft32(float2 (&x)[32]) {
#pragma unroll
  for (int n0 = 0; n0 < 4; ++n0) {
    float2 v[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) v[n1] = x[4 * n1 + n0];
    dft8<INV>(v);
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0) {
      float2 y = v[k0];
      if (n0 * k0 != 0) {
        const float c = cos32(n0 * k0), s = INV ? sin32(n0 * k0) : -sin32(n0 * k0);
        y = make_float2(y.x * c - y.y * s, y.x * s + y.y * c);
      }
      x[4 * k0 + n0] = y;   // Y[n0][k0] parked at 4 k0 + n0
    }
  }
  float2 h[32];
#pragma unroll
  for (int k0 = 0; k0 < 8; ++k0) {
    float2 u0 = x[4 * k0], u1 = x[4 * k0 + 1], u2 = x[4 * k0 + 2], u3 = x[4 * k0 + 3];
    dft4<INV>(u0, u1, u2, u3);
    h[k0] = u0; h[k0 + 8] = u1; h[k0 + 16] = u2; h[k0 + 24] = u3;
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] = h[k];
}

// R = points per lane: 16 (256-point packed-complex transform, N = 512) or 32 (512 points, N = 1024); sixteen lanes per transform
template <int R> constexpr int fft16_tab_len() { return R * 16; }   // float2 entries of the inter-pass table: t1[c * 16 + b] = exp(-2 pi i b c / (16 R))
template <int R> constexpr int fft16_buf_len() { return R * 16; }   // float2 entries of one transform's transpose buffer

// builds the inter-pass table from the table of mcag_k_twiddles for N = 32 R (tw[n] = exp(-2 pi i n / N), n < N/2); the caller syncs
template <int R> __device__ __forceinline__ void fft16_load_table(float2 *s_t1, const float2 *__restrict__ tw, int tid, int nthreads) {
  for (int i = tid; i < R * 16; i += nthreads) s_t1[i] = tw_lookup<false>(tw, 2 * (i >> 4) * (i & 15), 16 * R);
}

// 16 R-point complex transform of one half-warp (l16 = lane & 15), in place in the registers: on entry v[a] = z[16 a + l16], on exit
// v[e] = Z[l16 + 16 e] (e < R).  xbuf: the transpose buffer of this half-warp (16-byte aligned, R x 16 float2); element (row c, col b)
// sits at c * 16 + (b ^ (c & 15)), which keeps the row-wise stores and the column-wise loads of a half-warp on sixteen distinct bank
// pairs.  Both half-warps of the warp must call this together.
template <int R, bool INV> __device__ __forceinline__ void fft_hw(float2 (&v)[R], float2 *xbuf, const float2 *s_t1, int l16) {
  static_assert(R == 16 || R == 32, "16 or 32 points per lane");
  if constexpr (R == 16) dft16<INV>(v); else dft32<INV>(v);
#pragma unroll
  for (int c = 1; c < R; ++c) {
    float2 w = s_t1[c * 16 + l16];
    if (INV) w.y = -w.y;
    v[c] = cmul(v[c], w);
  }
#pragma unroll
  for (int c = 0; c < R; ++c) xbuf[c * 16 + (l16 ^ (c & 15))] = v[c];
  __syncwarp();
#pragma unroll
  for (int h = 0; h < R / 16; ++h) {   // row l16 + 16 h: outputs k = l16 + 16 h + R d = l16 + 16 (h + (R / 16) d)
    float2 u[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) u[b] = xbuf[(l16 + 16 * h) * 16 + (b ^ l16)];
    dft16<INV>(u);
#pragma unroll
    for (int d = 0; d < 16; ++d) v[(R / 16) * d + h] = u[d];
  }
  __syncwarp();   // the buffer may be rewritten by the next transform
}

// Real post-processing of an N = 32 R sample frame packed as z[n] = x[2n] + i x[2n+1]: on entry v[e] = Z[c + 16 e] (c = l16); emit(e, X) is
// called with X[c + 16 e] for e = 0..R-1, and nyq = X[N/2] (real; meaningful on lane c = 0, whose X[0] has a zero imaginary part).
// wl = exp(-2 pi i c / N).  Bin k pairs with N/2 - k = (16 - c) + 16 (R - 1 - e): lane 16 - c, register R - 1 - e; lane 0 pairs with
// itself, register (R - e) mod R.  X[k] = (Z[k] + conj Z[N/2-k]) / 2 + W^k (-i/2) (Z[k] - conj Z[N/2-k]), the -i/2 folded into the twiddle.
template <int R, class Emit> __device__ __forceinline__ void fft_hw_real_post(const float2 (&v)[R], float2 wl, int l16, float &nyq, Emit emit) {
  const int src = ((16 - l16) & 15) | ((int)threadIdx.x & 16);
  float2 r[R];
#pragma unroll
  for (int i = 0; i < R; ++i) { r[i].x = __shfl_sync(0xffffffffu, v[i].x, src); r[i].y = __shfl_sync(0xffffffffu, v[i].y, src); }
  float2 pz[R];   // the partners first: emit may overwrite v[e]
#pragma unroll
  for (int e = 0; e < R; ++e) pz[e] = l16 == 0 ? v[(R - e) & (R - 1)] : r[R - 1 - e];
  nyq = v[0].x - v[0].y;
  const float2 wh = make_float2(0.5f * wl.y, -0.5f * wl.x);   // -i/2 wl
#pragma unroll
  for (int e = 0; e < R; ++e) {
    const float2 zk = v[e];
    const float2 sm = make_float2(zk.x + pz[e].x, zk.y - pz[e].y), df = make_float2(zk.x - pz[e].x, zk.y + pz[e].y);
    // -i/2 exp(-2 pi i (c + 16 e) / N): 16 e / N = e / (2 R)
    const float2 w = e == 0 ? wh : cmul(wh, R == 16 ? make_float2(cos32(e), -sin32(e)) : make_float2(cos64(e), -sin64(e)));
    float2 X = make_float2(fmaf(w.x, df.x, fmaf(-w.y, df.y, 0.5f * sm.x)), fmaf(w.x, df.y, fmaf(w.y, df.x, 0.5f * sm.y)));
    if (e == 0 && l16 == 0) X.y = 0.f;
    emit(e, X);
  }
}
// sum over the sixteen lanes of a half-warp, every lane gets it
__device__ __forceinline__ float hw_sum(float x) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

}  // namespace mcag

// Register-resident FFT engine for the N = 512 real transforms (BASELINE config 1: 16 kHz stereo, 512-sample frames).
//
// The shared-memory Stockham engine of fft.cuh moves every point of a 256-point packed-complex transform through shared memory three
// times (three radix passes) and once more for the real post-processing: 197 shared-memory wavefronts per frame, and ncu shows the STFT
// and the fused masking kernel bound by exactly that (r2_cfg1l_kernels: l1tex data-pipe 76 % of peak at 39 % of the DRAM peak).  Here a
// transform belongs to SIXTEEN lanes holding sixteen points each: 256 = 16 x 16, so both radix-16 passes run in registers (all their
// twiddles are immediates), the only exchange is ONE swizzled 16 x 16 transpose through shared memory, and the real post-processing
// pairs bin k with bin 256 - k, which live in lane c and lane 16 - c of the same half-warp: a shuffle, not a round trip.  A warp carries
// two transforms (two frames of a row); nothing wider than __syncwarp is needed.
//
//   forward   lane b holds z[16 a + b] (a = register)  -> dft16 over a -> * W256^(b c) -> transpose -> dft16 over b -> Z[c + 16 d] in lane c
//   inverse   the same schedule with conjugated twiddles takes Z[c + 16 d] in lane c back to z[16 a + b] in lane b (unscaled)
//
// Checked against numpy in tools/proto/fft16_halfwarp.py (index maps, lane pairing, bank-conflict freedom of the transpose, and the inverse
// pre-processing).  Used by stft512_hw_kernel.  A one-warp-per-stream FastBinauralMasking kernel on this engine (left / right channel in
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

constexpr int kFft16TabLen = 256;   // float2 entries of the inter-pass table: t1[c * 16 + b] = exp(-2 pi i b c / 256)
constexpr int kFft16BufLen = 256;   // float2 entries of one transform's transpose buffer

// builds the inter-pass table from the N = 512 table of mcag_k_twiddles (tw[n] = exp(-2 pi i n / 512), n < 256); the caller syncs
__device__ __forceinline__ void fft16_load_table(float2 *s_t1, const float2 *__restrict__ tw512, int tid, int nthreads) {
  for (int i = tid; i < kFft16TabLen; i += nthreads) s_t1[i] = tw_lookup<false>(tw512, 2 * (i >> 4) * (i & 15), 256);
}

// 256-point complex transform of one half-warp (l16 = lane & 15), in place in the registers as described above.  xbuf: the 2 KB transpose
// buffer of this half-warp (16-byte aligned); element (row, col) sits at row * 16 + (col ^ row), which keeps the row-wise stores and the
// column-wise loads of a half-warp on sixteen distinct bank pairs.  Both half-warps of the warp must call this together.
template <bool INV> __device__ __forceinline__ void fft256_hw(float2 (&v)[16], float2 *xbuf, const float2 *s_t1, int l16) {
  dft16<INV>(v);
#pragma unroll
  for (int c = 1; c < 16; ++c) {
    float2 w = s_t1[c * 16 + l16];   // the same address in both half-warps: one wavefront per warp
    if (INV) w.y = -w.y;
    v[c] = cmul(v[c], w);
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) xbuf[c * 16 + (l16 ^ c)] = v[c];
  __syncwarp();
#pragma unroll
  for (int b = 0; b < 16; ++b) v[b] = xbuf[l16 * 16 + (b ^ l16)];
  __syncwarp();   // the buffer may be rewritten by the next transform
  dft16<INV>(v);
}

// Real post-processing of a 512-sample frame packed as z[n] = x[2n] + i x[2n+1]: on entry v[d] = Z[c + 16 d] (c = l16); emit(d, X) is called
// with X[c + 16 d] for d = 0..15, and nyq = X[256] (real; meaningful on lane c = 0, whose X[0] has a zero imaginary part).
// wl = exp(-2 pi i c / 512).  Bin k pairs with 256 - k = (16 - c) + 16 (15 - d): lane 16 - c, register 15 - d; lane 0 pairs with itself,
// register (16 - d) & 15.  X[k] = (Z[k] + conj Z[256-k]) / 2 + W^k (-i/2) (Z[k] - conj Z[256-k]), the -i/2 folded into the twiddle.
template <class Emit> __device__ __forceinline__ void fft16_real_post(const float2 (&v)[16], float2 wl, int l16, float &nyq, Emit emit) {
  const int src = ((16 - l16) & 15) | ((int)threadIdx.x & 16);
  float2 r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { r[i].x = __shfl_sync(0xffffffffu, v[i].x, src); r[i].y = __shfl_sync(0xffffffffu, v[i].y, src); }
  nyq = v[0].x - v[0].y;
  const float2 wh = make_float2(0.5f * wl.y, -0.5f * wl.x);   // -i/2 wl
  float2 pz[16];   // the partners first: emit may overwrite v[d]
#pragma unroll
  for (int d = 0; d < 16; ++d) pz[d] = l16 == 0 ? v[(16 - d) & 15] : r[15 - d];
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    const float2 zk = v[d];
    const float2 sm = make_float2(zk.x + pz[d].x, zk.y - pz[d].y), df = make_float2(zk.x - pz[d].x, zk.y + pz[d].y);
    const float2 w = d == 0 ? wh : cmul(wh, make_float2(cos32(d), -sin32(d)));   // -i/2 exp(-2 pi i (c + 16 d) / 512)
    float2 X = make_float2(fmaf(w.x, df.x, fmaf(-w.y, df.y, 0.5f * sm.x)), fmaf(w.x, df.y, fmaf(w.y, df.x, 0.5f * sm.y)));
    if (d == 0 && l16 == 0) X.y = 0.f;
    emit(d, X);
  }
}

// sum over the sixteen lanes of a half-warp, every lane gets it
__device__ __forceinline__ float hw_sum(float x) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

}  // namespace mcag

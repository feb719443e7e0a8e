// Per-band step of FastBinauralMasking shared by the staged scan kernel (mask.cu) and the fused kernel (mask_fused.cu): the
// first-order power tracker, the temporal / spatial decisions and the gains of one (frame, band), FastBinauralMasking.cpp:159-313,
// 369-376, 437-459, 477-493 with the constants of FastBinauralMasking.h:111-128.  method / alg enums: ArrayModules.h:81,89.
#pragma once
#include "common.cuh"

namespace mcag {

constexpr int MS_NSTAT = 6;   // pw2, num, eL, eR, pL, pR:
//   st[0] = sum_{k<N/2} H2 |(L+R)/2|^2      (getFramePower :496-538)
//   st[1] = sum_{k<K}   H2 Re(R conj L)     (normaliseFFTCorrelation :437-441)
//   st[2], st[3] = sum_{k<K} H2 |L|^2, |R|^2      (:446-452, maskFrameByScaling :262-264)
//   st[4], st[5] = sum_{k<N/2} H2 |L|^2, |R|^2    (noisyFrame -> getPower :222,520-538)

// Updates Q (and noise on the first frame), returns the decision (2 = spatial mask, 1 = temporal mask, 0 = pass) and the gains of the
// left / right band signals.  first_call = frames processed before this one (the reference's _firstCall counter, saturated at 2).
__device__ __forceinline__ int mask_band_step(const float *st, int N, int method, int alg, float thr, int first_call, float &Q, float &noise,
                                              float &gl, float &gr) {
  const float K = (float)(N / 2 + 1), NH = (float)(N / 2);
  const float lam = 0.04f, keep = 1.0f - 0.04f, reject = 0.999f, rho = 0.01f;   // FastBinauralMasking.h:114,122,126
  const float pw = sqrtf(st[0] / NH);
  Q = Q * lam + keep * pw;                                        // temportalMasking :489
  bool temp = pw < reject * Q;                                    // :492
  bool spat = false;
  if (alg == 0 || alg == 1) {                                     // BOTH / SPATIAL :159-166
    const float num = st[1] / K;
    float ncorr;
    if (num == 0.f) ncorr = 0.f;
    else { const float den = sqrtf((st[2] / K) * (st[3] / K)); ncorr = (den == 0.f) ? 1.f : num / den; }
    spat = ncorr < thr;                                           // :374
    if (alg == 1) temp = false;
  }
  gl = 1.f; gr = 1.f;                                             // enhanceFrame: _enhanceFactor = 1
  const int dec = spat ? 2 : (temp ? 1 : 0);
  if (dec) {
    switch (method) {
      case 3: gl = gr = 1.0f / 1000.0f; break;                    // FULL  :214-217
      case 0: gl = gr = 1.0f / (spat ? 10.0f : 3.0f); break;      // FACTOR :284-287, .h:117-118
      case 1: {                                                   // RELATIVE :246-282 (uses the updated Q)
        if (Q < 1e-10f) gl = gr = sqrtf(rho);
        else { gl = sqrtf((st[2] / K) * rho / Q); gr = sqrtf((st[3] / K) * rho / Q); }
      } break;
      case 4: {                                                   // NOISY :220-243
        if (first_call >= 2) {
          const float pl = sqrtf(st[4] / NH), pr = sqrtf(st[5] / NH);
          gl = pl > 0.f ? noise / pl : 1.f;
          gr = pr > 0.f ? noise / pr : 1.f;
        }
      } break;
      default: break;
    }
  }
  if (first_call + 1 < 2) noise = Q;                              // :193-197: the first frame snapshots the noise estimate
  return dec;
}

}  // namespace mcag

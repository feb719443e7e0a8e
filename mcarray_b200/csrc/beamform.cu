// K4 ds_beamform: Beamformer::processFrame (Beamformer.cpp:51-71),
//   Y[k] = (1/M) sum_c X_c[k] exp(j k phi_c),  phi_c = 2 pi fs/N/c * x_c * cos(DOA + pi/2).
// The per-bin phase increment phi_c/(2 pi) arrives as 0.64 fixed-point turns (computed in double on the host from the
// float-rounded grid angle the reference would use), so k*phi_c wraps exactly.
//   k_ds_select : steer each frame to the S cells picked by selectDOA (BeamformingSeparationAndLocalisation.cpp:103-119)
//   k_ds_fan    : steer every frame to all D directions (BASELINE config 3), CUDA-core tile version
#include "common.cuh"
#include "kernels.h"

namespace mcag {

// steering table tab[d][c][k] = exp(j 2 pi k fx[d][c]) for the selected-cell beamformer (D*M*KP float2, L2 resident)
__global__ void steer_table_kernel(const uint64_t *__restrict__ fx, int DM, int N, float2 *__restrict__ tab) {
  const int KP = spec_pitch(N), K = N / 2 + 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)DM * KP) return;
  const int k = (int)(i % KP);
  tab[i] = (k < K) ? phase_ramp(fx[i / KP], k) : make_float2(0.f, 0.f);
}
int k_steer_table(const uint64_t *fx, int DM, int N, float2 *tab, cudaStream_t st) {
  const long long n = (long long)DM * spec_pitch(N);
  steer_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fx, DM, N, tab);
  MCAG_CHECK_LAUNCH();
  return 0;
}

// one CTA per frame; out[b][t][s][k] for s < S, zeros for S <= s < C_out
__global__ void __launch_bounds__(256) ds_select_kernel(const float2 *__restrict__ spec, int M, int N, const float2 *__restrict__ tab,
                                                         const int32_t *__restrict__ cells, int S, int C_out, float2 *__restrict__ out) {
  const int KP = spec_pitch(N), K = N / 2 + 1;
  const long long bt = blockIdx.x;
  const float2 *X = spec + bt * M * KP;
  const float invM = 1.0f / (float)M;
  for (int s = 0; s < C_out; ++s) {
    float2 *o = out + (bt * C_out + s) * KP;
    if (s >= S || s >= M) {
      for (int k = threadIdx.x; k < KP; k += blockDim.x) o[k] = make_float2(0.f, 0.f);
      continue;
    }
    const float2 *A = tab + (size_t)cells[bt * S + s] * M * KP;
    for (int k = threadIdx.x; k < KP; k += blockDim.x) {
      float2 acc = make_float2(0.f, 0.f);
      if (k < K)
        for (int c = 0; c < M; ++c) acc = cadd(acc, cmul(X[(size_t)c * KP + k], A[(size_t)c * KP + k]));   // channel order as :56-68
      o[k] = make_float2(acc.x * invM, acc.y * invM);
    }
  }
}

int k_ds_select(const float2 *spec, int B, int T, int M, int N, const float2 *steer_tab, const int32_t *cells, int S, int C_out, float2 *out,
                cudaStream_t st) {
  const long long BT = (long long)B * T;
  if (BT <= 0) return 0;
  ds_select_kernel<<<(unsigned)BT, 256, 0, st>>>(spec, M, N, steer_tab, cells, S, C_out, out);
  MCAG_CHECK_LAUNCH();
  return 0;
}

// fan: CTA = (8-frame tile, 32-bin chunk, stream).  The spectra tile [8 frames][M][32 bins] is staged in shared memory; a
// warp owns 4 directions at a time and each lane one bin, i.e. a register tile of 4 directions x 8 frames per thread (32 complex
// accumulators): per microphone the 4 steering phasors are generated once (32-bit fixed-point phase, exact wrap, MUFU sin / cos)
// and reused for the 8 frames, the 8 spectrum values are read once and reused for the 4 directions.  Channels are added in
// index order like Beamformer.cpp:56-68.  Output rows are written 256 bytes at a time (32 consecutive bins).
// LOADED = true is the filter-and-sum beamformer: the per-bin complex weights W[d][c][k] come from memory instead of being generated as
// delay phasors, Y[t][d][k] = (1/M) sum_c X_c[t][k] W[d][c][k] (same channel order, same 1/M): with W = exp(j k phi_c(d)) it IS the
// delay-and-sum fan.  The reference has no filter-and-sum class; BASELINE.json's north_star names it next to delay-and-sum.
constexpr int FAN_TF = 8, FAN_TD = 4, FAN_KC = 32;
template <bool LOADED>
__global__ void __launch_bounds__(256, 2) ds_fan_kernel(const float2 *__restrict__ spec, int T, int M, int N, const uint64_t *__restrict__ fx,
                                                         const float2 *__restrict__ W, int D, float2 *__restrict__ out, int KPo) {
  extern __shared__ float2 s_X[];   // [FAN_TF][M][FAN_KC]
  const int KP = spec_pitch(N), K = N / 2 + 1;   // KPo >= KP: pitch of the output rows (bins K .. KPo - 1 are written as zeros)
  const int t0 = blockIdx.x * FAN_TF, k0 = blockIdx.y * FAN_KC, b = blockIdx.z;
  for (int i = threadIdx.x; i < FAN_TF * M * FAN_KC; i += blockDim.x) {
    const int kk = i % FAN_KC, c = (i / FAN_KC) % M, f = i / (FAN_KC * M);
    const int t = t0 + f, k = k0 + kk;
    s_X[i] = (t < T && k < K) ? spec[(((long long)b * T + t) * M + c) * KP + k] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int k = k0 + lane;
  if (k >= KPo) return;   // no block-level synchronisation below
  const float invM = 1.0f / (float)M;
  for (int d0 = warp * FAN_TD; d0 < D; d0 += nwarp * FAN_TD) {
    float2 acc[FAN_TD][FAN_TF];
#pragma unroll
    for (int i = 0; i < FAN_TD; ++i)
#pragma unroll
      for (int f = 0; f < FAN_TF; ++f) acc[i][f] = make_float2(0.f, 0.f);
    for (int c = 0; c < M; ++c) {
      float2 a[FAN_TD];
#pragma unroll
      for (int i = 0; i < FAN_TD; ++i) {
        if constexpr (LOADED) {
          a[i] = __ldg(W + ((size_t)min(d0 + i, D - 1) * M + c) * KP + min(k, KP - 1));                         // lanes = consecutive bins: 256-byte rows
        } else {
          const uint64_t f64 = __ldg(fx + (size_t)min(d0 + i, D - 1) * M + c);
          const int32_t ph = (int32_t)((uint32_t)((f64 + 0x80000000ull) >> 32) * (uint32_t)k);   // signed turns * 2^32, wraps exactly
          __sincosf((float)ph * 1.4629180792671596e-09f, &a[i].y, &a[i].x);                      // 2 pi / 2^32
        }
      }
#pragma unroll
      for (int f = 0; f < FAN_TF; ++f) {
        const float2 xv = s_X[(f * M + c) * FAN_KC + lane];
#pragma unroll
        for (int i = 0; i < FAN_TD; ++i) {
          acc[i][f].x = fmaf(xv.x, a[i].x, fmaf(-xv.y, a[i].y, acc[i][f].x));
          acc[i][f].y = fmaf(xv.x, a[i].y, fmaf(xv.y, a[i].x, acc[i][f].y));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < FAN_TD; ++i) {
      if (d0 + i >= D) break;
#pragma unroll
      for (int f = 0; f < FAN_TF; ++f)
        if (t0 + f < T)
          out[(((long long)b * T + t0 + f) * D + d0 + i) * KPo + k] = (k < K) ? make_float2(acc[i][f].x * invM, acc[i][f].y * invM) : make_float2(0.f, 0.f);
    }
  }
}

int k_ds_fan(const float2 *spec, int B, int T, int M, int N, const uint64_t *steer_fx, int D, float2 *out, cudaStream_t st, int out_pitch) {
  if (B <= 0 || T <= 0) return 0;
  const int KP = out_pitch > 0 ? out_pitch : spec_pitch(N);
  if (KP < spec_pitch(N)) return mcag_set_error(1, "ds_fan: output pitch below the spectrum pitch");
  size_t smem = sizeof(float2) * FAN_TF * M * FAN_KC;
  if (smem > 200 * 1024) return mcag_set_error(1, "ds_fan: too many channels");
  cudaFuncSetAttribute(ds_fan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((T + FAN_TF - 1) / FAN_TF, (KP + FAN_KC - 1) / FAN_KC, B);
  ds_fan_kernel<false><<<grid, 256, smem, st>>>(spec, T, M, N, steer_fx, nullptr, D, out, KP);
  MCAG_CHECK_LAUNCH();
  return 0;
}

// filter-and-sum fan: weights [D][M][KP] float2 (pad bins ignored)
int k_fs_fan(const float2 *spec, int B, int T, int M, int N, const float2 *weights, int D, float2 *out, cudaStream_t st, int out_pitch) {
  if (B <= 0 || T <= 0) return 0;
  const int KP = out_pitch > 0 ? out_pitch : spec_pitch(N);
  if (KP < spec_pitch(N)) return mcag_set_error(1, "fs_fan: output pitch below the spectrum pitch");
  size_t smem = sizeof(float2) * FAN_TF * M * FAN_KC;
  if (smem > 200 * 1024) return mcag_set_error(1, "fs_fan: too many channels");
  cudaFuncSetAttribute(ds_fan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((T + FAN_TF - 1) / FAN_TF, (KP + FAN_KC - 1) / FAN_KC, B);
  ds_fan_kernel<true><<<grid, 256, smem, st>>>(spec, T, M, N, nullptr, weights, D, out, KP);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

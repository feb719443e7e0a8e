// K3 srp_contract, CUDA-core tile version (the tcgen05 version lives in srp_tc.cu): channel form of the SRP-PHAT pair
// sum of SteeringBeamforming::computeCorrelations (SteeringBeamforming.cpp:104-130), SURVEY.md §8a row A4:
//   sum_{i<j} Re(G_ij e^{+j w tau_ij(d)}) = 1/2 ( |sum_m U_m e^{-j w tau_m(d)}|^2 - nz ),  U = X/|X|,  tau_ij = tau_j - tau_i,
// nz = number of non-zero channels in the bin.  srp[b][t][d] = sum over all K one-sided bins.
#include "common.cuh"
#include "kernels.h"

namespace mcag {

constexpr int SRP_TF = 8, SRP_KC = 32;

// CTA = (8-frame tile, direction tile of blockDim/32 * DPW directions, stream); lane = bin within the 32-bin chunk.
template <int DPW>
__global__ void __launch_bounds__(256) srp_channel_kernel(const float2 *__restrict__ spec, int T, int M, int N, const uint64_t *__restrict__ mic_fx,
                                                           int D, float *__restrict__ srp) {
  extern __shared__ float2 s_U[];   // [SRP_TF][M][SRP_KC], then nz [SRP_TF][SRP_KC] floats
  float *s_nz = reinterpret_cast<float *>(s_U + SRP_TF * M * SRP_KC);
  const int KP = spec_pitch(N), K = N / 2 + 1;
  const int t0 = blockIdx.x * SRP_TF, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int d0 = (blockIdx.y * nwarp + warp) * DPW;
  float acc[DPW][SRP_TF];
#pragma unroll
  for (int q = 0; q < DPW; ++q)
#pragma unroll
    for (int f = 0; f < SRP_TF; ++f) acc[q][f] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SRP_KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < SRP_TF * M * SRP_KC; i += blockDim.x) {
      const int kk = i % SRP_KC, c = (i / SRP_KC) % M, f = i / (SRP_KC * M);
      const int t = t0 + f, k = k0 + kk;
      s_U[i] = (t < T && k < K) ? whiten(spec[(((long long)b * T + t) * M + c) * KP + k]) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SRP_TF * SRP_KC; i += blockDim.x) {
      const int kk = i % SRP_KC, f = i / SRP_KC;
      float n = 0.f;
      for (int c = 0; c < M; ++c) { float2 u = s_U[(f * M + c) * SRP_KC + kk]; n += (u.x != 0.f || u.y != 0.f) ? 1.f : 0.f; }
      s_nz[i] = n;
    }
    __syncthreads();
    const int k = k0 + lane;
#pragma unroll
    for (int q = 0; q < DPW; ++q) {
      const int d = d0 + q;
      if (d >= D) continue;
      float2 y[SRP_TF];
#pragma unroll
      for (int f = 0; f < SRP_TF; ++f) y[f] = make_float2(0.f, 0.f);
      for (int c = 0; c < M; ++c) {
        const float2 a = phase_ramp(mic_fx[(size_t)d * M + c], k);
#pragma unroll
        for (int f = 0; f < SRP_TF; ++f) y[f] = cadd(y[f], cmul(s_U[(f * M + c) * SRP_KC + lane], a));
      }
#pragma unroll
      for (int f = 0; f < SRP_TF; ++f) acc[q][f] += 0.5f * (y[f].x * y[f].x + y[f].y * y[f].y - s_nz[f * SRP_KC + lane]);
    }
  }
#pragma unroll
  for (int q = 0; q < DPW; ++q) {
    const int d = d0 + q;
#pragma unroll
    for (int f = 0; f < SRP_TF; ++f) {
      float v = warp_sum(acc[q][f]);
      if (lane == 0 && d < D && t0 + f < T) srp[((long long)b * T + t0 + f) * D + d] = v;
    }
  }
}

int k_srp_channel(const float2 *spec, int B, int T, int M, int N, const uint64_t *mic_fx, int D, float *srp, cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  constexpr int DPW = 2;
  size_t smem = sizeof(float2) * SRP_TF * M * SRP_KC + sizeof(float) * SRP_TF * SRP_KC;
  if (smem > 200 * 1024) return mcag_set_error(1, "srp: too many channels");
  auto kern = srp_channel_kernel<DPW>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int dper = 8 * DPW;
  dim3 grid((T + SRP_TF - 1) / SRP_TF, (D + dper - 1) / dper, B);
  kern<<<grid, 256, smem, st>>>(spec, T, M, N, mic_fx, D, srp);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

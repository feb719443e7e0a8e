// K4 delay-and-sum fan on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only: Beamformer::processFrame
// (Beamformer.cpp:51-71) steered to every azimuth of a fan (BASELINE config 3),
//     Y[t][d][k] = 1/M sum_c X_c[t][k] exp(j k phi_c(d)),
// per bin a complex [frames x M] . [M x D] contraction, run as a real GEMM with Re / Im stacked along K (2M) and N (2 x directions)
// and the 3xTF32 split (A_hi B_hi + A_lo B_hi + A_hi B_lo) like the SRP kernels of srp_tc.cu.
//
// The bin index is the BATCH index of the contraction, but the consumer (an inverse STFT per beam) needs [B][T][D][K] rows with the bin
// innermost.  A work item therefore keeps FOUR consecutive bins of a (128-frame, 64-direction) tile resident in TMEM (4 x 128 columns =
// all 512) and the epilogue writes, per (frame, direction), the four bins as 32 contiguous bytes — one full sector — instead of the
// isolated 8-byte stores of an accumulator that holds a single bin (round 1: 3.6 ms against 0.88 ms on CUDA cores).  No pre-pass, no
// workspace: 16 producer warps build the frames operand straight from the spectra (32-byte loads = the same four bins of one
// channel), 4 warps generate the steering operand from fixed-point phase increments, one thread issues the MMAs, 8 warps drain TMEM.
//   warp 0        MMA issuer: per (K chunk, bin) stage 12 tcgen05.mma.kind::tf32 M128 N128 K8 into the bin's 128 TMEM columns
//   warps 1-16    frames operand: thread = (frame row, 4 microphones of the 16-microphone K chunk)
//   warps 17-20   steering operand: thread = (direction, 8 microphones of the chunk), __sincosf of an exactly wrapped phase
//   warps 21-28   epilogue: TMEM lane = (direction, Re | Im), column = frame; neighbouring lanes swap parts, 16-byte stores that pair up
//                 into full sectors, 16 directions of one frame per store instruction
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace mcag {

constexpr int FT_BM = 128, FT_BD = 64, FT_NB = 4, FT_KC = 32, FT_STAGES = 3;
constexpr int FT_A_BYTES = FT_BM * FT_KC * 4;        // 16 KB
constexpr int FT_B_BYTES = 2 * FT_BD * FT_KC * 4;    // 16 KB
constexpr int FT_STAGE_BYTES = 2 * FT_A_BYTES + 2 * FT_B_BYTES;   // 64 KB
constexpr int FT_SMEM = FT_STAGES * FT_STAGE_BYTES + 1024 + 256;
constexpr int FT_A_THREADS = 512, FT_B_THREADS = 128, FT_EPI_THREADS = 256, FT_THREADS = 32 + FT_A_THREADS + FT_B_THREADS + FT_EPI_THREADS;
constexpr uint32_t FT_IDESC = umma_idesc_tf32(128, 128);

struct FtParams {
  const float2 *spec;     // [BT][M][KP]
  float2 *beams;          // [BT][D][KP]
  long long BT;
  int D, M, K, KP;
  int KPo, wide;          // pitch of the beam rows (complex bins); wide: rows and bin groups are 32-byte aligned -> one 256-bit store per row
  int n_tt, n_dt, n_kg;   // frame tiles, direction tiles, bin groups
  const uint64_t *steer_fx;   // [D][M] 0.64 fixed-point turns per bin
  float out_scale;        // 1 / M
};

// item -> (bin group kg, frame tile tt, direction tile dt).  The four bin groups of one 128-byte line of the beam rows are the fastest
// index and the direction tiles the next one: the ~12 CTAs that run neighbouring items at the same time complete whole lines in L2
// before they are evicted (with the bin group as the slowest index DRAM saw 24 M isolated 32-byte writes) and read the same spectra.
// Returns false for the padding items of the last line.
__device__ __forceinline__ bool ft_item(const FtParams &p, int item, int &kg, int &tt, int &dt) {
  const int kg_lo = item & 3;
  int rest = item >> 2;
  dt = rest % p.n_dt; rest /= p.n_dt;
  tt = rest % p.n_tt;
  kg = (rest / p.n_tt) * 4 + kg_lo;
  return kg < p.n_kg;
}

template <int NKC>   // K chunks per bin = M / 16
__global__ void __launch_bounds__(FT_THREADS, 1) ds_fan_tc_kernel(const FtParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + FT_STAGES * FT_STAGE_BYTES);
  uint64_t *full_a = bars, *full_b = bars + FT_STAGES, *empty = bars + 2 * FT_STAGES, *tmem_full = bars + 3 * FT_STAGES, *tmem_empty = tmem_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < FT_STAGES; ++s) { mbar_init(&full_a[s], FT_A_THREADS / 32); mbar_init(&full_b[s], FT_B_THREADS / 32); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_empty, FT_EPI_THREADS / 32);
  }
  if (warp == 0) {   // all 512 TMEM columns: four bins x (64 Re + 64 Im) columns; 1 CTA per SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.n_tt * p.n_dt * ((p.n_kg + 3) / 4) * 4;

  if (warp == 0) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int kg, tt, dt;
        if (!ft_item(p, item, kg, tt, dt)) continue;
        mbar_wait_bounded(tmem_empty, acc_phase ^ 1);   // the epilogue has drained the previous item
        tc_fence_after();
        for (int cc = 0; cc < NKC; ++cc)
          for (int b = 0; b < FT_NB; ++b) {
            mbar_wait_bounded(&full_a[stage], phase);
            mbar_wait_bounded(&full_b[stage], phase);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)b * 128u;
            const uint32_t sa = smem_u32(smem + stage * FT_STAGE_BYTES);
            const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + FT_A_BYTES);
            const uint64_t b_hi = umma_desc_sw128(sa + 2 * FT_A_BYTES), b_lo = umma_desc_sw128(sa + 2 * FT_A_BYTES + FT_B_BYTES);
#pragma unroll
            for (int j = 0; j < FT_KC / 8; ++j) {   // K = 8 per instruction: 32 bytes along the swizzled row -> +2 in the address field
              // steering rows on the M side (TMEM lanes), frames on the N side (TMEM columns): D[2 d + {Re, Im}][t]
              umma_tf32(d_tmem, b_hi + 2 * j, a_hi + 2 * j, FT_IDESC, (cc | j) != 0);
              umma_tf32(d_tmem, b_hi + 2 * j, a_lo + 2 * j, FT_IDESC, 1);
              umma_tf32(d_tmem, b_lo + 2 * j, a_hi + 2 * j, FT_IDESC, 1);
            }
            umma_commit(&empty[stage]);
            if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
          }
        umma_commit(tmem_full);
        acc_phase ^= 1;
      }
    }
  } else if (warp <= FT_A_THREADS / 32) {
    // ===== frames operand: raw spectra -> hi / lo K-major tiles, 128-byte swizzle =====
    const int g = tid - 32;
    // microphones 4 qm .. 4 qm + 3 of the chunk, frame row of the tile.  qm runs fastest along the lanes: a warp load then covers 8
    // frames x 4 microphone rows (two pages) instead of 32 frames 256 KB apart (one page per lane)
    const int qm = g & 3, row = g >> 2;
    const uint32_t a_row = (uint32_t)row * 128u, a_sw = (uint32_t)(row & 7);
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int kg, tt, dt;
      if (!ft_item(p, item, kg, tt, dt)) continue;
      const int k0 = kg * FT_NB;
      const long long t = (long long)tt * FT_BM + row;
      const bool valid = t < p.BT;
      const bool second = k0 + 2 < p.KP;   // the last group of a row holds two bins only (KP = N/2 + 2)
      for (int cc = 0; cc < NKC; ++cc) {
        const float4 *src = reinterpret_cast<const float4 *>(p.spec + ((valid ? t : 0) * p.M + cc * 16 + qm * 4) * p.KP + k0);
        float4 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u][0] = __ldg(src + (size_t)u * (p.KP >> 1));
          v[u][1] = second ? __ldg(src + (size_t)u * (p.KP >> 1) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int b = 0; b < FT_NB; ++b) {
          float x[8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 w = v[u][b >> 1];
            x[2 * u] = valid ? ((b & 1) ? w.z : w.x) : 0.f;
            x[2 * u + 1] = valid ? ((b & 1) ? w.w : w.y) : 0.f;
          }
          mbar_wait_bounded(&empty[stage], phase ^ 1);
          unsigned char *st = smem + stage * FT_STAGE_BYTES;
#pragma unroll
          for (int q = 0; q < 2; ++q) {   // two microphones = one 16-byte chunk of the frame's row
            const float h0 = tf32_hi(x[4 * q]), h1 = tf32_hi(x[4 * q + 1]), h2 = tf32_hi(x[4 * q + 2]), h3 = tf32_hi(x[4 * q + 3]);
            const uint32_t off = ((uint32_t)(qm * 2 + q) ^ a_sw) << 4;
            *reinterpret_cast<float4 *>(st + a_row + off) = make_float4(h0, h1, h2, h3);
            *reinterpret_cast<float4 *>(st + FT_A_BYTES + a_row + off) = make_float4(x[4 * q] - h0, x[4 * q + 1] - h1, x[4 * q + 2] - h2, x[4 * q + 3] - h3);
          }
          fence_async_smem();
          mbar_arrive_warp(&full_a[stage]);
          if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp <= (FT_A_THREADS + FT_B_THREADS) / 32) {
    // ===== steering operand, generated: S[2d][2m] = cos, S[2d][2m+1] = -sin (Re rows); S[2d+1][2m] = sin, S[2d+1][2m+1] = cos (Im rows) =====
    const int g = tid - 32 - FT_A_THREADS;
    const int dl = g & (FT_BD - 1), grp = g >> 6;     // direction of the tile; microphones 8 grp .. 8 grp + 7 of the chunk
    // rows 2 dl (Re) and 2 dl + 1 (Im): the two parts of a direction land in neighbouring TMEM lanes of the same epilogue warp
    const uint32_t b_re = (uint32_t)(2 * dl) * 128u, b_im = (uint32_t)(2 * dl + 1) * 128u;
    const uint32_t sw_re = (uint32_t)((2 * dl) & 7), sw_im = (uint32_t)((2 * dl + 1) & 7);
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int kg, tt, dt;
      if (!ft_item(p, item, kg, tt, dt)) continue;
      const int k0 = kg * FT_NB;
      const int d = min(dt * FT_BD + dl, p.D - 1);
      uint32_t fx[NKC * 8];   // 32-bit fixed-point turns per bin; k * fx wraps exactly mod one turn
#pragma unroll
      for (int cc = 0; cc < NKC; ++cc)
#pragma unroll
        for (int u = 0; u < 8; ++u) fx[cc * 8 + u] = (uint32_t)((p.steer_fx[(size_t)d * p.M + cc * 16 + grp * 8 + u] + 0x80000000ull) >> 32);
#pragma unroll
      for (int cc = 0; cc < NKC; ++cc)
        for (int b = 0; b < FT_NB; ++b) {
          const uint32_t k = (uint32_t)(k0 + b);
          float c[8], s[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int32_t ph = (int32_t)(fx[cc * 8 + u] * k);                           // signed turns * 2^32
            __sincosf((float)ph * 1.4629180792671596e-09f, &s[u], &c[u]);               // 2 pi / 2^32
          }
          mbar_wait_bounded(&empty[stage], phase ^ 1);
          unsigned char *bh = smem + stage * FT_STAGE_BYTES + 2 * FT_A_BYTES, *bl = bh + FT_B_BYTES;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float ch0 = tf32_hi(c[2 * u]), sh0 = tf32_hi(s[2 * u]), ch1 = tf32_hi(c[2 * u + 1]), sh1 = tf32_hi(s[2 * u + 1]);
            const float cl0 = c[2 * u] - ch0, sl0 = s[2 * u] - sh0, cl1 = c[2 * u + 1] - ch1, sl1 = s[2 * u + 1] - sh1;
            const uint32_t q_re = ((uint32_t)(grp * 4 + u) ^ sw_re) << 4, q_im = ((uint32_t)(grp * 4 + u) ^ sw_im) << 4;
            *reinterpret_cast<float4 *>(bh + b_re + q_re) = make_float4(ch0, -sh0, ch1, -sh1);
            *reinterpret_cast<float4 *>(bh + b_im + q_im) = make_float4(sh0, ch0, sh1, ch1);
            *reinterpret_cast<float4 *>(bl + b_re + q_re) = make_float4(cl0, -sl0, cl1, -sl1);
            *reinterpret_cast<float4 *>(bl + b_im + q_im) = make_float4(sl0, cl0, sl1, cl1);
          }
          fence_async_smem();
          mbar_arrive_warp(&full_b[stage]);
          if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
        }
    }
  } else {
    // ===== epilogue: 4 bins of a (frame, direction) leave as 32 contiguous bytes =====
    // TMEM lane = steering row (even lanes Re, odd lanes Im of one direction), column = frame.  A store instruction of a warp covers
    // one frame x 16 directions x one full 32-byte sector: even lanes write bins k0, k0+1, odd lanes bins k0+2, k0+3 after swapping the
    // parts they lack with their neighbour.  What this pattern can reach is measured in tools/ubench/scatter_store2.cu: 16 rows x 32 B
    // per store instruction run at 1.8 TB/s (16 B per lane-row: 1.2 TB/s, 128 B per row: 3.8 TB/s, 512 B: 4.6 TB/s), whoever issues them
    // - a TMA store of the same 32-byte rows was slower still (0.83 ms for the kernel against 0.59 ms).  The stores are therefore 0.3 ms
    // of this kernel.  Sixteen bins per row resident in TMEM (N = 32 frames per MMA, 128-byte rows behind a shared-memory transpose) was
    // built three ways in round 2 and stayed behind this kernel, see DESIGN.md 4 "what did not work": with a quarter of the frames per
    // MMA the operand stages quadruple, and the shared-memory write bandwidth of the generated steering operand (32 KB per stage), the
    // per-instruction issue cost of tcgen05.mma from one elected lane and the 32-byte gathers of the frames operand each became the bound.
    const int e = warp - (1 + (FT_A_THREADS + FT_B_THREADS) / 32);
    const int quarter = warp & 3, fh = e >> 2;   // TMEM lanes 32 quarter..+31 = directions 16 quarter..+15; frames 64 fh..+63 of the tile
    const bool odd = lane & 1;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int kg, tt, dt;
      if (!ft_item(p, item, kg, tt, dt)) continue;
      const int k0 = kg * FT_NB;
      const bool second = k0 + 2 < p.KP;
      const int d = dt * FT_BD + quarter * 16 + (lane >> 1);
      const bool store = d < p.D && (!odd || second);
      float2 *col = p.beams + (size_t)d * p.KPo + k0 + ((odd && !p.wide) ? 2 : 0);   // + t * D * KPo per frame
      mbar_wait_bounded(tmem_full, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(fh * 64);
#pragma unroll 2
      for (int j = 0; j < 8; ++j) {   // 8 frames at a time: 4 bins x 8 columns = 32 registers
        uint32_t r[FT_NB][8];
#pragma unroll
        for (int b = 0; b < FT_NB; ++b) tmem_ld8_nowait(taddr + b * 128 + j * 8, r[b]);
        tmem_ld_wait();
        if (p.wide) {
          // 32-byte aligned rows: ONE 256-bit store per (frame, direction) row - st.global.v8.f32 reaches 4.3 TB/s on 32-byte row pieces where
          // two 16-byte stores of a lane pair reach 1.8 (tools/ubench/scatter_store3.cu).  The Re lane writes the even frame of a pair, the
          // Im lane the odd one; each hands the other its four bins of the frame it does not write.
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float mine[4], give[4], got[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const float ve = __uint_as_float(r[b][i]) * p.out_scale, vo = __uint_as_float(r[b][i + 1]) * p.out_scale;
              mine[b] = odd ? vo : ve;
              give[b] = odd ? ve : vo;
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) got[b] = __shfl_xor_sync(0xffffffffu, give[b], 1);
            const long long t = (long long)tt * FT_BM + fh * 64 + j * 8 + i + (odd ? 1 : 0);
            if (d < p.D && t < p.BT) {
              float2 *dst = col + (size_t)t * p.D * p.KPo;
              if (odd) st_global_v8(dst, got[0], mine[0], got[1], mine[1], got[2], mine[2], got[3], mine[3]);
              else st_global_v8(dst, mine[0], got[0], mine[1], got[1], mine[2], got[2], mine[3], got[3]);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long t = (long long)tt * FT_BM + fh * 64 + j * 8 + i;
            const float v0 = __uint_as_float(r[0][i]) * p.out_scale, v1 = __uint_as_float(r[1][i]) * p.out_scale;
            const float v2 = __uint_as_float(r[2][i]) * p.out_scale, v3 = __uint_as_float(r[3][i]) * p.out_scale;
            // even lane (Re) lacks Im of bins 0, 1 and gives Re of bins 2, 3; odd lane (Im) the other way round
            const float g0 = __shfl_xor_sync(0xffffffffu, odd ? v0 : v2, 1), g1 = __shfl_xor_sync(0xffffffffu, odd ? v1 : v3, 1);
            const float4 o = odd ? make_float4(g0, v2, g1, v3) : make_float4(v0, g0, v1, g1);
            if (store && t < p.BT) *reinterpret_cast<float4 *>(col + (size_t)t * p.D * p.KPo) = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive_warp(tmem_empty);
      acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

bool k_ds_fan_tensor_supported(int M) { return M % 16 == 0 && M >= 16 && M <= 64; }

// Supported shapes: M in {16, 32, 48, 64}; anything else runs the CUDA-core tile kernel (still on the GPU).  The spectra's pad bin
// (index N/2 + 1) must be zero, as every producer of spectra in this library leaves it: it comes out as the beams' pad bin.
int k_ds_fan_tensor(const float2 *spec, int B, int T, int M, int N, const uint64_t *steer_fx, int D, float2 *out, cudaStream_t st, int out_pitch) {
  if (B <= 0 || T <= 0 || D <= 0) return 0;
  if (!k_ds_fan_tensor_supported(M)) return k_ds_fan(spec, B, T, M, N, steer_fx, D, out, st, out_pitch);
  FtParams p;
  p.spec = spec; p.beams = out; p.BT = (long long)B * T; p.D = D; p.M = M; p.K = N / 2 + 1; p.KP = spec_pitch(N);
  p.KPo = out_pitch > 0 ? out_pitch : p.KP;
  if (p.KPo < p.KP) return mcag_set_error(1, "ds_fan_tensor: output pitch below the spectrum pitch");
  // rows that start on 32-byte boundaries and hold whole four-bin groups: the 256-bit store path (fan_out_pitch(N) rows of MCAG_KIND_DSFAN)
  p.wide = ((p.KPo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) ? 1 : 0;
  if (!p.wide && p.KPo != p.KP) return mcag_set_error(1, "ds_fan_tensor: a padded output pitch must be a multiple of 4 bins on a 32-byte aligned buffer");
  p.n_tt = (int)((p.BT + FT_BM - 1) / FT_BM); p.n_dt = (D + FT_BD - 1) / FT_BD; p.n_kg = (p.KP + FT_NB - 1) / FT_NB;
  p.steer_fx = steer_fx; p.out_scale = 1.0f / (float)M;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)p.n_tt * p.n_dt * ((p.n_kg + 3) / 4) * 4;
  const unsigned grid = (unsigned)(items < sms ? items : sms);
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    kern<<<grid, FT_THREADS, FT_SMEM, st>>>(p);
  };
  switch (M / 16) {
    case 1: launch(ds_fan_tc_kernel<1>); break;
    case 2: launch(ds_fan_tc_kernel<2>); break;
    case 3: launch(ds_fan_tc_kernel<3>); break;
    default: launch(ds_fan_tc_kernel<4>); break;
  }
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

extern "C" int mcag_k_ds_fan_tensor(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_steer_fx, int D, void *d_out, void *stream) {
  return mcag::k_ds_fan_tensor((const float2 *)d_spec, B, T, M, N, d_steer_fx, D, (float2 *)d_out, (cudaStream_t)stream, 0);
}

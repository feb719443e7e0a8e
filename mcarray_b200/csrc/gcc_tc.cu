// K2 gcc_phat on a tau grid, on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// dsp::GeneralisedCrossCorrelation::calculateCorrelationsForPrecomputedTauMatrix as the reference drives it
// (SteeringBeamforming.cpp:104-130 per pair, BinauralLocalisation.cpp:438-444 for the two-microphone localiser):
//     corr[t][p][d] = Re sum_{k<K} G_p[t][k] exp(+j 2 pi k tau_pd / N),      G_p = PHAT(X_i conj X_j).
// For one pair this is a plain real GEMM with the BINS as the contraction dimension (Re / Im interleaved: 2K = N + 2 deep),
//     D[t][d] = sum_c A[t][c] B[d][c],     A[t][2k] = Re G, A[t][2k+1] = Im G,     B[d][2k] = cos, B[d][2k+1] = -sin,
// frames on the UMMA M side (128 rows), delays on the N side (64 rows), 3xTF32 (A_hi B_hi + A_lo B_hi + A_hi B_lo).  This is the
// steered-response contraction of the north_star for the pair form (the channel form of srp_tc.cu contracts over microphones per bin
// instead); the CUDA-core register-tile kernel gcc_tau_kernel spends 0.92 ms on cfg1l's 256 k frames x 61 delays, FP32-issue bound.
//
// Per CTA (persistent over work items = pair x 128-frame tile):
//   warp 0       MMA issuer: per 16-bin chunk 12 tcgen05.mma.kind::tf32 M128 N64 K8 into one of two 64-column TMEM accumulators
//   warp 1       table loader: the delay operand is the same for every frame tile, so it is tabulated once per processor in the
//                canonical swizzled K-major layout (gcc_tc_table_kernel) and fetched per chunk by one 16 KB bulk copy (TMA engine)
//   warps 2-17   cross-spectrum producers: thread = (two frame rows, two bins): 16-byte loads of both channels with the BINS along the
//                lanes (eight lanes cover 128 contiguous bytes of a spectrum row), PHAT, 3xTF32 split, swizzled 16-byte stores
//   warps 18-21  epilogue: tcgen05.ld the finished accumulator, store corr rows
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace mcag {

constexpr int GT2_BM = 128, GT2_BD = 64, GT2_KC = 32, GT2_STAGES = 4;
constexpr int GT2_A_BYTES = GT2_BM * GT2_KC * 4;   // 16 KB
constexpr int GT2_B_BYTES = GT2_BD * GT2_KC * 4;   // 8 KB
constexpr int GT2_STAGE_BYTES = 2 * GT2_A_BYTES + 2 * GT2_B_BYTES;   // 48 KB
constexpr int GT2_SMEM = GT2_STAGES * GT2_STAGE_BYTES + 1024 + 256;
// 16 producer warps: the operand build is a latency chain (load, whiten, store, fence, arrive) per chunk and warp; 8 warps ran cfg1l in
// 0.41 ms, 16 in 0.35 ms, 4 in 0.62 ms
constexpr int GT2_PROD_THREADS = 512, GT2_RPT = 128 * 8 / GT2_PROD_THREADS /* frame rows per producer thread */, GT2_EPI_THREADS = 128, GT2_THREADS = 64 + GT2_PROD_THREADS + GT2_EPI_THREADS;
constexpr uint32_t GT2_IDESC = umma_idesc_tf32(128, 64);

__host__ __device__ constexpr int gt2_chunks(int N) { return (N + 2 + GT2_KC - 1) / GT2_KC; }   // 2K = N + 2 floats deep

// table[p][chunk][hi | lo][64 rows][32 floats], rows in the 128-byte-swizzled K-major layout the UMMA descriptors expect
__global__ void gcc_tc_table_kernel(const uint64_t *__restrict__ pair_fx, int P, int D, int N, float *__restrict__ table) {
  const int NCH = gt2_chunks(N), K = N / 2 + 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (p, chunk, row d, 16-byte unit q): two bins
  if (i >= (long long)P * NCH * GT2_BD * 8) return;
  const int q = (int)(i & 7), d = (int)((i >> 3) & (GT2_BD - 1));
  const int c = (int)((i >> 9) % NCH), p = (int)((i >> 9) / NCH);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (d < D)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = c * 16 + 2 * q + u;
      if (k < K) { const float2 w = phase_ramp(pair_fx[(size_t)p * D + d], k); v[2 * u] = w.x; v[2 * u + 1] = -w.y; }
    }
  float h[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) h[u] = tf32_hi(v[u]);
  float *blk = table + ((size_t)p * NCH + c) * (2 * GT2_B_BYTES / 4);
  const int off = d * 32 + ((q ^ (d & 7)) << 2);
  *reinterpret_cast<float4 *>(blk + off) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4 *>(blk + GT2_B_BYTES / 4 + off) = make_float4(v[0] - h[0], v[1] - h[1], v[2] - h[2], v[3] - h[3]);
}

struct Gt2Params {
  const float2 *spec;   // [BT][M][KP]
  long long BT;
  int M, P, D, K, KP, NCH, n_tt;
  const float *table;   // gcc_tc_table_kernel
  float *corr;          // [BT][P][D]
};

__global__ void __launch_bounds__(GT2_THREADS, 1) gcc_tau_tc_kernel(const Gt2Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + GT2_STAGES * GT2_STAGE_BYTES);
  uint64_t *full_a = bars, *full_b = bars + GT2_STAGES, *empty = bars + 2 * GT2_STAGES, *tmem_full = bars + 3 * GT2_STAGES, *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < GT2_STAGES; ++s) { mbar_init(&full_a[s], GT2_PROD_THREADS / 32); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], GT2_EPI_THREADS / 32); }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.P * p.n_tt;   // item -> (pair, frame tile), pair fastest: neighbouring CTAs read the same spectra

  if (warp == 0) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait_bounded(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 64u;
        for (int c = 0; c < p.NCH; ++c) {
          mbar_wait_bounded(&full_a[stage], phase);
          mbar_wait_bounded(&full_b[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * GT2_STAGE_BYTES);
          const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + GT2_A_BYTES);
          const uint64_t b_hi = umma_desc_sw128(sa + 2 * GT2_A_BYTES), b_lo = umma_desc_sw128(sa + 2 * GT2_A_BYTES + GT2_B_BYTES);
#pragma unroll
          for (int j = 0; j < GT2_KC / 8; ++j) {
            umma_tf32(d_tmem, a_hi + 2 * j, b_hi + 2 * j, GT2_IDESC, (c | j) != 0);
            umma_tf32(d_tmem, a_lo + 2 * j, b_hi + 2 * j, GT2_IDESC, 1);
            umma_tf32(d_tmem, a_hi + 2 * j, b_lo + 2 * j, GT2_IDESC, 1);
          }
          umma_commit(&empty[stage]);
          if (++stage == GT2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== delay-operand loader: one bulk copy (hi + lo tile) per chunk =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int pr = item % p.P;
        for (int c = 0; c < p.NCH; ++c) {
          mbar_wait_bounded(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full_b[stage], 2 * GT2_B_BYTES);
          bulk_g2s(smem + stage * GT2_STAGE_BYTES + 2 * GT2_A_BYTES, p.table + ((size_t)pr * p.NCH + c) * (2 * GT2_B_BYTES / 4), 2 * GT2_B_BYTES,
                   &full_b[stage]);
          if (++stage == GT2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 2 + GT2_PROD_THREADS / 32) {
    // ===== cross-spectrum producers =====
    const int g = tid - 64;
    const int q = g & 7, r0 = g >> 3;   // 16-byte unit (two bins) of the chunk; frame rows r0, r0 + 32, r0 + 64, r0 + 96 of the tile
    int stage = 0; uint32_t phase = 0;
    auto decode = [&](int item, int &tt, int &mi, int &mj) {   // pair -> (i, j), i < j lexicographic (SteeringBeamforming.cpp:63-65)
      int rem = item % p.P;
      tt = item / p.P; mi = 0;
      while (rem >= p.M - 1 - mi) { rem -= p.M - 1 - mi; ++mi; }
      mj = mi + 1 + rem;
    };
    // always a load (rows past the last frame and units past the row end read frame 0 / unit 0 and are zeroed when they are USED): a select
    // against zero right behind the load would make the prefetch wait for its own data
    auto load_chunk = [&](int tt, int mi, int mj, int c, float4 (&l)[GT2_RPT], float4 (&r)[GT2_RPT]) {
      const int k0 = c * 16 + 2 * q;
      const bool inrow = k0 < p.KP;   // KP is even: a 16-byte unit is inside the row or entirely past it
#pragma unroll
      for (int i = 0; i < GT2_RPT; ++i) {
        const long long t = (long long)tt * GT2_BM + r0 + (GT2_PROD_THREADS / 8) * i;
        const bool ok = inrow && t < p.BT;
        const float2 *row = p.spec + (ok ? t : 0) * p.M * p.KP + (ok ? k0 : 0);
        l[i] = __ldg(reinterpret_cast<const float4 *>(row + (size_t)mi * p.KP));
        r[i] = __ldg(reinterpret_cast<const float4 *>(row + (size_t)mj * p.KP));
      }
    };
    // The spectra of the NEXT chunk (or of the next item's first chunk) are requested before the current chunk is whitened and stored
    int item = blockIdx.x;
    if (item < n_items) {
      int tt, mi, mj;
      decode(item, tt, mi, mj);
      float4 l[GT2_RPT], r[GT2_RPT], nl[GT2_RPT], nr[GT2_RPT];
      load_chunk(tt, mi, mj, 0, l, r);
      while (item < n_items) {
        const int nitem = item + gridDim.x;
        int ntt = 0, nmi = 0, nmj = 1;
        if (nitem < n_items) decode(nitem, ntt, nmi, nmj);
        for (int c = 0; c < p.NCH; ++c) {
          if (c + 1 < p.NCH) load_chunk(tt, mi, mj, c + 1, nl, nr);
          else if (nitem < n_items) load_chunk(ntt, nmi, nmj, 0, nl, nr);
          const int k0 = c * 16 + 2 * q;
          const bool inrow = k0 < p.KP;
          mbar_wait_bounded(&empty[stage], phase ^ 1);
          unsigned char *st = smem + stage * GT2_STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < GT2_RPT; ++i) {
            const int row = r0 + (GT2_PROD_THREADS / 8) * i;
            const bool ok = inrow && (long long)tt * GT2_BM + row < p.BT;
            float2 g0 = whiten(cmulc(make_float2(l[i].x, l[i].y), make_float2(r[i].x, r[i].y)));
            float2 g1 = whiten(cmulc(make_float2(l[i].z, l[i].w), make_float2(r[i].z, r[i].w)));
            if (!ok || k0 >= p.K) g0 = make_float2(0.f, 0.f);        // the pad bin never contributes
            if (!ok || k0 + 1 >= p.K) g1 = make_float2(0.f, 0.f);
            const float h0 = tf32_hi(g0.x), h1 = tf32_hi(g0.y), h2 = tf32_hi(g1.x), h3 = tf32_hi(g1.y);
            const uint32_t off = (uint32_t)row * 128u + (((uint32_t)q ^ (uint32_t)(row & 7)) << 4);
            *reinterpret_cast<float4 *>(st + off) = make_float4(h0, h1, h2, h3);
            *reinterpret_cast<float4 *>(st + GT2_A_BYTES + off) = make_float4(g0.x - h0, g0.y - h1, g1.x - h2, g1.y - h3);
          }
          fence_async_smem();
          mbar_arrive_warp(&full_a[stage]);
          if (++stage == GT2_STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
          for (int i = 0; i < GT2_RPT; ++i) { l[i] = nl[i]; r[i] = nr[i]; }
        }
        item = nitem; tt = ntt; mi = nmi; mj = nmj;
      }
    }
  } else {
    // ===== epilogue =====
    const int quarter = warp & 3;   // TMEM lanes 32 quarter..+31 (four consecutive warps cover the four quarters)
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int pr = item % p.P, tt = item / p.P;
      const long long t = (long long)tt * GT2_BM + quarter * 32 + lane;
      mbar_wait_bounded(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 64);
      float *dst = p.corr + ((t < p.BT ? t : 0) * p.P + pr) * p.D;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[16];
        tmem_ld16(taddr + j * 16, v);
        if (t < p.BT) {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (j * 16 + i < p.D) dst[j * 16 + i] = v[i];
        }
      }
      tc_fence_before();
      mbar_arrive_warp(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
}

bool k_gcc_tau_tc_supported(int D) { return D >= 1 && D <= GT2_BD; }
size_t k_gcc_tau_tc_table_bytes(int P, int N) { return (size_t)P * gt2_chunks(N) * 2 * GT2_B_BYTES; }

int k_gcc_tau_tc_build(const uint64_t *pair_fx, int P, int D, int N, float *table, cudaStream_t st) {
  if (!k_gcc_tau_tc_supported(D)) return mcag_set_error(1, "gcc_tau_tc: at most 64 delays");
  const long long n = (long long)P * gt2_chunks(N) * GT2_BD * 8;
  gcc_tc_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pair_fx, P, D, N, table);
  MCAG_CHECK_LAUNCH();
  return 0;
}

int k_gcc_tau_tc(const float2 *spec, int B, int T, int M, int N, const float *table, int D, float *corr, cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  Gt2Params p;
  p.spec = spec; p.BT = (long long)B * T; p.M = M; p.P = M * (M - 1) / 2; p.D = D; p.K = N / 2 + 1; p.KP = spec_pitch(N); p.NCH = gt2_chunks(N);
  p.n_tt = (int)((p.BT + GT2_BM - 1) / GT2_BM); p.table = table; p.corr = corr;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)p.P * p.n_tt;
  cudaFuncSetAttribute(gcc_tau_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT2_SMEM);
  gcc_tau_tc_kernel<<<(unsigned)(items < sms ? items : sms), GT2_THREADS, GT2_SMEM, st>>>(p);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

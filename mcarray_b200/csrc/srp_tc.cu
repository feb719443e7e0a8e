// K3 srp_contract on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Channel form of the SRP-PHAT pair sum of SteeringBeamforming::computeCorrelations (SteeringBeamforming.cpp:104-130,
// SURVEY.md §8a row A4):   srp[t][d] = 1/2 sum_k ( |Y_k[t][d]|^2 - nz_k[t] ),   Y_k[t][d] = sum_m U_m[t][k] a_m[d][k],
// U = X/|X| (PHAT whitening), a_m[d][k] = exp(-j 2 pi k tau_m(d) / N).  Per bin k this is a complex [T x M] . [M x D]
// contraction, run as a real GEMM with the real / imaginary parts stacked along K (2M) and along N (2 x directions):
//
//     D[t][n] += A[t][c] * B[n][c]        A = frames (UMMA M = 128 rows, K-major)      c = 2m + {re, im}
//                                          B = steering (UMMA N = 256 rows, K-major)    n = d (Re rows 0..127), 128 + d (Im rows)
//     B[d][2m] = Re a, B[d][2m+1] = -Im a;   B[128+d][2m] = Im a, B[128+d][2m+1] = Re a
//
// fp32 accuracy comes from the 3xTF32 split: x = hi + lo with hi = the TF32 the tensor core would read (low 13 mantissa bits
// cleared) and lo = x - hi (exact); D += A_hi B_hi + A_lo B_hi + A_hi B_lo (the dropped lo*lo term is ~2^-22 relative).
//
// Data flow per CTA (persistent over work items = frame tile x direction tile x bin range, 1 CTA per SM):
//   warp 0      TMA producer: the frames operand comes from a bin-major, pre-whitened, pre-split copy of the spectra
//               (srp_prepare_kernel) as 128B-swizzled [128 x 32] fp32 boxes, hi and lo, per 32-float K chunk
//   warp 1      MMA issuer: one thread issues 12 tcgen05.mma.kind::tf32 (M128 N256 K8) per K chunk into one of two
//               256-column TMEM accumulators, commits the smem stage back to the producers and the accumulator to the epilogue
//   warps 4-11  steering generators: the B operand is never loaded - each thread keeps the 32-bit fixed-point phase
//               increments of its (direction, microphone) elements in registers, evaluates sin/cos of k*increment for the
//               current bin and writes the hi / lo tiles straight into shared memory in the canonical swizzled layout
//   warps 12-19 epilogue: tcgen05.ld the finished accumulator, square, accumulate over the bins of the work item in registers
//               (|Y|^2 is not linear, so it cannot stay in TMEM), and store the partial energy map at the end of the item
// A second small kernel adds the partial maps in a fixed order (deterministic) and applies the -nz/2 term.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

#define OK_RC(call)        \
  do {                     \
    int rc__ = (call);     \
    if (rc__) return rc__; \
  } while (0)

namespace mcag {

constexpr int TC_BM = 128;        // frames per tile (UMMA M)
constexpr int TC_BD = 128;        // directions per tile (UMMA N = 2 * TC_BD)
constexpr int TC_KC = 32;         // floats per K chunk = one 128-byte swizzle row
constexpr int TC_STAGES = 2;
constexpr int TC_GEN_THREADS = 256, TC_EPI_THREADS = 256;
constexpr int TC_THREADS = 128 + TC_GEN_THREADS + TC_EPI_THREADS;
constexpr int TC_A_BYTES = TC_BM * TC_KC * 4;          // 16 KB
constexpr int TC_B_BYTES = 2 * TC_BD * TC_KC * 4;      // 32 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 96 KB
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

// ---- pre-pass: whiten, split, transpose to bin-major --------------------------------------------------------------------
// spec [BT][M][KP] float2  ->  U_hi / U_lo [K][BTpad][2M] fp32 (rows t >= BT stay zero), nzsum[t] = sum_k #(non-zero channels)
__global__ void __launch_bounds__(256) srp_prepare_kernel(const float2 *__restrict__ spec, long long BT, long long BTpad, int M, int N,
                                                           float *__restrict__ Uhi, float *__restrict__ Ulo, float *__restrict__ nzsum) {
  __shared__ float2 s_t[32][65];   // [bin in chunk][mic], M <= 64
  __shared__ float s_nz[8];
  const int KP = spec_pitch(N), K = N / 2 + 1;
  const long long t = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (t >= BT) {   // rows of the last frame tile beyond BT must read as zeros
    for (int i = tid; i < K * M; i += 256) {
      const int k = i / M, m = i - k * M;
      const long long o = ((long long)k * BTpad + t) * (2 * M) + 2 * m;
      *reinterpret_cast<float2 *>(Uhi + o) = make_float2(0.f, 0.f);
      *reinterpret_cast<float2 *>(Ulo + o) = make_float2(0.f, 0.f);
    }
    return;
  }
  float nz = 0.f;
  // the spectra of the NEXT 32-bin chunk are requested before the current one is transposed and written: one global round trip per chunk
  // was exposed before (0.26 ms for cfg4's 403 MB = 1.5 TB/s)
  float2 cur[8], nxt[8];   // microphones warp, warp + 8, ... (M <= 64)
  auto load_chunk = [&](int k0, float2 (&v)[8]) {
    const int k = k0 + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = warp + 8 * i;
      v[i] = (m < M && k < K) ? __ldg(spec + (t * M + m) * KP + k) : make_float2(0.f, 0.f);
    }
  };
  load_chunk(0, cur);
  for (int k0 = 0; k0 < K; k0 += 32) {
    if (k0 + 32 < K) load_chunk(k0 + 32, nxt);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = warp + 8 * i;
      if (m < M) {
        const float2 u = whiten(cur[i]);   // whiten(0) = 0: bins past K stay zero
        nz += (u.x != 0.f || u.y != 0.f) ? 1.f : 0.f;
        s_t[lane][m] = u;
      }
    }
    __syncthreads();
    for (int i = tid; i < 32 * M; i += 256) {   // coalesced along m (8 bytes per mic)
      const int kk = i / M, m = i - kk * M, k = k0 + kk;
      if (k < K) {
        const float2 u = s_t[kk][m];
        const float2 h = make_float2(tf32_hi(u.x), tf32_hi(u.y));
        const long long o = ((long long)k * BTpad + t) * (2 * M) + 2 * m;
        *reinterpret_cast<float2 *>(Uhi + o) = h;
        *reinterpret_cast<float2 *>(Ulo + o) = make_float2(u.x - h.x, u.y - h.y);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
  }
  nz = warp_sum(nz);
  if (lane == 0) s_nz[warp] = nz;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_nz[w];
    nzsum[t] = s;
  }
}

// ---- main kernel -----------------------------------------------------------------------------------------------------------
struct TcParams {
  long long BT;          // frames (all streams)
  int D, M, K;           // directions, microphones, one-sided bins
  int n_tt, n_dt, n_ks;  // frame tiles, direction tiles, bin ranges
  int bins_per_range;
  const uint64_t *mic_fx;   // [D][M] 0.64 fixed-point turns per bin
  float *partial;           // [n_ks][BT][D] partial energy maps
};

template <int NKC>   // K chunks per bin = 2M / 32 = M / 16
__global__ void __launch_bounds__(TC_THREADS, 1) srp_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                                                                const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array itself: an integer round-trip of the pointer loses the address
  // space and the operand-tile stores become generic ST.E.128 (long-scoreboard WAR stalls in the producers, ncu s4_cfg5_small)
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t *full_a = bars, *full_b = bars + TC_STAGES, *empty = bars + 2 * TC_STAGES, *tmem_full = bars + 3 * TC_STAGES,
           *tmem_empty = bars + 3 * TC_STAGES + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * TC_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_a[s], 1); mbar_init(&full_b[s], TC_GEN_THREADS / 32); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], TC_EPI_THREADS / 32); }
  }
  if (warp == 1) {   // TMEM: all 512 columns (two 256-column accumulators); 1 CTA per SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = p.n_tt * p.n_dt * p.n_ks;

  if (warp == 0) {
    // ===== TMA producer (frames operand) =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ks = item / (p.n_tt * p.n_dt), tt = (item / p.n_dt) % p.n_tt;
        const int k_begin = ks * p.bins_per_range, k_end = min(p.K, k_begin + p.bins_per_range);
        for (int k = k_begin; k < k_end; ++k)
          for (int cc = 0; cc < NKC; ++cc) {
            mbar_wait_bounded(&empty[stage], phase ^ 1);
            unsigned char *st = smem + stage * TC_STAGE_BYTES;
            mbar_expect_tx(&full_a[stage], 2 * TC_A_BYTES);
            tma_load_3d(st, &map_hi, &full_a[stage], cc * TC_KC, tt * TC_BM, k);
            tma_load_3d(st + TC_A_BYTES, &map_lo, &full_a[stage], cc * TC_KC, tt * TC_BM, k);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ks = item / (p.n_tt * p.n_dt);
        const int k_begin = ks * p.bins_per_range, k_end = min(p.K, k_begin + p.bins_per_range);
        for (int k = k_begin; k < k_end; ++k) {
          mbar_wait_bounded(&tmem_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
          for (int cc = 0; cc < NKC; ++cc) {
            mbar_wait_bounded(&full_a[stage], phase);
            mbar_wait_bounded(&full_b[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * TC_STAGE_BYTES);
            const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + TC_A_BYTES);
            const uint64_t b_hi = umma_desc_sw128(sa + 2 * TC_A_BYTES), b_lo = umma_desc_sw128(sa + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
            for (int j = 0; j < TC_KC / 8; ++j) {   // K = 8 per instruction: 32 bytes along the swizzled row -> +2 in the address field
              umma_tf32(d_tmem, a_hi + 2 * j, b_hi + 2 * j, TC_IDESC, (cc | j) != 0);
              umma_tf32(d_tmem, a_lo + 2 * j, b_hi + 2 * j, TC_IDESC, 1);
              umma_tf32(d_tmem, a_hi + 2 * j, b_lo + 2 * j, TC_IDESC, 1);
            }
            umma_commit(&empty[stage]);   // frees the smem stage once these MMAs have read it
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tmem_full[acc]);    // accumulator of bin k complete
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + TC_GEN_THREADS / 32) {
    // ===== steering generators (B operand) =====
    const int g = tid - 128;
    const int dl = g & (TC_BD - 1), grp = g >> 7;   // direction within the tile; which half of the microphone pairs of a chunk
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ks = item / (p.n_tt * p.n_dt), dt = item % p.n_dt;
      const int k_begin = ks * p.bins_per_range, k_end = min(p.K, k_begin + p.bins_per_range);
      const int d = min(dt * TC_BD + dl, p.D - 1);
      // 32-bit fixed-point turns per bin of this thread's (direction, microphone) elements; k * fx wraps exactly mod 1 turn
      uint32_t fx[NKC * 8];
#pragma unroll
      for (int cc = 0; cc < NKC; ++cc)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int m = cc * 16 + grp * 8 + u;
          const uint64_t f = p.mic_fx[(size_t)d * p.M + m];
          fx[cc * 8 + u] = (uint32_t)((f + 0x80000000ull) >> 32);
        }
      const uint32_t row_re = (uint32_t)dl * 128u, row_im = (uint32_t)(TC_BD + dl) * 128u;
      const uint32_t sw = (uint32_t)(dl & 7);   // (128 + dl) & 7 == dl & 7
      for (int k = k_begin; k < k_end; ++k)
#pragma unroll
        for (int cc = 0; cc < NKC; ++cc) {
          float c[8], s[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int32_t ph = (int32_t)(fx[cc * 8 + u] * (uint32_t)k);                       // signed turns * 2^32
            __sincosf((float)ph * 1.4629180792671596e-09f, &s[u], &c[u]);                     // 2 pi / 2^32
          }
          mbar_wait_bounded(&empty[stage], phase ^ 1);
          unsigned char *bh = smem + stage * TC_STAGE_BYTES + 2 * TC_A_BYTES, *bl = bh + TC_B_BYTES;
#pragma unroll
          for (int u = 0; u < 4; ++u) {   // two microphones = one 16-byte chunk of the Re row and of the Im row
            const float ch0 = tf32_hi(c[2 * u]), sh0 = tf32_hi(s[2 * u]), ch1 = tf32_hi(c[2 * u + 1]), sh1 = tf32_hi(s[2 * u + 1]);
            const float cl0 = c[2 * u] - ch0, sl0 = s[2 * u] - sh0, cl1 = c[2 * u + 1] - ch1, sl1 = s[2 * u + 1] - sh1;
            const uint32_t q = ((uint32_t)(grp * 4 + u) ^ sw) << 4;   // 16-byte chunk, 128B swizzle
            *reinterpret_cast<float4 *>(bh + row_re + q) = make_float4(ch0, -sh0, ch1, -sh1);
            *reinterpret_cast<float4 *>(bh + row_im + q) = make_float4(sh0, ch0, sh1, ch1);
            *reinterpret_cast<float4 *>(bl + row_re + q) = make_float4(cl0, -sl0, cl1, -sl1);
            *reinterpret_cast<float4 *>(bl + row_im + q) = make_float4(sl0, cl0, sl1, cl1);
          }
          fence_async_smem();            // generic-proxy stores -> visible to the tensor core (async proxy)
          mbar_arrive_warp(&full_b[stage]);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp >= 4 + TC_GEN_THREADS / 32) {
    // ===== epilogue: TMEM -> registers, square, accumulate over bins =====
    const int e = warp - (4 + TC_GEN_THREADS / 32);
    const int quarter = warp & 3, half = e >> 2;   // TMEM lanes 32*quarter..+31 are the ones this warp may touch; which 64 directions
    constexpr int HD = TC_BD / 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ks = item / (p.n_tt * p.n_dt), tt = (item / p.n_dt) % p.n_tt, dt = item % p.n_dt;
      const int k_begin = ks * p.bins_per_range, k_end = min(p.K, k_begin + p.bins_per_range);
      const long long t = (long long)tt * TC_BM + quarter * 32 + lane;
      float sum[HD];
#pragma unroll
      for (int i = 0; i < HD; ++i) sum[i] = 0.f;
      for (int k = k_begin; k < k_end; ++k) {
        mbar_wait_bounded(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + half * HD);
#pragma unroll
        for (int j = 0; j < HD / 16; ++j) {
          float vr[16], vi[16];
          tmem_ld16(taddr + j * 16, vr);              // Re Y of 16 directions
          tmem_ld16(taddr + TC_BD + j * 16, vi);      // Im Y of the same directions
#pragma unroll
          for (int i = 0; i < 16; ++i) sum[j * 16 + i] = fmaf(vi[i], vi[i], fmaf(vr[i], vr[i], sum[j * 16 + i]));
        }
        tc_fence_before();
        mbar_arrive_warp(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (t < p.BT) {
        const int d0 = dt * TC_BD + half * HD;
        float *dst = p.partial + (((long long)ks * p.BT + t) * p.D) + d0;
        const int nd = min(HD, p.D - d0);
        if (nd == HD && (p.D & 3) == 0) {
#pragma unroll
          for (int i = 0; i < HD; i += 4) *reinterpret_cast<float4 *>(dst + i) = make_float4(sum[i], sum[i + 1], sum[i + 2], sum[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < HD; ++i) if (i < nd) dst[i] = sum[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// srp[t][d] = 1/2 ( sum_s partial[s][t][d] - nzsum[t] ), partial maps added in index order
__global__ void srp_reduce_kernel(const float *__restrict__ partial, int n_part, long long BT, int D, const float *__restrict__ nzsum,
                                  float *__restrict__ srp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BT * D) return;
  float acc = 0.f;
  for (int s = 0; s < n_part; ++s) acc += partial[(long long)s * BT * D + i];
  srp[i] = 0.5f * (acc - nzsum[i / D]);
}

// ---- small arrays: M = 16 microphones, D <= 64 directions (the mcbeam-shaped streams of BASELINE config 5) ---------------------
// With one direction tile the frames operand is read exactly once, so the bin-major hi / lo copy of srp_prepare_kernel (write
// 2 x 4*2M*K bytes per frame, read them again) costs more than the contraction itself.  This variant has no pre-pass and no TMA:
//   warps 4-11 (256 producer threads) build BOTH operands of a bin in shared memory: the frames operand straight from the
//               spectra (thread = (frame row, 8 microphones): 16-byte loads of two bins at a time, one bin pair prefetched,
//               whiten, 3xTF32 split, swizzled 16-byte stores; the other half of every 32-byte sector is an L1 hit one
//               iteration later - the kernel keeps 2 stages so that ~96 KB stay L1), and the generated steering operand
//               (thread = (direction, 4 microphones)); they also count the non-zero (bin, microphone) entries per frame (nz term)
//   warp 1      MMA issuer: 12 tcgen05.mma.kind::tf32 (M128 N128 K8) per bin into one of two 128-column TMEM accumulators
//   warps 12-19 epilogue: |Y|^2 accumulated over the bins of the work item (frame tile x bin range), partial map stored
constexpr int TS_BD = 64;
constexpr int TS_STAGES = 2;
constexpr int TS_A_BYTES = TC_BM * TC_KC * 4;          // 16 KB
constexpr int TS_B_BYTES = 2 * TS_BD * TC_KC * 4;      // 16 KB
constexpr int TS_STAGE_BYTES = 2 * TS_A_BYTES + 2 * TS_B_BYTES;   // 64 KB
constexpr int TS_SMEM = TS_STAGES * TS_STAGE_BYTES + 1024 + 256;
// warp 0: MMA issuer; warps 1-16: 512 producer threads (the operand build is the critical path: 16 warps hide its latencies where 8
// ran at 0.43 IPC); warps 17-24: epilogue (any 4 consecutive warps cover the four TMEM lane quarters)
constexpr int TS_PROD_THREADS = 512, TS_EPI_THREADS = 256, TS_THREADS = 32 + TS_PROD_THREADS + TS_EPI_THREADS;
constexpr uint32_t TS_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

struct TsParams {
  const float2 *spec;    // [BT][16][KP]
  long long BT;
  int D, K, KP;
  int n_tt, n_ks, bins_per_range;   // bins_per_range is even: bin pairs never straddle two ranges
  const uint64_t *mic_fx;           // [D][16]
  float *partial;                   // [n_ks][BT][D]
  float *nzsum;                     // [BT], zeroed before the launch
};

__global__ void __launch_bounds__(TS_THREADS, 1) srp_tc_small_kernel(const TsParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array itself: an integer round-trip of the pointer loses the address
  // space and the operand-tile stores become generic ST.E.128 (long-scoreboard WAR stalls in the producers, ncu s4_cfg5_small)
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TS_STAGES * TS_STAGE_BYTES);
  uint64_t *full = bars, *empty = bars + TS_STAGES, *tmem_full = bars + 2 * TS_STAGES, *tmem_empty = bars + 2 * TS_STAGES + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TS_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < TS_STAGES; ++s) { mbar_init(&full[s], TS_PROD_THREADS / 32); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], TS_EPI_THREADS / 32); }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.n_tt * p.n_ks;

  if (warp == 0) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ks = item / p.n_tt;
        const int k_begin = ks * p.bins_per_range, k_end = min(p.K, k_begin + p.bins_per_range);
        for (int k = k_begin; k < k_end; ++k) {
          mbar_wait_bounded(&tmem_empty[acc], acc_phase ^ 1);
          mbar_wait_bounded(&full[stage], phase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * 128u;
          const uint32_t sa = smem_u32(smem + stage * TS_STAGE_BYTES);
          const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + TS_A_BYTES);
          const uint64_t b_hi = umma_desc_sw128(sa + 2 * TS_A_BYTES), b_lo = umma_desc_sw128(sa + 2 * TS_A_BYTES + TS_B_BYTES);
#pragma unroll
          for (int j = 0; j < TC_KC / 8; ++j) {
            umma_tf32(d_tmem, a_hi + 2 * j, b_hi + 2 * j, TS_IDESC, j != 0);
            umma_tf32(d_tmem, a_lo + 2 * j, b_hi + 2 * j, TS_IDESC, 1);
            umma_tf32(d_tmem, a_hi + 2 * j, b_lo + 2 * j, TS_IDESC, 1);
          }
          umma_commit(&empty[stage]);
          umma_commit(&tmem_full[acc]);
          if (++stage == TS_STAGES) { stage = 0; phase ^= 1; }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp <= TS_PROD_THREADS / 32) {
    // ===== producers: frames operand from the spectra + generated steering operand =====
    const int g = tid - 32;
    const int row = g & (TC_BM - 1), qm = g >> 7;     // frames operand: frame row of the tile, microphones 4*qm .. 4*qm+3
    const int dl = g & (TS_BD - 1), grp = g >> 6;     // steering operand: direction, microphones 2*grp, 2*grp+1
    const int d = min(dl, p.D - 1);
    uint32_t fx[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) fx[u] = (uint32_t)((p.mic_fx[(size_t)d * 16 + grp * 2 + u] + 0x80000000ull) >> 32);
    const uint32_t a_row = (uint32_t)row * 128u, a_sw = (uint32_t)(row & 7);
    const uint32_t b_re = (uint32_t)dl * 128u, b_im = (uint32_t)(TS_BD + dl) * 128u, b_sw = (uint32_t)(dl & 7);
    const int kp4 = p.KP >> 1;   // float4 (two bins) per spectrum row
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ks = item / p.n_tt, tt = item - ks * p.n_tt;
      const int k_begin = ks * p.bins_per_range, k_end = min(p.K, k_begin + p.bins_per_range);
      const long long t = (long long)tt * TC_BM + row;
      const bool valid = t < p.BT;
      const float4 *src = reinterpret_cast<const float4 *>(p.spec + ((valid ? t : 0) * 16 + qm * 4) * p.KP);
      float4 cur[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) cur[u] = __ldg(src + (size_t)u * kp4 + (k_begin >> 1));   // rows past BT read frame 0 and are zeroed below
      float cnt = 0.f;
      for (int k = k_begin; k < k_end; k += 2) {
        float4 nxt[4];
        // always a load (the last pair of a range re-reads itself): a select against zero made the compiler clear the destination
        // registers first, and that write waited on the address reads of the loads still queued in the LSU (ncu s4_cfg5_small2)
        const int kn = (k + 2 < k_end) ? k + 2 : k;
#pragma unroll
        for (int u = 0; u < 4; ++u) nxt[u] = __ldg(src + (size_t)u * kp4 + (kn >> 1));
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int kb = k + kk;
          if (kb < k_end) {
            float ah[8], al[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float2 w = whiten(kk ? make_float2(cur[u].z, cur[u].w) : make_float2(cur[u].x, cur[u].y));
              if (!valid) w = make_float2(0.f, 0.f);
              cnt += (w.x != 0.f || w.y != 0.f) ? 1.f : 0.f;
              ah[2 * u] = tf32_hi(w.x); ah[2 * u + 1] = tf32_hi(w.y);
              al[2 * u] = w.x - ah[2 * u]; al[2 * u + 1] = w.y - ah[2 * u + 1];
            }
            float c[2], sn[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int32_t ph = (int32_t)(fx[u] * (uint32_t)kb);
              __sincosf((float)ph * 1.4629180792671596e-09f, &sn[u], &c[u]);
            }
            mbar_wait_bounded(&empty[stage], phase ^ 1);
            unsigned char *st = smem + stage * TS_STAGE_BYTES;
            unsigned char *ahp = st + a_row, *alp = st + TS_A_BYTES + a_row;
#pragma unroll
            for (int q = 0; q < 2; ++q) {   // two microphones = one 16-byte chunk of the frame's row
              const uint32_t off = ((uint32_t)(qm * 2 + q) ^ a_sw) << 4;
              *reinterpret_cast<float4 *>(ahp + off) = make_float4(ah[4 * q], ah[4 * q + 1], ah[4 * q + 2], ah[4 * q + 3]);
              *reinterpret_cast<float4 *>(alp + off) = make_float4(al[4 * q], al[4 * q + 1], al[4 * q + 2], al[4 * q + 3]);
            }
            unsigned char *bh = st + 2 * TS_A_BYTES, *bl = bh + TS_B_BYTES;
            {
              constexpr int q = 0;
              const float ch0 = tf32_hi(c[2 * q]), sh0 = tf32_hi(sn[2 * q]), ch1 = tf32_hi(c[2 * q + 1]), sh1 = tf32_hi(sn[2 * q + 1]);
              const float cl0 = c[2 * q] - ch0, sl0 = sn[2 * q] - sh0, cl1 = c[2 * q + 1] - ch1, sl1 = sn[2 * q + 1] - sh1;
              const uint32_t off = ((uint32_t)grp ^ b_sw) << 4;
              *reinterpret_cast<float4 *>(bh + b_re + off) = make_float4(ch0, -sh0, ch1, -sh1);
              *reinterpret_cast<float4 *>(bh + b_im + off) = make_float4(sh0, ch0, sh1, ch1);
              *reinterpret_cast<float4 *>(bl + b_re + off) = make_float4(cl0, -sl0, cl1, -sl1);
              *reinterpret_cast<float4 *>(bl + b_im + off) = make_float4(sl0, cl0, sl1, cl1);
            }
            fence_async_smem();
            mbar_arrive_warp(&full[stage]);
            if (++stage == TS_STAGES) { stage = 0; phase ^= 1; }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
      }
      if (valid && cnt != 0.f) atomicAdd(p.nzsum + t, cnt);   // small integers: the float sum is exact in any order
    }
  } else {
    // ===== epilogue =====
    const int e = warp - (1 + TS_PROD_THREADS / 32);
    const int quarter = warp & 3, half = e >> 2;   // TMEM lanes 32*quarter..+31; directions 32*half..+31
    constexpr int HD = TS_BD / 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ks = item / p.n_tt, tt = item - ks * p.n_tt;
      const int k_begin = ks * p.bins_per_range, k_end = min(p.K, k_begin + p.bins_per_range);
      const long long t = (long long)tt * TC_BM + quarter * 32 + lane;
      float sum[HD];
#pragma unroll
      for (int i = 0; i < HD; ++i) sum[i] = 0.f;
      for (int k = k_begin; k < k_end; ++k) {
        mbar_wait_bounded(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 128 + half * HD);
#pragma unroll
        for (int j = 0; j < HD / 8; ++j) {   // 8 columns at a time: the 800-thread CTA leaves 72 registers per thread
          float vr[8], vi[8];
          tmem_ld8(taddr + j * 8, vr);
          tmem_ld8(taddr + TS_BD + j * 8, vi);
#pragma unroll
          for (int i = 0; i < 8; ++i) sum[j * 8 + i] = fmaf(vi[i], vi[i], fmaf(vr[i], vr[i], sum[j * 8 + i]));
        }
        tc_fence_before();
        mbar_arrive_warp(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (t < p.BT) {
        const int d0 = half * HD;
        float *dst = p.partial + (((long long)ks * p.BT + t) * p.D) + d0;
        const int nd = min(HD, p.D - d0);
#pragma unroll
        for (int i = 0; i < HD; ++i) if (i < nd) dst[i] = sum[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

// work decomposition of the small variant: frame tiles x bin ranges, the range count chosen for full waves of persistent CTAs
static bool ts_supported(int M, int D) { return M == 16 && D >= 1 && D <= TS_BD; }
static void ts_plan(long long BT, int N, int D, int sms, TsParams &p) {
  const int K = N / 2 + 1;
  p.BT = BT; p.D = D; p.K = K; p.KP = spec_pitch(N);
  p.n_tt = (int)((BT + TC_BM - 1) / TC_BM);
  int best_bpr = (K + 1) & ~1;
  double best_score = -1.0;
  for (int bpr = 8; bpr <= ((K + 1) & ~1); bpr += 2) {
    const int n_ks = (K + bpr - 1) / bpr;
    const long long items = (long long)p.n_tt * n_ks;
    const long long grid = items < sms ? items : sms;
    const long long rounds = (items + grid - 1) / grid;
    // busy fraction of the CTA slots, minus a small charge per work item (pipeline fill, partial map store)
    const double score = (double)items / (double)(rounds * sms) - 0.004 * (double)n_ks;
    if (score > best_score) { best_score = score; best_bpr = bpr; }
  }
  p.bins_per_range = best_bpr;
  p.n_ks = (K + best_bpr - 1) / best_bpr;
}
static size_t ts_workspace_bytes(long long BT, int N, int D) {
  auto al = [](size_t n) { return (n + 1023) & ~(size_t)1023; };
  const int K = N / 2 + 1;
  return al((size_t)((K + 7) / 8) * BT * D * 4) + al((size_t)BT * 4);   // partial maps for the largest possible range count + nzsum
}
static int ts_launch(const float2 *spec, long long BT, int N, const uint64_t *mic_fx, int D, float *srp, void *workspace, size_t ws_bytes,
                     cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  TsParams p;
  ts_plan(BT, N, D, sms, p);
  auto al = [](size_t n) { return (n + 1023) & ~(size_t)1023; };
  const size_t part_bytes = al((size_t)p.n_ks * BT * D * 4);
  if (part_bytes + al((size_t)BT * 4) > ws_bytes) return mcag_set_error(4, "srp_tensor: workspace too small");
  unsigned char *w = static_cast<unsigned char *>(workspace);
  p.spec = spec; p.mic_fx = mic_fx; p.partial = reinterpret_cast<float *>(w); p.nzsum = reinterpret_cast<float *>(w + part_bytes);
  if (cudaMemsetAsync(p.nzsum, 0, (size_t)BT * 4, st) != cudaSuccess) return mcag_set_cuda_error(cudaGetLastError());
  // per launch, like launch_tc: function attributes are per device and a process may drive more than one
  cudaFuncSetAttribute(srp_tc_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM);
  cudaFuncSetAttribute(srp_tc_small_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 58);   // 132 KB shared, ~96 KB stay L1 (3 stages / no L1: 0.57 ms instead of 0.31)
  const long long items = (long long)p.n_tt * p.n_ks;
  srp_tc_small_kernel<<<(unsigned)(items < sms ? items : sms), TS_THREADS, TS_SMEM, st>>>(p);
  MCAG_CHECK_LAUNCH();
  const long long n = BT * D;
  srp_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.partial, p.n_ks, BT, D, p.nzsum, srp);
  MCAG_CHECK_LAUNCH();
  return 0;
}

static int encode_map(CUtensorMap *map, const float *base, long long BTpad, int M, int K) {
  EncodeTiledFn cuTensorMapEncodeTiled = encode_tiled_fn();
  if (!cuTensorMapEncodeTiled) return mcag_set_error(2, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)(2 * M), (cuuint64_t)BTpad, (cuuint64_t)K};
  cuuint64_t strides[2] = {(cuuint64_t)(2 * M) * 4, (cuuint64_t)BTpad * (2 * M) * 4};
  cuuint32_t box[3] = {TC_KC, TC_BM, 1}, estr[3] = {1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mcag_set_error(2, "cuTensorMapEncodeTiled failed");
  return 0;
}

template <int NKC> static void launch_tc(int grid, cudaStream_t st, const CUtensorMap &hi, const CUtensorMap &lo, const TcParams &p) {
  auto kern = srp_tc_kernel<NKC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  kern<<<grid, TC_THREADS, TC_SMEM, st>>>(hi, lo, p);
}

// Work decomposition shared by the workspace query and the launcher.
static void tc_plan(long long BT, int M, int N, int D, TcParams &p, long long &BTpad) {
  BTpad = (BT + TC_BM - 1) / TC_BM * TC_BM;
  const int K = N / 2 + 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  p.BT = BT; p.D = D; p.M = M; p.K = K;
  p.n_tt = (int)(BTpad / TC_BM); p.n_dt = (D + TC_BD - 1) / TC_BD;
  // bin ranges: 3 to 8 waves of work items for the persistent CTAs (one per SM), at least 16 bins each; among those the count that
  // leaves the last wave fullest (cfg4 on one GPU: 232 tiles x 3 ranges = 4.7 waves, x 5 = 7.8; the sharded grid at 8 GPUs: 32 tiles x
  // 19 ranges = 4.1 waves, a fifth wave one ninth full, x 18 = 3.9)
  const long long tiles = (long long)p.n_tt * p.n_dt;
  const int n_max = K / 16 > 0 ? K / 16 : 1;
  int best = (int)((4LL * sms + tiles - 1) / tiles);   // small problems (never 3 waves): as many ranges as it takes to reach ~4 waves, capped
  if (best > n_max) best = n_max;
  if (best < 1) best = 1;
  double best_eff = -1.0;
  for (int n = 1; n <= n_max; ++n) {
    const int bpr = (K + n - 1) / n, nn = (K + bpr - 1) / bpr;   // the count the rounding of bins_per_range really gives
    const long long items = tiles * nn, waves = (items + sms - 1) / sms;
    if (waves < 3 || waves > 8) continue;
    const double eff = (double)items / (double)(waves * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = n; }
  }
  p.bins_per_range = (K + best - 1) / best;
  p.n_ks = (K + p.bins_per_range - 1) / p.bins_per_range;
}
bool k_srp_tensor_supported(int M) { return M % 16 == 0 && M >= 16 && M <= 64; }
// bytes of scratch k_srp_tensor_ws needs for up to BT frames: [U_hi | U_lo | partial | nzsum], each 1 KB aligned
size_t k_srp_tensor_workspace_bytes(long long BT, int M, int N, int D) {
  if (!k_srp_tensor_supported(M) || BT <= 0) return 0;
  if (ts_supported(M, D)) return ts_workspace_bytes(BT, N, D);
  TcParams p; long long BTpad;
  tc_plan(BT, M, N, D, p, BTpad);
  auto al = [](size_t n) { return (n + 1023) & ~(size_t)1023; };
  const int K = N / 2 + 1;
  // n_ks shrinks as BT grows, so the partial maps of any smaller call fit when sized with the largest bin-range count
  return 2 * al((size_t)K * BTpad * 2 * M * 4) + al((size_t)(K / 16 > 0 ? K / 16 : 1) * BT * D * 4) + al((size_t)BT * 4);
}

// Supported shapes: M in {16, 32, 48, 64}; anything else runs the CUDA-core tile kernel (still on the GPU).
int k_srp_tensor_ws(const float2 *spec, int B, int T, int M, int N, const uint64_t *mic_fx, int D, float *srp, void *workspace, size_t ws_bytes,
                    cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  if (!k_srp_tensor_supported(M)) return k_srp_channel(spec, B, T, M, N, mic_fx, D, srp, st);
  const long long BT = (long long)B * T;
  if (ts_supported(M, D)) return ts_launch(spec, BT, N, mic_fx, D, srp, workspace, ws_bytes, st);
  TcParams p; long long BTpad;
  tc_plan(BT, M, N, D, p, BTpad);
  p.mic_fx = mic_fx;
  const int K = p.K;
  auto al = [](size_t n) { return (n + 1023) & ~(size_t)1023; };
  const size_t u_bytes = al((size_t)K * BTpad * 2 * M * sizeof(float)), part_bytes = al((size_t)p.n_ks * BT * D * sizeof(float));
  if (2 * u_bytes + part_bytes + al((size_t)BT * 4) > ws_bytes) return mcag_set_error(4, "srp_tensor: workspace too small");
  unsigned char *w = static_cast<unsigned char *>(workspace);
  float *Uhi = reinterpret_cast<float *>(w), *Ulo = reinterpret_cast<float *>(w + u_bytes), *partial = reinterpret_cast<float *>(w + 2 * u_bytes),
        *nzsum = reinterpret_cast<float *>(w + 2 * u_bytes + part_bytes);
  p.partial = partial;
  srp_prepare_kernel<<<(unsigned)BTpad, 256, 0, st>>>(spec, BT, BTpad, M, N, Uhi, Ulo, nzsum);
  MCAG_CHECK_LAUNCH();
  CUtensorMap map_hi, map_lo;
  OK_RC(encode_map(&map_hi, Uhi, BTpad, M, K));
  OK_RC(encode_map(&map_lo, Ulo, BTpad, M, K));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)p.n_tt * p.n_dt * p.n_ks;
  const int grid = (int)(items < sms ? items : sms);
  switch (M / 16) {
    case 1: launch_tc<1>(grid, st, map_hi, map_lo, p); break;
    case 2: launch_tc<2>(grid, st, map_hi, map_lo, p); break;
    case 3: launch_tc<3>(grid, st, map_hi, map_lo, p); break;
    default: launch_tc<4>(grid, st, map_hi, map_lo, p); break;
  }
  MCAG_CHECK_LAUNCH();
  const long long n = BT * D;
  srp_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, p.n_ks, BT, D, nzsum, srp);
  MCAG_CHECK_LAUNCH();
  return 0;
}

// kernel-level entry without a caller-owned workspace: stream-ordered scratch from the device's memory pool
int k_srp_tensor(const float2 *spec, int B, int T, int M, int N, const uint64_t *mic_fx, int D, float *srp, cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  if (!k_srp_tensor_supported(M)) return k_srp_channel(spec, B, T, M, N, mic_fx, D, srp, st);
  static bool pool_kept = false;
  if (!pool_kept) {   // keep freed scratch in the pool across synchronisations instead of returning it to the driver
    int dev = 0; cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) { unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
    pool_kept = true;
  }
  const size_t bytes = k_srp_tensor_workspace_bytes((long long)B * T, M, N, D);
  void *ws = nullptr;
  if (cudaMallocAsync(&ws, bytes, st) != cudaSuccess) return mcag_set_cuda_error(cudaGetLastError());
  const int rc = k_srp_tensor_ws(spec, B, T, M, N, mic_fx, D, srp, ws, bytes, st);
  cudaFreeAsync(ws, st);
  return rc;
}

}  // namespace mcag

extern "C" int mcag_k_srp_tensor(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_mic_fx, int D, float *d_srp, void *stream) {
  return mcag::k_srp_tensor((const float2 *)d_spec, B, T, M, N, d_mic_fx, D, d_srp, (cudaStream_t)stream);
}

// Shared-memory Stockham FFT engine (complex, NC = 128/256/512/1024 points) for the real FFTs of the STFT,
// the inverse STFT and the GCC-PHAT lag transform.  NC/8 threads cooperate on one transform; each thread
// owns 8 points per pass (one radix-8, two radix-4 or four radix-2 butterflies) in registers, passes exchange
// through a padded float2 buffer in shared memory.  Real N = 2*NC transforms use the packed-complex trick.
//
// Twiddles come from two tables computed in double on the host (capi.cu: mcag_k_twiddles), stored back to back:
//   tw[n]  = exp(-2*pi*i*n/N), n < N/2 (= NC entries): the real <-> packed-complex pre/post-processing
//   twp[slot][j], j < NC/8: the inter-pass twiddles of thread j, one slot per (pass >= 1, block b, r >= 1).  A thread
//   multiplies by the same factors in every transform it ever runs, so the table is laid out per thread: a warp reads
//   consecutive float2 (conflict-free) instead of gathering tw[k*r*stride] (8-way bank conflicts, ncu r1_cfg2_full).
#pragma once
#include "common.cuh"

namespace mcag {

// buffer index swizzle: XOR with bits 3..6 keeps every access pattern of the Stockham passes (stride-8 stores, contiguous
// loads at stride NC/R, 64*(j>>3)+(j&7) stores) at the ideal two wavefronts per 64-bit warp access for all four sizes
// (checked exhaustively on the host; the old i + (i>>3) padding cost 3-4 wavefronts on the contiguous loads).
__host__ __device__ constexpr int fft_pad(int i) { return i ^ ((i >> 3) & 15); }
__host__ __device__ constexpr int fft_buf_len(int NC) { return NC; }

// Transform buffers are addressed by their 32-bit shared-memory address and must be aligned to their size (8*NC bytes): then
// base + 8*idx == base ^ 8*idx, and together with the XOR-affine swizzle below every point of a pass is ONE LOP3 away from a
// per-pass address (fft_align_smem carves such a region out of the dynamic shared memory).
typedef uint32_t fft_buf_t;
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ float2 fft_buf_get(fft_buf_t buf, int i) { return lds64(buf ^ (8u * (uint32_t)fft_pad(i))); }
// first byte of the dynamic shared memory rounded up to `align` (a power of two); kernels add `align` bytes to their request
__device__ __forceinline__ unsigned char *fft_align_smem(unsigned char *raw, uint32_t align) {
  const uint32_t a = smem_u32(raw);
  return raw + (((a + align - 1u) & ~(align - 1u)) - a);
}

template <int NC> struct FftPlan;
template <> struct FftPlan<128>  { static constexpr int NP = 3; static constexpr int R[4] = {8, 8, 2, 1}; };
template <> struct FftPlan<256>  { static constexpr int NP = 3; static constexpr int R[4] = {8, 8, 4, 1}; };
template <> struct FftPlan<512>  { static constexpr int NP = 3; static constexpr int R[4] = {8, 8, 8, 1}; };
template <> struct FftPlan<1024> { static constexpr int NP = 4; static constexpr int R[4] = {8, 8, 8, 2}; };

// slots of the per-thread twiddle table: (8/R) * (R-1) per pass after the first
template <int NC> __host__ __device__ constexpr int fft_twp_slots() {
  using P = FftPlan<NC>;
  int n = 0;
  for (int p = 1; p < P::NP; ++p) n += (8 / P::R[p]) * (P::R[p] - 1);
  return n;
}
__host__ __device__ constexpr int fft_twp_slots_rt(int NC) {
  return NC == 128 ? fft_twp_slots<128>() : NC == 256 ? fft_twp_slots<256>() : NC == 512 ? fft_twp_slots<512>() : fft_twp_slots<1024>();
}
// float2 entries of the combined table for frame size N: tw[N/2] then twp[slots][N/16]
__host__ __device__ constexpr int fft_table_len(int N) { return N / 2 + fft_twp_slots_rt(N / 2) * (N / 16); }

// sync among the NC/8 threads of one transform: a warp (or less) syncs itself, larger groups use a named barrier
template <int TPF> __device__ __forceinline__ void group_sync(int group) {
  if constexpr (TPF == 32) {
    __syncwarp();
  } else if constexpr (TPF < 32) {
    // several transforms share one warp and may diverge from each other: sync only this transform's lanes
    const unsigned lane = threadIdx.x & 31u;
    __syncwarp((((1u << TPF) - 1u)) << (lane & ~(unsigned)(TPF - 1)));
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(TPF) : "memory");
  }
}

// W_N^(n) for n in [0, N): table holds the first half
template <bool INV> __device__ __forceinline__ float2 tw_lookup(const float2 *tw, int n, int half) {
  float2 w = n < half ? tw[n] : tw[n - half];
  if (n >= half) { w.x = -w.x; w.y = -w.y; }
  if (INV) w.y = -w.y;
  return w;
}

template <bool INV> __device__ __forceinline__ void dft2(float2 &a, float2 &b) {
  float2 t = a; a = cadd(t, b); b = csub(t, b);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ float2 rot90(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

template <bool INV> __device__ __forceinline__ void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
  float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = rot90<INV>(csub(v1, v3));
  v0 = cadd(a0, a2); v2 = csub(a0, a2); v1 = cadd(a1, a3); v3 = csub(a1, a3);
}
template <bool INV> __device__ __forceinline__ void dft8(float2 *v) {
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4<INV>(e0, e1, e2, e3);
  dft4<INV>(o0, o1, o2, o3);
  const float h = 0.70710678118654752440f;
  // W8^1 = (1 -+ i)/sqrt2, W8^2 = -+i, W8^3 = (-1 -+ i)/sqrt2
  float2 t1 = INV ? make_float2((o1.x - o1.y) * h, (o1.x + o1.y) * h) : make_float2((o1.x + o1.y) * h, (o1.y - o1.x) * h);
  float2 t2 = rot90<INV>(o2);
  float2 t3 = INV ? make_float2((-o3.x - o3.y) * h, (o3.x - o3.y) * h) : make_float2((o3.y - o3.x) * h, (-o3.x - o3.y) * h);
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
  v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
  v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}

// The swizzle is XOR-affine in the thread index for every access pattern of the passes (checked exhaustively for all four
// plans): fft_pad(idx(j, b, r)) == fft_pad(idx(j, b, 0)) ^ C(b, r) with C independent of j.  Each pass therefore swizzles ONE
// index per block at run time and reaches its other R-1 points with a compile-time XOR (one LOP3 instead of add + shift +
// and + xor per access; the integer pipe was 44 % busy with index arithmetic before, ncu r1b_cfg2_fused).
template <int NC, int R, int NS> __host__ __device__ constexpr int fft_store_xor(int b, int r) {
  const int jj = b * (NC / 8), k = jj & (NS - 1), j0 = (jj - k) * R + k;
  return fft_pad(j0 + r * NS) ^ fft_pad(j0);
}
template <int NC, int R> __host__ __device__ constexpr int fft_load_xor(int b, int r) {
  const int jj = b * (NC / 8);
  return fft_pad(jj + r * (NC / R)) ^ fft_pad(jj);
}

// One Stockham pass over the 8 points this thread holds.  On entry v[b*R + r] = in[jj_b + r*NC/R] with
// jj_b = j + b*NC/8; on exit the results are stored to buf at their autosort positions.
template <int NC, int R, int NS, int SLOT0, bool INV>
__device__ __forceinline__ void fft_pass_store(float2 *v, fft_buf_t buf, const float2 *twp, int j) {
  constexpr int NB = 8 / R, TPF = NC / 8;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int jj = j + b * (NC / 8);
    const int k = jj & (NS - 1);
    float2 *u = v + b * R;
    if (NS > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) {
        // the factor depends on the thread only through k = j & (NS-1) when NS < TPF: lanes with the same k then read the SAME
        // table entry (a broadcast: 8 distinct float2 per warp in the NS = 8 pass is one wavefront instead of two)
        float2 w = twp[(SLOT0 + b * (R - 1) + (r - 1)) * TPF + (NS < TPF ? (j & (NS - 1)) : j)];
        if (INV) w.y = -w.y;
        u[r] = cmul(u[r], w);
      }
    }
    if constexpr (R == 8) dft8<INV>(u);
    else if constexpr (R == 4) dft4<INV>(u[0], u[1], u[2], u[3]);
    else dft2<INV>(u[0], u[1]);
    const uint32_t base = buf ^ (8u * (uint32_t)fft_pad((jj - k) * R + k));
#pragma unroll
    for (int r = 0; r < R; ++r) sts64(base ^ (8u * (uint32_t)fft_store_xor<NC, R, NS>(b, r)), u[r]);
  }
}
template <int NC, int R> __device__ __forceinline__ void fft_pass_load(float2 *v, fft_buf_t buf, int j) {
  constexpr int NB = 8 / R;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const uint32_t base = buf ^ (8u * (uint32_t)fft_pad(j + b * (NC / 8)));
#pragma unroll
    for (int r = 0; r < R; ++r) v[b * R + r] = lds64(base ^ (8u * (uint32_t)fft_load_xor<NC, R>(b, r)));
  }
}

// first slot of pass `pass` (>= 1) in the per-thread twiddle table
template <int NC> __host__ __device__ constexpr int fft_slot0(int pass) {
  using P = FftPlan<NC>;
  int n = 0;
  for (int p = 1; p < pass; ++p) n += (8 / P::R[p]) * (P::R[p] - 1);
  return n;
}

// All passes but the last.  The caller has already placed the first pass's inputs in v[] (v[r] = in[j + r*NC/8], the first
// pass is always radix 8), so loading, windowing and packing fuse into the caller.  On return the inputs of the last pass
// sit in buf (in[jj + r*NC/R_last] at fft_pad(jj + r*NC/R_last)) and the group is synchronised.
// `twp` is the per-thread twiddle table (shared memory, [slots][NC/8]).
template <int NC, bool INV>
__device__ __forceinline__ void fft_run_head(float2 *v, fft_buf_t buf, const float2 *twp, int j, int group) {
  using P = FftPlan<NC>;
  constexpr int TPF = NC / 8;
  fft_pass_store<NC, 8, 1, 0, INV>(v, buf, twp, j);
  group_sync<TPF>(group);
  fft_pass_load<NC, P::R[1]>(v, buf, j);
  group_sync<TPF>(group);
  fft_pass_store<NC, P::R[1], 8, fft_slot0<NC>(1), INV>(v, buf, twp, j);
  group_sync<TPF>(group);
  if constexpr (P::NP == 4) {
    fft_pass_load<NC, P::R[2]>(v, buf, j);
    group_sync<TPF>(group);
    fft_pass_store<NC, P::R[2], 8 * P::R[1], fft_slot0<NC>(2), INV>(v, buf, twp, j);
    group_sync<TPF>(group);
  }
}

// Full transform: on return the spectrum sits in buf (natural order, padded indexing) and the group is synchronised.
template <int NC, bool INV>
__device__ __forceinline__ void fft_run(float2 *v, fft_buf_t buf, const float2 *twp, int j, int group) {
  using P = FftPlan<NC>;
  constexpr int TPF = NC / 8, L = P::NP - 1;
  constexpr int NSL = (L == 2) ? 8 * P::R[1] : 8 * P::R[1] * P::R[2];
  fft_run_head<NC, INV>(v, buf, twp, j, group);
  fft_pass_load<NC, P::R[L]>(v, buf, j);
  group_sync<TPF>(group);
  fft_pass_store<NC, P::R[L], NSL, fft_slot0<NC>(L), INV>(v, buf, twp, j);
  group_sync<TPF>(group);
}

// Output-pruned last pass of an INVERSE transform: only the first output row (r = 0) and the last one (r = R-1) of the
// radix-R butterfly of block `b` are formed, straight from the buffer fft_run_head left, without storing anything.  These
// rows hold outputs jj and jj + (R-1)*NC/R (jj = j + b*NC/8): the lowest and highest NC/R outputs, which is where the
// non-negative and the negative lags of a short GCC lag window live.
template <int NC, int B_>
__device__ __forceinline__ void fft_inv_last_pruned(fft_buf_t buf, const float2 *twp, int j, bool want_first, bool want_last, float2 &y_first,
                                                    float2 &y_last) {
  using P = FftPlan<NC>;
  constexpr int TPF = NC / 8, L = P::NP - 1, R = P::R[L], SLOT = fft_slot0<NC>(L);
  const int jj = j + B_ * (NC / 8);
  float2 u[R];
  const uint32_t base = buf ^ (8u * (uint32_t)fft_pad(jj));
#pragma unroll
  for (int r = 0; r < R; ++r) u[r] = lds64(base ^ (8u * (uint32_t)fft_load_xor<NC, R>(B_, r)));
#pragma unroll
  for (int r = 1; r < R; ++r) {
    float2 w = twp[(SLOT + B_ * (R - 1) + (r - 1)) * TPF + j];
    u[r] = cmul(u[r], make_float2(w.x, -w.y));
  }
  if (want_first) {
    float2 a = u[0];
#pragma unroll
    for (int r = 1; r < R; ++r) a = cadd(a, u[r]);
    y_first = a;
  }
  if (want_last) {   // sum_q u_q exp(+2 pi i (R-1) q / R) = sum_q u_q exp(-2 pi i q / R)
    if constexpr (R == 2) {
      y_last = csub(u[0], u[1]);
    } else if constexpr (R == 4) {
      const float2 a = csub(u[0], u[2]), b = csub(u[1], u[3]);   // a - i b
      y_last = make_float2(a.x + b.y, a.y - b.x);
    } else {
      const float h = 0.70710678118654752440f;
      const float2 a = csub(u[0], u[4]), b = csub(u[2], u[6]), c = csub(u[1], u[5]), d = csub(u[3], u[7]);
      // a - i b + (1 - i)/sqrt2 c - (1 + i)/sqrt2 d
      const float2 e = make_float2((c.x + c.y) * h, (c.y - c.x) * h), f = make_float2((d.x - d.y) * h, (d.x + d.y) * h);
      y_last = make_float2(a.x + b.y + e.x - f.x, a.y - b.x + e.y - f.y);
    }
  }
}

// copy the combined table (tw then twp) from global to shared memory; the caller syncs
template <int N> __device__ __forceinline__ void fft_load_tables(float2 *s_tab, const float2 *__restrict__ tab_g, int tid, int nthreads) {
  constexpr int LEN = fft_table_len(N);
  const float4 *src = reinterpret_cast<const float4 *>(tab_g);
  float4 *dst = reinterpret_cast<float4 *>(s_tab);
  for (int i = tid; i < LEN / 2; i += nthreads) dst[i] = src[i];
}

}  // namespace mcag

// C ABI (include/mcarray_b200.h): processor handles that chain the kernels of this directory, plus thin extern "C"
// wrappers around the kernel launchers.  Host code only orchestrates: every arithmetic step of the hot path runs in a
// CUDA kernel; there is no CPU fallback (a missing device or a failed launch is an error, never a silent detour).
#include "../../include/mcarray_b200.h"
#include "kernels.h"
#include "fft.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {
thread_local std::string g_err;
}
int mcag_set_error(int code, const char *msg) { g_err = msg ? msg : ""; return code; }
int mcag_set_cuda_error(cudaError_t e) { g_err = std::string("CUDA: ") + cudaGetErrorString(e); return MCAG_ERR_CUDA; }

#define CU(call)                                              \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return mcag_set_cuda_error(e__);  \
  } while (0)
#define OK(call)                   \
  do {                             \
    int r__ = (call);              \
    if (r__ != MCAG_OK) return r__; \
  } while (0)

namespace mcag {

// turns (any real) -> unsigned 0.64 fixed point of frac(turns)
static uint64_t turns_to_fx(double turns) {
  double f = turns - std::floor(turns);          // [0, 1)
  long double s = (long double)f * 18446744073709551616.0L;
  if (s >= 18446744073709551615.0L) return 0;    // wrapped to a full turn
  return (uint64_t)s;
}

// ---- small device helpers owned by the processors ----------------------------------------------------------------
// power gate: SoundLocalisationImpl state machine (BeamformingSeparationAndLocalisation.cpp:55-101,
// BinauralLocalisation.cpp:387-404,429-434).  One thread per stream, sequential over the frames of the call.
struct GateState { double acc; double floor; int samples; int estimated; };

__global__ void __launch_bounds__(256) gate_kernel(const float *__restrict__ chan_pow, const float *__restrict__ chan_raw, int B, int T, int M, int N,
                                                   int use_floor, int ccs_mode, float margin_db, int needed, GateState *__restrict__ gs,
                                                   float *__restrict__ power_db, unsigned char *__restrict__ active,
                                                   unsigned char *__restrict__ est /* optional: _noiseEstimated after the frame's floor update */) {
  // one CTA per stream: thread 0 walks the frames that still feed the noise-floor estimate (a strictly sequential
  // accumulation, at most floor_seconds of audio per stream), then all threads gate the remaining frames in parallel.
  __shared__ GateState s_state;
  __shared__ int s_t0;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    GateState s = gs[b];
    const bool estimate = ccs_mode ? true : (use_floor != 0);   // FreqGCC estimates the floor even when it will not use it (:429)
    int t = 0;
    for (; t < T && !s.estimated && estimate; ++t) {
      if (ccs_mode) {
        const float *cr = chan_raw + ((long long)b * T + t) * M;
        double raw = 0.0;
        for (int m = 0; m < M; ++m) raw += (double)cr[m] / (double)(N + 2);
        s.acc += raw / (double)M * (double)N + 1e-10;
      } else {
        const float *cp = chan_pow + ((long long)b * T + t) * M;
        double lin = 0.0;
        for (int m = 0; m < M; ++m) lin += (double)cp[m];
        s.acc += lin / (double)M * (double)N;
      }
      s.samples += N;
      if (s.samples >= needed) {
        s.estimated = 1;
        s.acc /= (double)s.samples;
        s.acc = 10.0 * log10(s.acc) + (double)margin_db;
      }
      s.floor = s.acc;
      power_db[(long long)b * T + t] = (float)s.floor;
      active[(long long)b * T + t] = use_floor ? 0 : 1;   // power == floor here, so `power > floor` is false
      if (est) est[(long long)b * T + t] = (unsigned char)s.estimated;   // the frame that completes the estimate already counts as silent (:528)
    }
    gs[b] = s;
    s_state = s;
    s_t0 = t;
  }
  __syncthreads();
  const double floor_db = s_state.floor;
  for (int t = s_t0 + threadIdx.x; t < T; t += blockDim.x) {
    const float *cp = chan_pow + ((long long)b * T + t) * M;
    double lin = 0.0;
    for (int m = 0; m < M; ++m) lin += (double)cp[m];
    const double power = 10.0 * log10(lin / (double)M);
    power_db[(long long)b * T + t] = (float)power;
    active[(long long)b * T + t] = (power > floor_db || !use_floor) ? 1 : 0;
    if (est) est[(long long)b * T + t] = 1;
  }
}

// hold the last active selection on gated-off frames (_currentDOA / _prob persist, BSAL.cpp:87-95).  One CTA per stream: every frame
// looks back to the nearest active frame of the call (usually itself), or to the carried state when there is none.
__global__ void __launch_bounds__(128) carry_cells_kernel(const int32_t *__restrict__ raw_idx, const float *__restrict__ raw_prob,
                                                           const unsigned char *__restrict__ active, int B, int T, int S,
                                                           int32_t *__restrict__ cell_state, float *__restrict__ prob_state,
                                                           int32_t *__restrict__ cells, float *__restrict__ prob) {
  const int b = blockIdx.x;
  const unsigned char *act = active + (long long)b * T;
  for (int i = threadIdx.x; i < T * S; i += blockDim.x) {
    const int t = i / S, s = i - t * S;
    int u = t;
    while (u >= 0 && !act[u]) --u;
    const long long o = ((long long)b * T + t) * S + s;
    if (u >= 0) { const long long src = ((long long)b * T + u) * S + s; cells[o] = raw_idx[src]; prob[o] = raw_prob[src]; }
    else { cells[o] = cell_state[b * S + s]; prob[o] = prob_state[b * S + s]; }
  }
  __syncthreads();   // every read of the carried state is done
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const long long o = ((long long)b * T + (T - 1)) * S + s;
    cell_state[b * S + s] = cells[o]; prob_state[b * S + s] = prob[o];
  }
}

int k_frame_power_raw(const float2 *spec, long long rows, int N, float *raw, cudaStream_t st);   // below

__global__ void frame_raw_kernel(const float2 *__restrict__ spec, long long rows, int N, float *__restrict__ raw) {
  const int KP = spec_pitch(N), K = N / 2 + 1;
  const long long row = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float2 *s = spec + row * KP;
  float acc = 0.f;
  for (int k = threadIdx.x & 31; k < K; k += 32) { float2 v = s[k]; acc += v.x * v.x + v.y * v.y; }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) raw[row] = acc;
}
int k_frame_power_raw(const float2 *spec, long long rows, int N, float *raw, cudaStream_t st) {
  if (rows <= 0) return 0;
  frame_raw_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(spec, rows, N, raw);
  MCAG_CHECK_LAUNCH();
  return 0;
}

}  // namespace mcag

using namespace mcag;

template <int NC> static void fill_thread_twiddles(std::vector<float2> &tw) {
  using P = FftPlan<NC>;
  constexpr int N = 2 * NC, TPF = NC / 8;
  int slot = 0, NS = 8;
  for (int p = 1; p < P::NP; ++p) {
    const int R = P::R[p], NB = 8 / R;
    for (int b = 0; b < NB; ++b)
      for (int r = 1; r < R; ++r, ++slot)
        for (int j = 0; j < TPF; ++j) {
          const int jj = j + b * (NC / 8), k = jj & (NS - 1);
          const double a = -2.0 * M_PI * (double)(k * r * (2 * NC / (NS * R))) / (double)N;
          tw[NC + slot * TPF + j] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    NS *= R;
  }
}

// ======================================================================================================================
struct DevBuf {
  void *p = nullptr; size_t bytes = 0;
  int alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) { p = nullptr; mcag_set_cuda_error(e); return MCAG_ERR_NOMEM; }
    bytes = n;
    return MCAG_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct mcag_proc_s {
  mcag_config cfg;
  int B, M, N, hop, K, KP, P, D, S, Cs /* synthesised channels */, Cout /* channels the caller sees */, Tmax, L;
  int rows;
  int srp_form = 0;   // SSL / SL: 1 pair form, 2 channel form on the tensor cores (mcag_config::srp_form resolved at create)
  cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_in[16] = {nullptr}, ev_done[16] = {nullptr};
  long long launches = 0, frames_total = 0;
  int frames_last = 0;
  // input FIFO (ping-pong), per row capacity fifo_cap floats, fill = carried samples (same for every row)
  DevBuf fifo[2]; int fifo_cur = 0; long long fifo_cap = 0; int fill = 0;
  DevBuf stage_in, stage_out;   // device staging for non-f32 input / output conversion
  DevBuf win, tw, spec, chan_pow, chan_raw, power_db, active, gate;
  DevBuf tau_tab;   // delay operand of the tcgen05 tau-grid GCC (gcc_tc.cu), built once from pair_fx
  DevBuf pair_fx, corr, esum, energy, energy_state, raw_idx, raw_prob, cells, prob, cell_state, prob_state;
  DevBuf steer_fx, steer_tab, beams, tail[2], out_dev; int tail_cur = 0;
  DevBuf lags, curves, curve_state, fg, est, track_doa, track_prob;
  DevBuf mic_fx, srp_ws;
  DevBuf H, H2, thr, stats, gains, Q, noise, dec, qtrace, mask_tab;
  int mask_n_h2c = 0, mask_n_hc = 0;
  bool mask_fused = false;   // MASK: analysis + mask + synthesis in one kernel (mask_fused.cu)
  DevBuf band_raw, band_energy, floor_pow, band_cells, mb_raw_cell, mb_raw_prob, mb_lohi;
  int mb_bw = 0, mb_kmin = 0, mb_kmax = 0; bool mb_fused = false;   // band supports (host copy of what mb_band_kernel derives from H)
  void *pin_in = nullptr, *pin_out = nullptr; size_t pin_in_bytes = 0, pin_out_bytes = 0;
  std::vector<double> h_window;
  // per-kernel CUDA-event timing (mcag_profile_*): events are recorded on the handle's own stream around each launch
  struct ProfRec { cudaEvent_t a, b; int id; };
  bool prof_on = false;
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[MCAG_PROF_COUNT] = {0};
  long long prof_n[MCAG_PROF_COUNT] = {0};
};

namespace {
struct ProfScope {
  mcag_proc p; int id; cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t get(mcag_proc p) {
    if (!p->prof_pool.empty()) { cudaEvent_t e = p->prof_pool.back(); p->prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr; cudaEventCreate(&e); return e;
  }
  ProfScope(mcag_proc p_, int id_) : p(p_), id(id_) {
    if (!p->prof_on) return;
    a = get(p); b = get(p);
    cudaEventRecord(a, p->stream);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, p->stream);
    p->prof_pending.push_back({a, b, id});
  }
};
}  // namespace
#define PROF(id) ProfScope prof_scope_##id(p, id)

static int upload(DevBuf &b, const void *src, size_t bytes, cudaStream_t st) {
  OK(b.alloc(bytes));
  CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
  return MCAG_OK;
}
static int upload_fx(DevBuf &b, const double *turns, size_t n, cudaStream_t st) {
  std::vector<uint64_t> fx(n);
  for (size_t i = 0; i < n; ++i) fx[i] = turns_to_fx(turns[i]);
  OK(upload(b, fx.data(), n * sizeof(uint64_t), st));
  CU(cudaStreamSynchronize(st));
  return MCAG_OK;
}

static int init_state(mcag_proc p) {
  cudaStream_t st = p->stream;
  p->fill = 0; p->fifo_cur = 0; p->tail_cur = 0; p->frames_total = 0; p->frames_last = 0;
  if (p->gate.p) {
    std::vector<GateState> gs((size_t)p->B);
    for (auto &g : gs) { g.acc = 0; g.floor = 0; g.samples = 0; g.estimated = p->cfg.noise_preestimated ? 1 : 0; }
    CU(cudaMemcpyAsync(p->gate.p, gs.data(), gs.size() * sizeof(GateState), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
  }
  if (p->energy_state.p) CU(cudaMemsetAsync(p->energy_state.p, 0, p->energy_state.bytes, st));
  if (p->curve_state.p) CU(cudaMemsetAsync(p->curve_state.p, 0, p->curve_state.bytes, st));
  if (p->fg.p) {
    std::vector<FgState> f((size_t)p->B);
    for (auto &v : f) { v.alpha = 0.f; v.dalpha = 0.f; v.silence = 0; v.prob = -1.f; v.doa = 0.0; }   // BinauralLocalisation.cpp:323-326,338-339
    CU(cudaMemcpyAsync(p->fg.p, f.data(), f.size() * sizeof(FgState), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
  }
  if (p->cell_state.p && p->cfg.kind == MCAG_KIND_MULTIBAND) {
    std::vector<int32_t> c((size_t)p->B, -1);              // no cell yet: _currentDOA = 0 rad (MultibandBinarualLocalisation.cpp:78)
    CU(cudaMemcpyAsync(p->cell_state.p, c.data(), c.size() * 4, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
  } else if (p->cell_state.p) {
    std::vector<int32_t> c((size_t)p->B * p->S, p->D);     // row D of the steering table = the initial DOA of 0 rad (BSAL.cpp:51)
    std::vector<float> pr((size_t)p->B * p->S, -1.0f);     // wipp::set(-1.0, _prob) (BSAL.cpp:52)
    CU(cudaMemcpyAsync(p->cell_state.p, c.data(), c.size() * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(p->prob_state.p, pr.data(), pr.size() * 4, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
  }
  for (int i = 0; i < 2; ++i) if (p->tail[i].p) CU(cudaMemsetAsync(p->tail[i].p, 0, p->tail[i].bytes, st));
  if (p->Q.p) { CU(cudaMemsetAsync(p->Q.p, 0, p->Q.bytes, st)); CU(cudaMemsetAsync(p->noise.p, 0, p->noise.bytes, st)); }
  CU(cudaStreamSynchronize(st));
  return MCAG_OK;
}

extern "C" {

const char *mcag_last_error(void) { return g_err.c_str(); }
int mcag_version(void) { return 100; }

void mcag_config_init(mcag_config *c) {
  std::memset(c, 0, sizeof(*c));
  c->frame_size = 512; c->hop = 256; c->n_channels = 2; c->n_streams = 1; c->max_frames_per_call = 256;
  c->n_sources = 1; c->energy_memory = 0.8f; c->corr_memory = 0.8f; c->noise_margin_db = 3.0f; c->floor_seconds = 3.0f;
  c->mask_method = 1; c->n_bands = 45; c->doa_memory = 0.6f;
}

int mcag_create(const mcag_config *cfg, mcag_proc *out) {
  if (!cfg || !out) return mcag_set_error(MCAG_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return mcag_set_error(MCAG_ERR_CUDA, "no CUDA device: mcarray_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return mcag_set_error(MCAG_ERR_INVALID, "bad device ordinal");
  const int N = cfg->frame_size;
  if (N != 256 && N != 512 && N != 1024 && N != 2048) return mcag_set_error(MCAG_ERR_INVALID, "frame_size must be 256, 512, 1024 or 2048");
  if (cfg->hop <= 0 || (N % cfg->hop) || (cfg->hop & 3) || N / cfg->hop > 4) return mcag_set_error(MCAG_ERR_INVALID, "hop must divide N (N/hop <= 4) and be a multiple of 4");
  if (cfg->n_channels < 1 || cfg->n_streams < 1 || cfg->max_frames_per_call < 1) return mcag_set_error(MCAG_ERR_INVALID, "bad channel / stream / frame counts");
  const int kind = cfg->kind;
  if ((kind == MCAG_KIND_MASK || kind == MCAG_KIND_FREQGCC || kind == MCAG_KIND_MULTIBAND) && cfg->n_channels != 2)
    return mcag_set_error(MCAG_ERR_INVALID, "Binaural masking is only working for 2 channels.");   // FastBinauralMasking.cpp:88-91
  CU(cudaSetDevice(cfg->device));

  mcag_proc p = new mcag_proc_s();
  p->cfg = *cfg;
  p->B = cfg->n_streams; p->M = cfg->n_channels; p->N = N; p->hop = cfg->hop; p->K = N / 2 + 1; p->KP = spec_pitch(N);
  p->P = p->M * (p->M - 1) / 2; p->D = cfg->n_dirs; p->S = cfg->n_sources > 0 ? cfg->n_sources : 1; p->Tmax = cfg->max_frames_per_call;
  p->L = 2 * cfg->max_lag + 1; p->rows = p->B * p->M;
  p->Cs = 0; p->Cout = 0;
  if (kind == MCAG_KIND_SSL) { p->Cs = p->S < p->M ? p->S : p->M; p->Cout = p->M; }
  if (kind == MCAG_KIND_MASK) { p->Cs = 2; p->Cout = 2; }
  int rc = MCAG_OK;
  auto fail = [&](int code) { mcag_destroy(p); return code; };
  // CUDA failures after this point must release the handle (streams, events, every buffer allocated so far): never the bare CU()
#define CUF(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(mcag_set_cuda_error(e__)); } while (0)
  if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(mcag_set_error(MCAG_ERR_CUDA, "stream creation failed"));
  if (cudaStreamCreateWithFlags(&p->copy_in, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&p->copy_out, cudaStreamNonBlocking) != cudaSuccess)
    return fail(mcag_set_error(MCAG_ERR_CUDA, "stream creation failed"));
  for (int i = 0; i < 16; ++i)
    if (cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming) != cudaSuccess)
      return fail(mcag_set_error(MCAG_ERR_CUDA, "event creation failed"));
  cudaStream_t st = p->stream;
  const size_t B = p->B, M = p->M, T = p->Tmax, KP = p->KP, D = p->D, P = p->P, S = p->S;

  // window + twiddles
  p->h_window.resize(N);
  for (int n = 0; n < N; ++n) p->h_window[n] = cfg->window ? cfg->window[n] : std::sqrt(0.5 * (1.0 - std::cos(2.0 * M_PI * n / N)));
  { std::vector<float> w(N); for (int n = 0; n < N; ++n) w[n] = (float)p->h_window[n];
    if ((rc = upload(p->win, w.data(), N * sizeof(float), st))) return fail(rc); CUF(cudaStreamSynchronize(st)); }
  if ((rc = p->tw.alloc(sizeof(float2) * fft_table_len(N)))) return fail(rc);
  if ((rc = mcag_k_twiddles(N, p->tw.p, st))) return fail(rc);

  // FIFO: carried samples (< N) + one call's worth of new samples; a call completing Tmax frames consumes Tmax*hop
  p->fifo_cap = ((long long)N + (long long)(T + 1) * p->hop + 3) & ~3LL;
  for (int i = 0; i < 2; ++i) if ((rc = p->fifo[i].alloc(sizeof(float) * p->rows * p->fifo_cap))) return fail(rc);
  // MASK: one fused kernel (spectra stay on chip) unless the caller wants the spectra, the method is NOTHING (a plain analysis /
  // synthesis pass) or the shape is outside the fused kernel; MCAG_MASK_STAGED=1 forces the staged kernels (tests compare the two)
  p->mask_fused = kind == MCAG_KIND_MASK && !(cfg->emit & MCAG_EMIT_SPECTRA) && cfg->mask_method != 5 && k_mask_fused_supported(N, cfg->hop, cfg->n_bands) &&
                  !getenv("MCAG_MASK_STAGED");
  if ((kind != MCAG_KIND_TDOA && !p->mask_fused) || (cfg->emit & MCAG_EMIT_SPECTRA) || (kind == MCAG_KIND_TDOA && !k_stft_tdoa_fits(p->M, N)))
    if ((rc = p->spec.alloc(sizeof(float2) * B * T * M * KP))) return fail(rc);
  if ((rc = p->chan_pow.alloc(sizeof(float) * B * T * M))) return fail(rc);
  if ((rc = p->power_db.alloc(sizeof(float) * B * T))) return fail(rc);
  if ((rc = p->active.alloc(B * T))) return fail(rc);
  if ((rc = p->gate.alloc(sizeof(GateState) * B))) return fail(rc);
  if (cfg->floor_ccs_power) if ((rc = p->chan_raw.alloc(sizeof(float) * B * T * M))) return fail(rc);

  const bool loc = kind == MCAG_KIND_SSL || kind == MCAG_KIND_SL;
  if (loc || kind == MCAG_KIND_FREQGCC) {
    if (D < 3 || !cfg->pair_tau) return fail(mcag_set_error(MCAG_ERR_INVALID, "pair_tau [P][D] with D >= 3 required"));
    std::vector<double> turns(P * D);
    for (size_t i = 0; i < P * D; ++i) turns[i] = cfg->pair_tau[i] / (double)N;   // exp(+j 2 pi k tau / N)
    if ((rc = upload_fx(p->pair_fx, turns.data(), turns.size(), st))) return fail(rc);
    if (loc) {
      // Channel form: valid when tau_ij(d) = tau_j(d) - tau_i(d) for per-microphone delays tau_m (tau_0 = 0,
      // tau_j = tau_0j).  The reference's scalar pair distances (SteeringBeamforming.cpp:67-73) satisfy this for linear arrays with
      // ascending coordinates, up to the float rounding of doaToDelayFarFieldSamples (~2e-6 samples).
      std::vector<double> mt(M * D, 0.0);
      for (size_t j = 1; j < M; ++j) for (size_t d = 0; d < D; ++d) mt[j * D + d] = cfg->pair_tau[(j - 1) * D + d];
      double worst = 0.0;
      size_t pi = 0;
      for (size_t i = 0; i < M; ++i) for (size_t j = i + 1; j < M; ++j, ++pi) for (size_t d = 0; d < D; ++d)
        worst = std::fmax(worst, std::fabs(cfg->pair_tau[pi * D + d] - (mt[j * D + d] - mt[i * D + d])));
      const bool can = worst <= 2e-5 && k_srp_tensor_supported(p->M);
      if (cfg->srp_form == 2 && !can)
        return fail(mcag_set_error(MCAG_ERR_INVALID, "srp_form = channel: needs 16/32/48/64 channels and pair delays consistent with per-microphone delays"));
      if (cfg->srp_form == 2 && (cfg->emit & MCAG_EMIT_CORR))
        return fail(mcag_set_error(MCAG_ERR_INVALID, "srp_form = channel does not produce per-pair correlations (MCAG_EMIT_CORR)"));
      p->srp_form = (cfg->srp_form == 2 || (cfg->srp_form == 0 && can && !(cfg->emit & MCAG_EMIT_CORR))) ? 2 : 1;
      if (p->srp_form == 2) {
        std::vector<double> mturns(D * M);   // device layout [D][M], exp(-j 2 pi k tau_m / N)
        for (size_t m = 0; m < M; ++m) for (size_t d = 0; d < D; ++d) mturns[d * M + m] = -mt[m * D + d] / (double)N;
        if ((rc = upload_fx(p->mic_fx, mturns.data(), mturns.size(), st))) return fail(rc);
        if ((rc = p->srp_ws.alloc(k_srp_tensor_workspace_bytes((long long)B * T, p->M, N, p->D)))) return fail(rc);
      }
    }
    if (p->srp_form != 2) {
      if ((rc = p->corr.alloc(sizeof(float) * B * T * P * D))) return fail(rc);
      // pair form on the tensor cores (gcc_tc.cu: a GEMM over the bins per pair) for grids of at most 64 delays; MCAG_GCC_CUDA_CORES=1
      // keeps the CUDA-core register-tile kernel (tests compare the two)
      if (k_gcc_tau_tc_supported((int)D) && !getenv("MCAG_GCC_CUDA_CORES")) {
        if ((rc = p->tau_tab.alloc(k_gcc_tau_tc_table_bytes((int)P, N)))) return fail(rc);
        if ((rc = k_gcc_tau_tc_build(p->pair_fx.as<uint64_t>(), (int)P, (int)D, N, p->tau_tab.as<float>(), st))) return fail(rc);
        CUF(cudaStreamSynchronize(st));
      }
    }
  }
  if (loc || kind == MCAG_KIND_SRP) {
    if ((rc = p->esum.alloc(sizeof(float) * B * T * D))) return fail(rc);
    if ((rc = p->energy.alloc(sizeof(float) * B * T * D))) return fail(rc);
    if ((rc = p->energy_state.alloc(sizeof(float) * B * D))) return fail(rc);
    if ((rc = p->raw_idx.alloc(4 * B * T * S)) || (rc = p->raw_prob.alloc(4 * B * T * S)) || (rc = p->cells.alloc(4 * B * T * S)) ||
        (rc = p->prob.alloc(4 * B * T * S)) || (rc = p->cell_state.alloc(4 * B * S)) || (rc = p->prob_state.alloc(4 * B * S)))
      return fail(rc);
  }
  if (kind == MCAG_KIND_SSL) {
    if (!cfg->steer_turns) return fail(mcag_set_error(MCAG_ERR_INVALID, "steer_turns [D+1][M] required (row D = initial DOA)"));
    if ((rc = upload_fx(p->steer_fx, cfg->steer_turns, (D + 1) * M, st))) return fail(rc);
    if ((rc = p->steer_tab.alloc(sizeof(float2) * (D + 1) * M * KP))) return fail(rc);
    if ((rc = k_steer_table(p->steer_fx.as<uint64_t>(), (int)((D + 1) * M), N, p->steer_tab.as<float2>(), st))) return fail(rc);
  }
  if (kind == MCAG_KIND_FREQGCC) {
    if ((rc = p->curves.alloc(sizeof(float) * B * T * D)) || (rc = p->curve_state.alloc(sizeof(float) * B * D)) || (rc = p->fg.alloc(sizeof(FgState) * B)) ||
        (rc = p->est.alloc(B * T)) || (rc = p->cells.alloc(4 * B * T)))
      return fail(rc);
    if (cfg->doa_tracker)
      if ((rc = p->track_doa.alloc(sizeof(double) * B * T)) || (rc = p->track_prob.alloc(sizeof(float) * B * T))) return fail(rc);
  }
  if (kind == MCAG_KIND_MULTIBAND) {
    const size_t nb = cfg->n_bands;
    if (D < 2 || !cfg->pair_tau || nb < 1 || !cfg->band_coefs) return fail(mcag_set_error(MCAG_ERR_INVALID, "pair_tau [D] and band_coefs [nb][N/2+1] required"));
    std::vector<double> turns(D);
    for (size_t i = 0; i < D; ++i) turns[i] = cfg->pair_tau[i] / (double)N;   // exp(+j 2 pi k tau / N)
    if ((rc = upload_fx(p->pair_fx, turns.data(), D, st))) return fail(rc);
    if ((rc = p->steer_tab.alloc(sizeof(float2) * D * KP))) return fail(rc);   // phasor table W[d][k], L2-resident (76 KB at D = 37, N = 512)
    if ((rc = k_steer_table(p->pair_fx.as<uint64_t>(), (int)D, N, p->steer_tab.as<float2>(), st))) return fail(rc);
    std::vector<float> H(nb * KP, 0.f);
    for (size_t b = 0; b < nb; ++b) for (int k = 0; k < p->K; ++k) H[b * KP + k] = (float)cfg->band_coefs[b * p->K + k];
    if ((rc = upload(p->H, H.data(), H.size() * 4, st))) return fail(rc);
    // band supports [lo, hi): first / one-past-last non-zero coefficient, as mb_band_kernel derives them
    std::vector<int> lohi(2 * nb);
    p->mb_kmin = p->K; p->mb_kmax = 0; p->mb_bw = 0;
    for (size_t b = 0; b < nb; ++b) {
      int lo = p->K, hi = 0;
      for (int k = 0; k < p->K; ++k) if (H[b * KP + k] != 0.f) { lo = std::min(lo, k); hi = k + 1; }
      if (hi <= lo) { lo = 0; hi = 0; }   // an all-zero band: empty support
      lohi[2 * b] = lo; lohi[2 * b + 1] = hi;
      if (hi > lo) { p->mb_kmin = std::min(p->mb_kmin, lo); p->mb_kmax = std::max(p->mb_kmax, hi); p->mb_bw = std::max(p->mb_bw, hi - lo); }
    }
    for (size_t b = 0; b < nb; ++b) if (lohi[2 * b + 1] == 0) lohi[2 * b] = lohi[2 * b + 1] = std::min(p->mb_kmin, p->K - 1);
    if ((rc = upload(p->mb_lohi, lohi.data(), lohi.size() * sizeof(int), st))) return fail(rc);
    // MCAG_MB_GENERAL=1 keeps the general three-kernel path (tests compare the two)
    p->mb_fused = p->mb_kmax > p->mb_kmin && k_mb_fused_supported((int)D, p->mb_bw, p->mb_kmin, p->mb_kmax, (int)nb) && !getenv("MCAG_MB_GENERAL");
    CUF(cudaStreamSynchronize(st));
    if ((rc = p->band_raw.alloc(p->mb_fused ? 16 : 4 * B * T * nb * D)) || (rc = p->curves.alloc(4 * B * T * nb * D)) || (rc = p->curve_state.alloc(4 * B * nb * D)) ||
        (rc = p->band_energy.alloc(4 * B * T * nb)) || (rc = p->floor_pow.alloc(4 * B * T)) || (rc = p->energy.alloc(4 * B * T * D)) ||
        (rc = p->band_cells.alloc(4 * B * T * nb)) || (rc = p->mb_raw_cell.alloc(4 * B * T)) || (rc = p->mb_raw_prob.alloc(4 * B * T)) ||
        (rc = p->cells.alloc(4 * B * T)) || (rc = p->prob.alloc(4 * B * T)) || (rc = p->cell_state.alloc(4 * B)))
      return fail(rc);
  }
  if (kind == MCAG_KIND_TDOA) {
    if (p->M < 2) return fail(mcag_set_error(MCAG_ERR_INVALID, "TDOA needs at least 2 channels"));
    if ((rc = p->lags.alloc(4 * B * T * P))) return fail(rc);
    if (cfg->emit & MCAG_EMIT_CURVES) if ((rc = p->curves.alloc(sizeof(float) * B * T * P * p->L))) return fail(rc);
  }
  if (kind == MCAG_KIND_DSFAN) {
    if (D < 1 || (!cfg->steer_turns && !cfg->fs_weights)) return fail(mcag_set_error(MCAG_ERR_INVALID, "steer_turns [D][M] or fs_weights [D][M][K] required"));
    if (cfg->fs_weights) {   // filter-and-sum: loaded per-bin weights, [D][M][KP] float2 on the device
      std::vector<float2> w(D * M * KP, make_float2(0.f, 0.f));
      for (size_t dm = 0; dm < D * M; ++dm)
        for (int k = 0; k < p->K; ++k) w[dm * KP + k] = make_float2((float)cfg->fs_weights[2 * (dm * p->K + k)], (float)cfg->fs_weights[2 * (dm * p->K + k) + 1]);
      if ((rc = upload(p->steer_tab, w.data(), w.size() * sizeof(float2), st))) return fail(rc);
      CUF(cudaStreamSynchronize(st));
    } else if ((rc = upload_fx(p->steer_fx, cfg->steer_turns, D * M, st))) return fail(rc);
    if ((rc = p->beams.alloc(sizeof(float2) * B * T * D * fan_out_pitch(N)))) return fail(rc);   // rows padded to 32 bytes: 256-bit stores
  }
  if (kind == MCAG_KIND_SRP) {
    if (D < 3 || !cfg->mic_tau) return fail(mcag_set_error(MCAG_ERR_INVALID, "mic_tau [M][D] with D >= 3 required"));
    std::vector<double> turns(D * M);   // device layout [D][M], exp(-j 2 pi k tau_m / N)
    for (size_t m = 0; m < M; ++m) for (size_t d = 0; d < D; ++d) turns[d * M + m] = -cfg->mic_tau[m * D + d] / (double)N;
    if ((rc = upload_fx(p->mic_fx, turns.data(), turns.size(), st))) return fail(rc);
    if (k_srp_tensor_supported(p->M))   // scratch of the tcgen05 contraction: bin-major whitened spectra (hi / lo) + partial energy maps
      if ((rc = p->srp_ws.alloc(k_srp_tensor_workspace_bytes((long long)B * T, p->M, N, p->D)))) return fail(rc);
  }
  if (p->Cs > 0) {
    if (kind != MCAG_KIND_MASK)   // MASK masks the analysis spectra in place (staged) or never materialises them (fused)
      if ((rc = p->beams.alloc(sizeof(float2) * B * T * p->Cs * KP))) return fail(rc);
    for (int i = 0; i < 2; ++i) if ((rc = p->tail[i].alloc(sizeof(float) * B * p->Cs * (N - p->hop)))) return fail(rc);
    if ((rc = p->out_dev.alloc(sizeof(float) * B * p->Cs * T * p->hop))) return fail(rc);
  }
  if (kind == MCAG_KIND_MASK) {
    const size_t nb = cfg->n_bands;
    if (nb < 1 || !cfg->band_coefs || !cfg->band_thresholds) return fail(mcag_set_error(MCAG_ERR_INVALID, "band_coefs / band_thresholds required"));
    std::vector<float> H(nb * KP, 0.f), H2(nb * KP, 0.f), thr(nb);
    for (size_t b = 0; b < nb; ++b) {
      for (int k = 0; k < p->K; ++k) { double h = cfg->band_coefs[b * p->K + k]; H[b * KP + k] = (float)h; H2[b * KP + k] = (float)(h * h); }
      thr[b] = (float)cfg->band_thresholds[b];
    }
    if ((rc = upload(p->H, H.data(), H.size() * 4, st)) || (rc = upload(p->H2, H2.data(), H2.size() * 4, st)) || (rc = upload(p->thr, thr.data(), nb * 4, st)))
      return fail(rc);
    // compact tables of the fused kernel: band supports [lo, hi) with their squared magnitudes (band-major, what mask_stats_kernel sums)
    // and, per bin, the covering bands [lo, hi) with their magnitudes at that bin (bin-major, what mask_apply_kernel sums)
    if (p->mask_fused) {
      std::vector<int> tab(4 * nb + p->K);
      std::vector<float> h2c, hc;
      for (size_t b = 0; b < nb; ++b) {
        int lo = p->K, hi = 0;
        for (int k = 0; k < p->K; ++k) if (H2[b * KP + k] != 0.f) { lo = std::min(lo, k); hi = k + 1; }
        if (hi <= lo) { lo = 0; hi = 0; }
        tab[4 * b] = lo; tab[4 * b + 1] = hi; tab[4 * b + 2] = (int)h2c.size(); tab[4 * b + 3] = 0;
        for (int k = lo; k < hi; ++k) h2c.push_back(H2[b * KP + k]);
      }
      for (int k = 0; k < p->K; ++k) {
        int lo = (int)nb, hi = 0;
        for (size_t b = 0; b < nb; ++b) if (H[b * KP + k] != 0.f) { lo = std::min(lo, (int)b); hi = (int)b + 1; }
        if (hi <= lo) { lo = 0; hi = 0; }
        tab[4 * nb + k] = lo | (hi << 8) | ((int)hc.size() << 16);
        for (int b = lo; b < hi; ++b) hc.push_back(H[(size_t)b * KP + k]);
      }
      if (nb > 255 || h2c.size() > 8192 || hc.size() > 8192) p->mask_fused = false;   // a dense bank: the staged kernels
      else {
        p->mask_n_h2c = (int)h2c.size(); p->mask_n_hc = (int)hc.size();
        const size_t n0 = tab.size();
        tab.resize(n0 + h2c.size() + hc.size());
        std::memcpy(tab.data() + n0, h2c.data(), h2c.size() * 4);
        std::memcpy(tab.data() + n0 + h2c.size(), hc.data(), hc.size() * 4);
        if ((rc = upload(p->mask_tab, tab.data(), tab.size() * 4, st))) return fail(rc);
      }
    }
    CUF(cudaStreamSynchronize(st));
    if ((rc = p->Q.alloc(4 * B * nb)) || (rc = p->noise.alloc(4 * B * nb))) return fail(rc);
    if (!p->mask_fused)
      if ((rc = p->stats.alloc(4 * B * T * nb * 6)) || (rc = p->gains.alloc(4 * B * T * nb * 2))) return fail(rc);
    if (!p->mask_fused || (cfg->emit & MCAG_EMIT_MASK_TRACE))
      if ((rc = p->dec.alloc(B * T * nb)) || (rc = p->qtrace.alloc(4 * B * T * nb))) return fail(rc);
  }
  if ((rc = init_state(p))) return fail(rc);
  *out = p;
  return MCAG_OK;
#undef CUF
}

void mcag_destroy(mcag_proc p) {
  if (!p) return;
  cudaSetDevice(p->cfg.device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->copy_in) cudaStreamSynchronize(p->copy_in);
  if (p->copy_out) cudaStreamSynchronize(p->copy_out);
  DevBuf *all[] = {&p->fifo[0], &p->fifo[1], &p->stage_in, &p->stage_out, &p->win, &p->tw, &p->spec, &p->chan_pow, &p->chan_raw, &p->power_db,
                   &p->active, &p->gate, &p->tau_tab, &p->pair_fx, &p->corr, &p->esum, &p->energy, &p->energy_state, &p->raw_idx, &p->raw_prob, &p->cells,
                   &p->prob, &p->cell_state, &p->prob_state, &p->steer_fx, &p->steer_tab, &p->beams, &p->tail[0], &p->tail[1], &p->out_dev,
                   &p->lags, &p->curves, &p->curve_state, &p->fg, &p->est, &p->track_doa, &p->track_prob, &p->mic_fx, &p->srp_ws, &p->H, &p->H2, &p->thr, &p->stats, &p->gains, &p->Q,
                   &p->noise, &p->dec, &p->qtrace, &p->mask_tab, &p->band_raw, &p->band_energy, &p->floor_pow, &p->band_cells, &p->mb_raw_cell, &p->mb_raw_prob, &p->mb_lohi};
  for (DevBuf *b : all) b->release();
  for (auto &r : p->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (cudaEvent_t e : p->prof_pool) cudaEventDestroy(e);
  if (p->pin_in) cudaFreeHost(p->pin_in);
  if (p->pin_out) cudaFreeHost(p->pin_out);
  for (int i = 0; i < 16; ++i) { if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]); if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]); }
  if (p->copy_in) cudaStreamDestroy(p->copy_in);
  if (p->copy_out) cudaStreamDestroy(p->copy_out);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

int mcag_reset(mcag_proc p) {
  if (!p) return mcag_set_error(MCAG_ERR_INVALID, "null handle");
  CU(cudaSetDevice(p->cfg.device));
  return init_state(p);
}

int mcag_flush_input(mcag_proc p) {
  if (!p) return mcag_set_error(MCAG_ERR_INVALID, "null handle");
  p->fill = 0;
  return MCAG_OK;
}

int mcag_get_info(mcag_proc p, mcag_info *i) {
  if (!p || !i) return mcag_set_error(MCAG_ERR_INVALID, "null argument");
  i->frame_size = p->hop; i->window_size = p->N; i->hop = p->hop; i->analysis_length = p->N + 2; i->one_sided_length = p->K;
  i->n_channels = p->M; i->n_streams = p->B; i->max_latency = p->N; i->n_dirs = p->D; i->n_pairs = p->P; i->n_sources = p->S;
  i->n_out_channels = p->Cout; i->spectrum_pitch = p->KP; i->max_frames_per_call = p->Tmax; i->srp_form = p->srp_form;
  i->beams_pitch = p->cfg.kind == MCAG_KIND_DSFAN ? fan_out_pitch(p->cfg.frame_size) : p->KP;
  return MCAG_OK;
}

int mcag_synchronize(mcag_proc p) {
  if (!p) return mcag_set_error(MCAG_ERR_INVALID, "null handle");
  CU(cudaStreamSynchronize(p->stream));
  return MCAG_OK;
}
int mcag_frames_done(mcag_proc p) { return p ? p->frames_last : 0; }
long long mcag_frames_total(mcag_proc p) { return p ? p->frames_total : 0; }
void *mcag_stream(mcag_proc p) { return p ? (void *)p->stream : nullptr; }
long long mcag_kernel_launches(mcag_proc p) { return p ? p->launches : 0; }

static const char *const k_prof_names[MCAG_PROF_COUNT] = {"stft", "gate", "gcc_tau", "energy", "select_doa", "ds_select", "istft", "curve_scan",
                                                        "stft_gcc", "ds_fan", "srp", "mask_stats", "mask_scan", "mask_apply", "mask_fused"};
const char *mcag_profile_name(int id) { return (id >= 0 && id < MCAG_PROF_COUNT) ? k_prof_names[id] : ""; }
int mcag_profile_enable(mcag_proc p, int on) {
  if (!p) return mcag_set_error(MCAG_ERR_INVALID, "null handle");
  p->prof_on = on != 0;
  return MCAG_OK;
}
int mcag_profile_read(mcag_proc p, double *ms, long long *count, int reset) {
  if (!p) return mcag_set_error(MCAG_ERR_INVALID, "null handle");
  CU(cudaSetDevice(p->cfg.device));
  CU(cudaStreamSynchronize(p->stream));
  for (auto &r : p->prof_pending) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { p->prof_ms[r.id] += t; p->prof_n[r.id]++; }
    p->prof_pool.push_back(r.a); p->prof_pool.push_back(r.b);
  }
  p->prof_pending.clear();
  for (int i = 0; i < MCAG_PROF_COUNT; ++i) {
    if (ms) ms[i] = p->prof_ms[i];
    if (count) count[i] = p->prof_n[i];
    if (reset) { p->prof_ms[i] = 0; p->prof_n[i] = 0; }
  }
  return MCAG_OK;
}

void *mcag_host_alloc(long long bytes) {
  void *ptr = nullptr;
  if (cudaMallocHost(&ptr, (size_t)bytes) != cudaSuccess) { mcag_set_error(MCAG_ERR_NOMEM, "cudaMallocHost failed"); return nullptr; }
  return ptr;
}
void mcag_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

void *mcag_dev_alloc(long long bytes) {
  void *ptr = nullptr;
  if (bytes <= 0) bytes = 16;
  if (cudaMalloc(&ptr, (size_t)bytes) != cudaSuccess) { mcag_set_error(MCAG_ERR_NOMEM, "cudaMalloc failed (mcarray_b200 has no CPU fallback)"); return nullptr; }
  cudaMemset(ptr, 0, (size_t)bytes);
  return ptr;
}
void mcag_dev_free(void *d_ptr) { if (d_ptr) cudaFree(d_ptr); }
int mcag_dev_upload(void *d_dst, const void *h_src, long long bytes) {
  CU(cudaMemcpy(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice));
  return MCAG_OK;
}
int mcag_dev_download(void *h_dst, const void *d_src, long long bytes) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
  return MCAG_OK;
}

}  // extern "C"

// ----------------------------------------------------------------------------------------------------------------------
// the frame pipeline: T complete frames are available at `x` (device, rows x pitch); results land in the handle's arrays,
// synthesised audio (if any) in out_dev [B*Cs][T*hop]
// ----------------------------------------------------------------------------------------------------------------------
// Streams [b0, b0 + nb) of the handle; every result / state array is stream-major, so a sub-batch is a pointer offset.
// process_host uses this to overlap the host->device copy of one group of streams with the kernels of the previous one.
// ext_out (optional): the caller's device buffer [B*ext_rows][ext_pitch]; the synthesis kernel writes its Cs channels of every stream
// straight into it instead of out_dev.
static int run_frames(mcag_proc p, const float *x_all, long long pitch, int T, int b0, int nb, float *ext_out = nullptr, long long ext_pitch = 0,
                      int ext_rows = 0) {
  cudaStream_t st = p->stream;
  const int B = nb, M = p->M, N = p->N, hop = p->hop, D = p->D, P = p->P, S = p->S, kind = p->cfg.kind, KP = p->KP, Cs = p->Cs;
  const long long BT = (long long)B * T, o = b0;   // o: stream offset
  const float *x = x_all + o * M * pitch;
  float2 *spec = p->spec.p ? p->spec.as<float2>() + o * T * M * KP : nullptr;
  float *chan_pow = p->chan_pow.as<float>() + o * T * M;
  float *chan_raw = p->chan_raw.p ? p->chan_raw.as<float>() + o * T * M : nullptr;
  unsigned char *active = p->active.as<unsigned char>() + o * T;
  const float *win = p->win.as<float>();
  const float2 *tw = p->tw.as<float2>();
  if (kind == MCAG_KIND_MASK && p->mask_fused) {
    // analysis + FastBinauralMasking + synthesis in one kernel: spectra, band statistics and gains never leave the SM
    {
      PROF(MCAG_PROF_MASK_FUSED);
      const int nb_ = p->cfg.n_bands;
      const long long ov = N - hop;
      float *outp = ext_out ? ext_out + o * ext_rows * ext_pitch : p->out_dev.as<float>() + o * 2 * T * hop;
      OK(k_mask_fused(x, pitch, B, T, N, hop, win, tw, p->mask_tab.as<int>(), p->mask_n_h2c, p->mask_n_hc, nb_,
                      p->cfg.mask_method, p->cfg.mask_alg, p->thr.as<float>(), p->Q.as<float>() + o * nb_, p->noise.as<float>() + o * nb_,
                      (int)(p->frames_total > 2 ? 2 : p->frames_total), p->tail[p->tail_cur].as<float>() + o * 2 * ov,
                      p->tail[p->tail_cur ^ 1].as<float>() + o * 2 * ov, outp, ext_out ? ext_pitch : (long long)T * hop, ext_out ? ext_rows : 2, chan_pow,
                      p->dec.p ? p->dec.as<unsigned char>() + o * T * nb_ : nullptr, p->qtrace.p ? p->qtrace.as<float>() + o * T * nb_ : nullptr, st));
      p->launches++;
    }
    PROF(MCAG_PROF_GATE);
    const int needed = (int)(p->cfg.floor_seconds * (float)p->cfg.sample_rate);
    gate_kernel<<<B, 256, 0, st>>>(chan_pow, chan_raw, B, T, M, N, p->cfg.use_power_floor, p->cfg.floor_ccs_power, p->cfg.noise_margin_db, needed,
                                   p->gate.as<GateState>() + o, p->power_db.as<float>() + o * T, active, nullptr);
    MCAG_CHECK_LAUNCH();
    p->launches++;
    return MCAG_OK;
  }
  if (kind == MCAG_KIND_TDOA) {
    // fused STFT -> GCC-PHAT -> lag argmax: the spectra stay in shared memory unless MCAG_EMIT_SPECTRA asks for them
    PROF(MCAG_PROF_TDOA);
    OK(k_stft_tdoa(x, pitch, B, T, M, N, hop, p->cfg.max_lag, win, tw, spec, chan_pow,
                   p->curves.p ? p->curves.as<float>() + o * T * P * p->L : nullptr, p->lags.as<int32_t>() + o * T * P, st));
    p->launches++;
  } else {
    PROF(MCAG_PROF_STFT);
    OK(k_stft(x, pitch, B * M, M, T, N, hop, win, tw, spec, chan_pow, chan_raw, st));   // chan_raw: only with floor_ccs_power (FreqGCC)
    p->launches += (N == 256) ? (chan_raw ? 3 : 2) : 1;
  }
  if (kind != MCAG_KIND_MULTIBAND) {
    PROF(MCAG_PROF_GATE);
    const int needed = (int)(p->cfg.floor_seconds * (float)p->cfg.sample_rate);
    gate_kernel<<<B, 256, 0, st>>>(chan_pow, chan_raw, B, T, M, N, p->cfg.use_power_floor, p->cfg.floor_ccs_power, p->cfg.noise_margin_db, needed,
                                   p->gate.as<GateState>() + o, p->power_db.as<float>() + o * T, active,
                                   p->est.p ? p->est.as<unsigned char>() + o * T : nullptr);
    MCAG_CHECK_LAUNCH();
    p->launches++;
  }
  float *energy = p->energy.p ? p->energy.as<float>() + o * T * D : nullptr;
  float *esum = p->esum.p ? p->esum.as<float>() + o * T * D : nullptr;
  float *energy_state = p->energy_state.p ? p->energy_state.as<float>() + o * D : nullptr;
  auto select_and_carry = [&]() -> int {
    PROF(MCAG_PROF_SELECT_DOA);
    int32_t *raw_idx = p->raw_idx.as<int32_t>() + o * T * S;
    float *raw_prob = p->raw_prob.as<float>() + o * T * S;
    OK(k_select_doa(energy, BT, D, P, S, raw_idx, raw_prob, st));
    carry_cells_kernel<<<B, 128, 0, st>>>(raw_idx, raw_prob, active, B, T, S, p->cell_state.as<int32_t>() + o * S,
                                                            p->prob_state.as<float>() + o * S, p->cells.as<int32_t>() + o * T * S,
                                                            p->prob.as<float>() + o * T * S);
    MCAG_CHECK_LAUNCH();
    p->launches += 2;
    return MCAG_OK;
  };
  auto synth = [&](const float2 *src, int C) -> int {
    PROF(MCAG_PROF_ISTFT);
    const long long ov = N - hop;
    if (ext_out)
      OK(k_istft(src, B, T, C, C, N, hop, win, tw, p->tail[p->tail_cur].as<float>() + o * C * ov, p->tail[p->tail_cur ^ 1].as<float>() + o * C * ov,
                 ext_out + o * ext_rows * ext_pitch, ext_pitch, ext_rows, st));
    else
      OK(k_istft(src, B, T, C, C, N, hop, win, tw, p->tail[p->tail_cur].as<float>() + o * C * ov, p->tail[p->tail_cur ^ 1].as<float>() + o * C * ov,
                 p->out_dev.as<float>() + o * C * T * hop, (long long)T * hop, C, st));
    p->launches++;
    return MCAG_OK;
  };

  if (kind == MCAG_KIND_SSL || kind == MCAG_KIND_SL) {
    const float a = p->cfg.energy_memory, b = 1.0f - p->cfg.energy_memory;   // float arithmetic as SteeringBeamforming.cpp:134,139
    if (p->srp_form == 2) {
      // channel form: the whole pair sum of computeCorrelations is one tcgen05 contraction per bin (srp_tc.cu)
      {
        PROF(MCAG_PROF_SRP);
        OK(k_srp_tensor_ws(spec, B, T, M, N, p->mic_fx.as<uint64_t>(), D, esum, p->srp_ws.p, p->srp_ws.bytes, st));
        p->launches += 3;
      }
      PROF(MCAG_PROF_ENERGY);
      OK(k_pair_sum(esum, BT, 1, D, b, energy, st));
      OK(k_energy_scan(energy, B, T, D, a, active, energy_state, energy, st));
      p->launches += 2;
    } else {
      float *corr = p->corr.as<float>() + o * T * P * D;
      {
        PROF(MCAG_PROF_GCC_TAU);
        if (p->tau_tab.p) { OK(k_gcc_tau_tc(spec, B, T, M, N, p->tau_tab.as<float>(), D, corr, st)); p->launches++; }
        else { OK(k_gcc_tau(spec, B, T, M, N, p->pair_fx.as<uint64_t>(), D, corr, st)); p->launches += (D <= 40) ? 1 : (D + 63) / 64; }
      }
      PROF(MCAG_PROF_ENERGY);
      OK(k_pair_sum(corr, BT, P, D, b, esum, st));
      OK(k_energy_scan(esum, B, T, D, a, active, energy_state, energy, st));
      p->launches += 2;
    }
    OK(select_and_carry());
    if (kind == MCAG_KIND_SSL) {
      float2 *beams = p->beams.as<float2>() + o * T * Cs * KP;
      {
        PROF(MCAG_PROF_DS_SELECT);
        OK(k_ds_select(spec, B, T, M, N, p->steer_tab.as<float2>(), p->cells.as<int32_t>() + o * T * S, S, Cs, beams, st));
        p->launches++;
      }
      OK(synth(beams, Cs));
    }
  } else if (kind == MCAG_KIND_FREQGCC) {
    float *corr = p->corr.as<float>() + o * T * D;
    {
      PROF(MCAG_PROF_GCC_TAU);
      if (p->tau_tab.p) { OK(k_gcc_tau_tc(spec, B, T, 2, N, p->tau_tab.as<float>(), D, corr, st)); p->launches++; }
      else { OK(k_gcc_tau(spec, B, T, 2, N, p->pair_fx.as<uint64_t>(), D, corr, st)); p->launches += (D <= 40) ? 1 : (D + 63) / 64; }
    }
    {
      PROF(MCAG_PROF_CURVE_SCAN);
      // windowsToDecay = secondsToDecay * fs / (analysisLength/2 - 1) with secondsToDecay = 3 (BinauralLocalisation.cpp:532-533)
      const int wtd = 3 * p->cfg.sample_rate / (N / 2);
      const float step = (float)(3 * M_PI / 180);   // _doaStep (:328); the tracker only exists on the reference's 61-cell grid
      OK(k_curve_scan_argmax(corr, B, T, D, p->cfg.corr_memory, p->cfg.doa_memory, wtd, step, p->track_doa.p != nullptr, active,
                             p->est.as<unsigned char>() + o * T, p->curve_state.as<float>() + o * D, p->fg.as<FgState>() + o,
                             p->curves.as<float>() + o * T * D, p->cells.as<int32_t>() + o * T,
                             p->track_doa.p ? p->track_doa.as<double>() + o * T : nullptr, p->track_prob.p ? p->track_prob.as<float>() + o * T : nullptr, st));
      p->launches += 3;
    }
  } else if (kind == MCAG_KIND_MULTIBAND) {
    const int nb_ = p->cfg.n_bands;
    float *band_raw = p->band_raw.as<float>() + o * T * nb_ * D, *curves = p->curves.as<float>() + o * T * nb_ * D;
    float *band_energy = p->band_energy.as<float>() + o * T * nb_, *floor_pow = p->floor_pow.as<float>() + o * T;
    int32_t *raw_cell = p->mb_raw_cell.as<int32_t>() + o * T;
    float *raw_prob = p->mb_raw_prob.as<float>() + o * T;
    if (p->mb_fused) {
      PROF(MCAG_PROF_GCC_TAU);
      OK(k_mb_fused(spec, B, T, N, p->H.as<float>(), p->mb_lohi.as<int>(), nb_, p->mb_bw, p->mb_kmin, p->mb_kmax, p->steer_tab.as<float2>(), D,
                    p->cfg.corr_memory, p->curve_state.as<float>() + o * nb_ * D, curves, band_energy, floor_pow, p->energy.as<float>() + o * T * D,
                    p->band_cells.as<int32_t>() + o * T * nb_, raw_cell, raw_prob, st));
      p->launches++;
    } else {
      {
        PROF(MCAG_PROF_GCC_TAU);
        OK(k_mb_band(spec, BT, N, p->H.as<float>(), nb_, p->steer_tab.as<float2>(), D, band_raw, band_energy, floor_pow, st));
        p->launches++;
      }
      {
        PROF(MCAG_PROF_CURVE_SCAN);
        OK(k_mb_scan(band_raw, B, T, nb_, D, p->cfg.corr_memory, p->curve_state.as<float>() + o * nb_ * D, curves, st));
        p->launches++;
      }
      PROF(MCAG_PROF_SELECT_DOA);
      OK(k_mb_summary(curves, band_energy, BT, nb_, D, p->energy.as<float>() + o * T * D, p->band_cells.as<int32_t>() + o * T * nb_, raw_cell, raw_prob, st));
      p->launches++;
    }
    {
      PROF(MCAG_PROF_SELECT_DOA);
      const int needed = (int)(p->cfg.floor_seconds * (float)p->cfg.sample_rate);
      OK(k_mb_gate(floor_pow, chan_pow, raw_cell, raw_prob, B, T, N, p->cfg.use_power_floor, p->cfg.noise_margin_db, needed,
                   p->gate.as<GateState>() + o, p->cell_state.as<int32_t>() + o, p->power_db.as<float>() + o * T, active,
                   p->cells.as<int32_t>() + o * T, p->prob.as<float>() + o * T, st));
      p->launches++;
    }
  } else if (kind == MCAG_KIND_DSFAN) {
    PROF(MCAG_PROF_DS_FAN);
    // 16 / 32 / 48 / 64 microphones: the contraction runs on the tensor cores with four consecutive bins of a tile resident in TMEM
    // (fan_tc.cu); other counts take the CUDA-core register-tile kernel.  MCAG_FAN_CUDA_CORES=1 forces the latter (tests compare the two).
    if (p->cfg.fs_weights)
      OK(k_fs_fan(spec, B, T, M, N, p->steer_tab.as<float2>(), D, p->beams.as<float2>() + o * T * D * fan_out_pitch(N), st, fan_out_pitch(N)));
    else if (k_ds_fan_tensor_supported(M) && !getenv("MCAG_FAN_CUDA_CORES"))
      OK(k_ds_fan_tensor(spec, B, T, M, N, p->steer_fx.as<uint64_t>(), D, p->beams.as<float2>() + o * T * D * fan_out_pitch(N), st, fan_out_pitch(N)));
    else
      OK(k_ds_fan(spec, B, T, M, N, p->steer_fx.as<uint64_t>(), D, p->beams.as<float2>() + o * T * D * fan_out_pitch(N), st, fan_out_pitch(N)));
    p->launches++;
  } else if (kind == MCAG_KIND_SRP) {
    {
      PROF(MCAG_PROF_SRP);
      OK(k_srp_tensor_ws(spec, B, T, M, N, p->mic_fx.as<uint64_t>(), D, esum, p->srp_ws.p, p->srp_ws.bytes, st));
      p->launches += 3;   // prepare (whiten / split / transpose), tcgen05 contraction, partial-map reduction
    }
    {
      PROF(MCAG_PROF_ENERGY);
      // the pair sum enters the smoothing scaled by (1 - a), like every pair correlation (SteeringBeamforming.cpp:139)
      OK(k_pair_sum(esum, BT, 1, D, 1.0f - p->cfg.energy_memory, energy, st));
      OK(k_energy_scan(energy, B, T, D, p->cfg.energy_memory, active, energy_state, energy, st));
      p->launches += 2;
    }
    OK(select_and_carry());
  } else if (kind == MCAG_KIND_MASK) {
    const int nb_ = p->cfg.n_bands;
    if (p->cfg.mask_method != 5) {   // NOTHING: pass-through (FastBinauralMasking.cpp:130-134)
      float *stats = p->stats.as<float>() + o * T * nb_ * 6, *gains = p->gains.as<float>() + o * T * nb_ * 2;
      {
        PROF(MCAG_PROF_MASK_STATS);
        OK(k_mask_stats(spec, BT, N, p->H2.as<float>(), nb_, stats, st));
        p->launches++;
      }
      {
        PROF(MCAG_PROF_MASK_SCAN);
        OK(k_mask_scan(stats, B, T, N, nb_, p->cfg.mask_method, p->cfg.mask_alg, p->thr.as<float>(), p->Q.as<float>() + o * nb_,
                       p->noise.as<float>() + o * nb_, (int)(p->frames_total > 2 ? 2 : p->frames_total), gains,
                       p->dec.as<unsigned char>() + o * T * nb_, p->qtrace.as<float>() + o * T * nb_, st));
        p->launches++;
      }
      {
        PROF(MCAG_PROF_MASK_APPLY);
        OK(k_mask_apply(spec, BT, N, p->H.as<float>(), nb_, gains, st));
        p->launches++;
      }
    }
    OK(synth(spec, 2));
  }
  return MCAG_OK;
}

// frames completed by appending nsamples to a FIFO that already holds `fill`
static int frames_for(const mcag_proc p, int nsamples) {
  const long long have = (long long)p->fill + nsamples;
  return have < p->N ? 0 : (int)((have - p->N) / p->hop + 1);
}

// core: new samples are on the device at d_new (rows x pitch).  Handles the FIFO, runs the frames, leaves audio in out_dev.
static int process_device_core(mcag_proc p, const float *d_new, long long pitch, int nsamples, int *T_out, float *d_out = nullptr,
                               long long out_pitch = 0) {
  cudaStream_t st = p->stream;
  const int T = frames_for(p, nsamples);
  if (T > p->Tmax) return mcag_set_error(MCAG_ERR_CAPACITY, "process: more frames than max_frames_per_call");
  if ((long long)p->fill + nsamples > p->fifo_cap) return mcag_set_error(MCAG_ERR_CAPACITY, "process: chunk larger than the input FIFO");
  const float *x; long long xp;
  float *cur = p->fifo[p->fifo_cur].as<float>(), *nxt = p->fifo[p->fifo_cur ^ 1].as<float>();
  if (p->fill == 0 && (pitch % 4) == 0 && ((uintptr_t)d_new % 16) == 0) {
    x = d_new; xp = pitch;                       // run straight from the caller's buffer
  } else {
    CU(cudaMemcpy2DAsync(cur + p->fill, p->fifo_cap * 4, d_new, pitch * 4, (size_t)nsamples * 4, p->rows, cudaMemcpyDeviceToDevice, st));
    x = cur; xp = p->fifo_cap;
  }
  if (T > 0) {
    float *ext = nullptr;
    if (p->Cs > 0 && d_out) {
      if ((long long)T * p->hop > out_pitch) return mcag_set_error(MCAG_ERR_CAPACITY, "process: output pitch too small");
      // channels >= Cs are zeros (BeamformingSeparationAndLocalisation.cpp:117-118); the synthesis kernel then writes the first Cs
      if (p->Cs != p->Cout) CU(cudaMemset2DAsync(d_out, out_pitch * 4, 0, (size_t)T * p->hop * 4, (size_t)p->B * p->Cout, st));
      ext = d_out;
    }
    OK(run_frames(p, x, xp, T, 0, p->B, ext, out_pitch, p->Cout));
    if (p->Cs > 0) p->tail_cur ^= 1;
  }
  // carry the unconsumed samples
  const long long have = (long long)p->fill + nsamples, consumed = (long long)T * p->hop, left = have - consumed;
  if (left > 0) {
    if (x == d_new) {
      CU(cudaMemcpy2DAsync(cur, p->fifo_cap * 4, d_new + consumed, pitch * 4, (size_t)left * 4, p->rows, cudaMemcpyDeviceToDevice, st));
    } else if (consumed > 0) {
      CU(cudaMemcpy2DAsync(nxt, p->fifo_cap * 4, cur + consumed, p->fifo_cap * 4, (size_t)left * 4, p->rows, cudaMemcpyDeviceToDevice, st));
      p->fifo_cur ^= 1;
    }
  }
  p->fill = (int)left;
  p->frames_last = T;
  p->frames_total += T;
  *T_out = T;
  return MCAG_OK;
}

template <class Tin> struct Conv;
template <> struct Conv<float> { static constexpr int id = 0; };
template <> struct Conv<double> { static constexpr int id = 1; };
template <> struct Conv<int16_t> { static constexpr int id = 2; };

static int ensure_pinned(void **ptr, size_t *have, size_t need) {
  if (*have >= need) return MCAG_OK;
  if (*ptr) cudaFreeHost(*ptr);
  *ptr = nullptr; *have = 0;
  CU(cudaMallocHost(ptr, need));
  *have = need;
  return MCAG_OK;
}

// out[r][off + i] = (float)in[r][i]: f64 / s16 input rows land in the FIFO in one pass
template <class Tin>
__global__ void convert_rows_kernel(const Tin *__restrict__ in, long long in_pitch, float *__restrict__ out, long long out_pitch, int n, int rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (i < n && r < rows) out[(long long)r * out_pitch + i] = (float)in[(long long)r * in_pitch + i];
}

// synthesised audio float -> the caller's sample type on the device: double widened, int16 rounded to nearest even and saturated
// (the reference's int16 process() overload narrows the same way through DSPONE)
template <class Tout> __device__ __forceinline__ Tout narrow_sample(float v);
template <> __device__ __forceinline__ double narrow_sample<double>(float v) { return (double)v; }
template <> __device__ __forceinline__ int16_t narrow_sample<int16_t>(float v) {
  v = rintf(v);
  return (int16_t)(v > 32767.f ? 32767.f : (v < -32768.f ? -32768.f : v));
}
template <class Tout>
__global__ void convert_out_kernel(const float *__restrict__ in, Tout *__restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = narrow_sample<Tout>(in[i]);
}

#ifndef MCAG_MAX_CHUNKS
#define MCAG_MAX_CHUNKS 8
#endif
constexpr int kMaxChunks = MCAG_MAX_CHUNKS;

// Host-buffer process call.  The streams of the handle are cut into up to kMaxChunks groups; group c+1 is copied host->device
// on the copy-in stream while the kernels of group c run on the compute stream, and the audio of group c goes back on the
// copy-out stream while group c+1 computes (PCIe is the bottleneck of the end-to-end path: 4 B per channel-sample in).
template <class Tio>
static int process_host(mcag_proc p, const Tio *const *in, const Tio *in_packed, long long in_pitch, int nsamples, Tio *const *out, Tio *out_packed,
                        long long out_pitch, int out_capacity, int *nsamples_out) {
  if (!p) return mcag_set_error(MCAG_ERR_INVALID, "null handle");
  if (nsamples < 0 || (!in && !in_packed && nsamples > 0)) return mcag_set_error(MCAG_ERR_INVALID, "bad input");
  CU(cudaSetDevice(p->cfg.device));
  cudaStream_t st = p->stream, sin = p->copy_in, sout = p->copy_out;
  const int rows = p->rows, B = p->B, M = p->M, Cs = p->Cs, Cout = p->Cout;
  const int T = frames_for(p, nsamples);
  if (T > p->Tmax) return mcag_set_error(MCAG_ERR_CAPACITY, "process: more frames than max_frames_per_call");
  if ((long long)p->fill + nsamples > p->fifo_cap) return mcag_set_error(MCAG_ERR_CAPACITY, "process: chunk larger than the input FIFO");
  const bool want_audio = Cs > 0 && (out || out_packed);
  if (want_audio && (long long)T * p->hop > out_capacity) return mcag_set_error(MCAG_ERR_CAPACITY, "process: output buffer too small");
  constexpr bool is_f32 = Conv<Tio>::id == 0;
  const int nout = T * p->hop;

  const size_t in_bytes = (size_t)rows * nsamples * sizeof(Tio);
  int nch = 1;
  if (B >= 2 && in_bytes >= (size_t)8 << 20) nch = B < kMaxChunks ? B : kMaxChunks;

  float *cur = p->fifo[p->fifo_cur].as<float>();
  if (!is_f32 && nsamples > 0 && p->stage_in.bytes < in_bytes) OK(p->stage_in.alloc(in_bytes));
  if (want_audio && !is_f32 && nout > 0 && p->stage_out.bytes < (size_t)B * Cs * nout * sizeof(Tio)) OK(p->stage_out.alloc((size_t)B * Cs * nout * sizeof(Tio)));

  // The FIFO may still be read or carried by work a previous mcag_process_device_f32 call left in flight on the compute stream:
  // the copy-in stream starts behind it.
  CU(cudaEventRecord(p->ev_done[0], st));
  CU(cudaStreamWaitEvent(sin, p->ev_done[0], 0));
  // ---- host -> device, one group of streams after the other on the copy-in stream
  for (int c = 0; c < nch && nsamples > 0; ++c) {
    const int r0 = (int)((long long)B * c / nch) * M, r1 = (int)((long long)B * (c + 1) / nch) * M;
    if (is_f32) {
      float *dst = cur + (long long)r0 * p->fifo_cap + p->fill;
      if (in_packed) CU(cudaMemcpy2DAsync(dst, p->fifo_cap * 4, in_packed + (long long)r0 * in_pitch, in_pitch * 4, (size_t)nsamples * 4, r1 - r0, cudaMemcpyHostToDevice, sin));
      else for (int r = r0; r < r1; ++r) CU(cudaMemcpyAsync(cur + (long long)r * p->fifo_cap + p->fill, in[r], (size_t)nsamples * 4, cudaMemcpyHostToDevice, sin));
    } else {
      Tio *dst = p->stage_in.as<Tio>() + (long long)r0 * nsamples;
      if (in_packed && in_pitch == nsamples)   // dense on both sides: one linear copy (the 2-D form of the same bytes ran below the pinned-copy rate)
        CU(cudaMemcpyAsync(dst, in_packed + (long long)r0 * in_pitch, (size_t)(r1 - r0) * nsamples * sizeof(Tio), cudaMemcpyHostToDevice, sin));
      else if (in_packed) CU(cudaMemcpy2DAsync(dst, (size_t)nsamples * sizeof(Tio), in_packed + (long long)r0 * in_pitch, in_pitch * sizeof(Tio), (size_t)nsamples * sizeof(Tio), r1 - r0, cudaMemcpyHostToDevice, sin));
      else for (int r = r0; r < r1; ++r) CU(cudaMemcpyAsync(p->stage_in.as<Tio>() + (long long)r * nsamples, in[r], (size_t)nsamples * sizeof(Tio), cudaMemcpyHostToDevice, sin));
    }
    CU(cudaEventRecord(p->ev_in[c], sin));
  }
  // ---- kernels per group on the compute stream; audio of the group back on the copy-out stream
  for (int c = 0; c < nch; ++c) {
    const int b0 = (int)((long long)B * c / nch), b1 = (int)((long long)B * (c + 1) / nch);
    if (nsamples > 0) {
      CU(cudaStreamWaitEvent(st, p->ev_in[c], 0));
      if (!is_f32) {
        const int r0 = b0 * M, nr = (b1 - b0) * M;
        dim3 grid((unsigned)((nsamples + 255) / 256), (unsigned)nr);
        convert_rows_kernel<Tio><<<grid, 256, 0, st>>>(p->stage_in.as<Tio>() + (long long)r0 * nsamples, nsamples, cur + (long long)r0 * p->fifo_cap + p->fill,
                                                       p->fifo_cap, nsamples, nr);
        MCAG_CHECK_LAUNCH();
        p->launches++;
      }
    }
    if (T > 0) OK(run_frames(p, cur, p->fifo_cap, T, b0, b1 - b0));
    if (want_audio && nout > 0) {
      const float *src = p->out_dev.as<float>() + (long long)b0 * Cs * nout;
      if constexpr (!is_f32) {
        // narrow on the device, then copy the caller's sample type straight into the caller's rows
        const long long cnt = (long long)(b1 - b0) * Cs * nout;
        Tio *dsto = p->stage_out.as<Tio>() + (long long)b0 * Cs * nout;
        convert_out_kernel<Tio><<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(src, dsto, cnt);
        MCAG_CHECK_LAUNCH();
        p->launches++;
      }
      CU(cudaEventRecord(p->ev_done[c], st));
      CU(cudaStreamWaitEvent(sout, p->ev_done[c], 0));
      const Tio *srco = is_f32 ? reinterpret_cast<const Tio *>(src) : p->stage_out.as<Tio>() + (long long)b0 * Cs * nout;
      if (out_packed && Cs == Cout) {
        CU(cudaMemcpy2DAsync(out_packed + (long long)b0 * Cout * out_pitch, out_pitch * sizeof(Tio), srco, (size_t)nout * sizeof(Tio), (size_t)nout * sizeof(Tio),
                             (size_t)(b1 - b0) * Cs, cudaMemcpyDeviceToHost, sout));
      } else if (out_packed) {   // Cs < Cout: the first Cs rows of every stream, one strided copy per channel
        for (int ch = 0; ch < Cs; ++ch)
          CU(cudaMemcpy2DAsync(out_packed + ((long long)b0 * Cout + ch) * out_pitch, (size_t)Cout * out_pitch * sizeof(Tio), srco + (long long)ch * nout,
                               (size_t)Cs * nout * sizeof(Tio), (size_t)nout * sizeof(Tio), (size_t)(b1 - b0), cudaMemcpyDeviceToHost, sout));
      } else {
        for (int b = b0; b < b1; ++b)
          for (int ch = 0; ch < Cs; ++ch) {
            Tio *dst = out[(long long)b * Cout + ch];
            if (dst) CU(cudaMemcpyAsync(dst, srco + ((long long)(b - b0) * Cs + ch) * nout, (size_t)nout * sizeof(Tio), cudaMemcpyDeviceToHost, sout));
          }
      }
    }
  }
  if (T > 0 && Cs > 0) p->tail_cur ^= 1;
  const long long have = (long long)p->fill + nsamples, consumed = (long long)T * p->hop, left = have - consumed;
  if (left > 0 && consumed > 0) {
    float *nxt = p->fifo[p->fifo_cur ^ 1].as<float>();
    CU(cudaMemcpy2DAsync(nxt, p->fifo_cap * 4, cur + consumed, p->fifo_cap * 4, (size_t)left * 4, rows, cudaMemcpyDeviceToDevice, st));
    p->fifo_cur ^= 1;
  }
  p->fill = (int)left;
  p->frames_last = T;
  p->frames_total += T;
  if (want_audio && nout > 0 && Cout > Cs) {
    // channels >= Cs are zeros (BeamformingSeparationAndLocalisation.cpp:117-118): filled by the host while the device works
    std::vector<Tio *> zr;
    for (int b = 0; b < B; ++b)
      for (int ch = Cs; ch < Cout; ++ch) {
        Tio *dst = out_packed ? out_packed + ((long long)b * Cout + ch) * out_pitch : out[(long long)b * Cout + ch];
        if (dst) zr.push_back(dst);
      }
    const size_t row_bytes = (size_t)nout * sizeof(Tio);
    const int nthr = (zr.size() * row_bytes >= ((size_t)8 << 20)) ? 8 : 1;
    auto work = [&](int w) { for (size_t i = (size_t)w; i < zr.size(); i += (size_t)nthr) std::memset(zr[i], 0, row_bytes); };
    if (nthr == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (int w = 1; w < nthr; ++w) th.emplace_back(work, w);
      work(0);
      for (auto &t_ : th) t_.join();
    }
  }
  CU(cudaStreamSynchronize(st));
  if (want_audio && nout > 0) {
    CU(cudaStreamSynchronize(sout));
  }
  if (nsamples_out) *nsamples_out = Cs > 0 ? nout : 0;
  return MCAG_OK;
}

extern "C" {

int mcag_process_f32(mcag_proc p, const float *const *in, int nsamples, float *const *out, int out_capacity, int *nsamples_out) {
  return process_host<float>(p, in, nullptr, 0, nsamples, out, nullptr, 0, out_capacity, nsamples_out);
}
int mcag_process_f64(mcag_proc p, const double *const *in, int nsamples, double *const *out, int out_capacity, int *nsamples_out) {
  return process_host<double>(p, in, nullptr, 0, nsamples, out, nullptr, 0, out_capacity, nsamples_out);
}
int mcag_process_s16(mcag_proc p, const int16_t *const *in, int nsamples, int16_t *const *out, int out_capacity, int *nsamples_out) {
  return process_host<int16_t>(p, in, nullptr, 0, nsamples, out, nullptr, 0, out_capacity, nsamples_out);
}
int mcag_process_packed_f32(mcag_proc p, const float *in, long long in_pitch, int nsamples, float *out, long long out_pitch, int *nsamples_out) {
  return process_host<float>(p, nullptr, in, in_pitch, nsamples, nullptr, out, out_pitch, (int)(out_pitch > 0x7fffffff ? 0x7fffffff : out_pitch), nsamples_out);
}

int mcag_process_packed_s16(mcag_proc p, const int16_t *in, long long in_pitch, int nsamples, int16_t *out, long long out_pitch, int *nsamples_out) {
  return process_host<int16_t>(p, nullptr, in, in_pitch, nsamples, nullptr, out, out_pitch, (int)(out_pitch > 0x7fffffff ? 0x7fffffff : out_pitch), nsamples_out);
}

int mcag_process_device_f32(mcag_proc p, const float *d_in, long long in_pitch, int nsamples, float *d_out, long long out_pitch, int *nsamples_out) {
  if (!p) return mcag_set_error(MCAG_ERR_INVALID, "null handle");
  CU(cudaSetDevice(p->cfg.device));
  int T = 0;
  OK(process_device_core(p, d_in, in_pitch, nsamples, &T, d_out, out_pitch));
  const int nout = T * p->hop;
  if (nsamples_out) *nsamples_out = p->Cs > 0 ? nout : 0;
  return MCAG_OK;
}

static const DevBuf *result_buf(mcag_proc p, int what) {
  switch (what) {
    case MCAG_OUT_SPECTRA: return &p->spec;
    case MCAG_OUT_POWER_DB: return &p->power_db;
    case MCAG_OUT_CORR: return &p->corr;
    case MCAG_OUT_ENERGY: return &p->energy;
    case MCAG_OUT_CELL: return &p->cells;
    case MCAG_OUT_PROB: return p->cfg.kind == MCAG_KIND_FREQGCC ? &p->track_prob : &p->prob;
    case MCAG_OUT_TRACK_DOA: return &p->track_doa;
    case MCAG_OUT_LAGS: return &p->lags;
    case MCAG_OUT_CURVES: return &p->curves;
    case MCAG_OUT_ACTIVE: return &p->active;
    case MCAG_OUT_BEAMS: return p->cfg.kind == MCAG_KIND_MASK ? &p->spec : &p->beams;
    case MCAG_OUT_MASK_Q: return &p->qtrace;
    case MCAG_OUT_MASK_DEC: return &p->dec;
    case MCAG_OUT_BAND_CELL: return &p->band_cells;
  }
  return nullptr;
}
const void *mcag_device_ptr(mcag_proc p, int what) {
  if (!p) return nullptr;
  const DevBuf *b = result_buf(p, what);
  return b ? b->p : nullptr;
}
int mcag_fetch(mcag_proc p, int what, void *dst, long long bytes) {
  if (!p || !dst) return mcag_set_error(MCAG_ERR_INVALID, "null argument");
  const DevBuf *b = result_buf(p, what);
  if (!b || !b->p) return mcag_set_error(MCAG_ERR_INVALID, "result not produced by this processor kind / emit flags");
  if (bytes < 0 || (size_t)bytes > b->bytes) return mcag_set_error(MCAG_ERR_CAPACITY, "fetch: more bytes than the result holds");
  CU(cudaSetDevice(p->cfg.device));
  CU(cudaMemcpyAsync(dst, b->p, (size_t)bytes, cudaMemcpyDeviceToHost, p->stream));
  CU(cudaStreamSynchronize(p->stream));
  return MCAG_OK;
}

// ---- kernel-level wrappers ------------------------------------------------------------------------------------------
int mcag_k_twiddle_count(int N) { return (N == 256 || N == 512 || N == 1024 || N == 2048) ? fft_table_len(N) : 0; }
int mcag_k_twiddles(int N, void *d_tw, void *stream) {
  if (!mcag_k_twiddle_count(N)) return mcag_set_error(MCAG_ERR_INVALID, "twiddles: frame size must be 256, 512, 1024 or 2048");
  std::vector<float2> tw(fft_table_len(N));
  for (int n = 0; n < N / 2; ++n) { double a = -2.0 * M_PI * (double)n / (double)N; tw[n] = make_float2((float)std::cos(a), (float)std::sin(a)); }
  switch (N) {
    case 256: fill_thread_twiddles<128>(tw); break;
    case 512: fill_thread_twiddles<256>(tw); break;
    case 1024: fill_thread_twiddles<512>(tw); break;
    default: fill_thread_twiddles<1024>(tw); break;
  }
  CU(cudaMemcpyAsync(d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return MCAG_OK;
}
int mcag_k_phase_fx(const double *h_turns, long long n, uint64_t *d_fx, void *stream) {
  std::vector<uint64_t> fx((size_t)n);
  for (long long i = 0; i < n; ++i) fx[(size_t)i] = turns_to_fx(h_turns[i]);
  CU(cudaMemcpyAsync(d_fx, fx.data(), (size_t)n * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return MCAG_OK;
}
int mcag_k_stft(const float *d_x, long long row_pitch, int rows, int M, int T, int N, int hop, const float *d_win, const void *d_tw, void *d_spec,
                float *d_chan_pow, void *stream) {
  return k_stft(d_x, row_pitch, rows, M, T, N, hop, d_win, (const float2 *)d_tw, (float2 *)d_spec, d_chan_pow, nullptr, (cudaStream_t)stream);
}
int mcag_k_istft(const void *d_spec, int B, int T, int C_in, int C_out, int N, int hop, const float *d_win, const void *d_tw, const float *d_tail_in,
                 float *d_tail_out, float *d_out, long long out_pitch, void *stream) {
  return k_istft((const float2 *)d_spec, B, T, C_in, C_out, N, hop, d_win, (const float2 *)d_tw, d_tail_in, d_tail_out, d_out, out_pitch, C_out, (cudaStream_t)stream);
}
int mcag_k_tdoa_lags(const void *d_spec, int B, int T, int M, int N, int max_lag, const void *d_tw, float *d_curves, int32_t *d_lags, float *d_peaks,
                     void *stream) {
  return k_tdoa_lags((const float2 *)d_spec, B, T, M, N, max_lag, (const float2 *)d_tw, d_curves, d_lags, d_peaks, (cudaStream_t)stream);
}
int mcag_k_stft_tdoa(const float *d_x, long long row_pitch, int B, int T, int M, int N, int hop, int max_lag, const float *d_win, const void *d_tw,
                     void *d_spec, float *d_chan_pow, float *d_curves, int32_t *d_lags, void *stream) {
  return k_stft_tdoa(d_x, row_pitch, B, T, M, N, hop, max_lag, d_win, (const float2 *)d_tw, (float2 *)d_spec, d_chan_pow, d_curves, d_lags, (cudaStream_t)stream);
}
int mcag_k_gcc_tau(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_pair_fx, int D, float *d_corr, void *stream) {
  return k_gcc_tau((const float2 *)d_spec, B, T, M, N, d_pair_fx, D, d_corr, (cudaStream_t)stream);
}
int mcag_k_gcc_tau_tensor(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_pair_fx, int D, float *d_corr, void *stream) {
  if (!k_gcc_tau_tc_supported(D)) return mcag_set_error(MCAG_ERR_INVALID, "gcc_tau_tensor: at most 64 delays");
  cudaStream_t st = (cudaStream_t)stream;
  void *tab = nullptr;
  const int P = M * (M - 1) / 2;
  CU(cudaMallocAsync(&tab, k_gcc_tau_tc_table_bytes(P, N), st));
  int rc = k_gcc_tau_tc_build(d_pair_fx, P, D, N, (float *)tab, st);
  if (!rc) rc = k_gcc_tau_tc((const float2 *)d_spec, B, T, M, N, (const float *)tab, D, d_corr, st);
  cudaFreeAsync(tab, st);
  return rc;
}
int mcag_k_pair_sum(const float *d_corr, long long BT, int P, int D, float scale, float *d_esum, void *stream) {
  return k_pair_sum(d_corr, BT, P, D, scale, d_esum, (cudaStream_t)stream);
}
int mcag_k_energy_scan(const float *d_esum, int B, int T, int D, float a, const unsigned char *d_active, float *d_state, float *d_energy, void *stream) {
  return k_energy_scan(d_esum, B, T, D, a, d_active, d_state, d_energy, (cudaStream_t)stream);
}
int mcag_k_select_doa(const float *d_energy, long long BT, int D, int n_pairs, int S, int32_t *d_idx, float *d_prob, void *stream) {
  return k_select_doa(d_energy, BT, D, n_pairs, S, d_idx, d_prob, (cudaStream_t)stream);
}
int mcag_k_argmax_pack(const float *d_map, long long rows, int D, int d_offset, long long *d_packed, void *stream) {
  return k_argmax_pack(d_map, rows, D, d_offset, d_packed, (cudaStream_t)stream);
}
int mcag_k_ds_fan(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_steer_fx, int D, void *d_out, void *stream) {
  return k_ds_fan((const float2 *)d_spec, B, T, M, N, d_steer_fx, D, (float2 *)d_out, (cudaStream_t)stream);
}
int mcag_k_fs_fan(const void *d_spec, int B, int T, int M, int N, const void *d_weights, int D, void *d_out, void *stream) {
  return k_fs_fan((const float2 *)d_spec, B, T, M, N, (const float2 *)d_weights, D, (float2 *)d_out, (cudaStream_t)stream);
}
int mcag_k_srp_channel(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_mic_fx, int D, float *d_srp, void *stream) {
  return k_srp_channel((const float2 *)d_spec, B, T, M, N, d_mic_fx, D, d_srp, (cudaStream_t)stream);
}

}  // extern "C"

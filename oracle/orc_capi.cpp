// TEST INFRASTRUCTURE ONLY (oracle).  extern "C" surface of the restatement (oracle/restated.hpp)
// -> oracle/_build/liboracle.so.  Loaded by tests/, __graft_entry__.smoke() and bench.py's CPU legs
// through ctypes; never by the product path.
#define ORC_PREFIX orc_
#include "capi.h"
#include "restated.hpp"

#include <cstring>
#include <thread>

using namespace orc;

double orc_doa_idx_to_angle(int idx, float doa_step) { return doa_idx_to_angle(idx, doa_step); }
double orc_angle_to_doa_idx(float angle, float doa_step) { return angle_to_doa_idx(angle, doa_step); }
double orc_doa_to_delay_samples(float doa, float mic_dist, int fs) { return doa_to_delay_far_field_samples(doa, mic_dist, fs); }
double orc_array_distance(const double *xyz, int, int i, int j) { return mic_distance(xyz, i, j); }
double orc_array_max_distance(const double *xyz, int M) { return max_mic_distance(xyz, M); }
int orc_frame_size(int fs, double frame_rate) { return 1 << dsp::ShortTimeProcess::calculateOrderFromSampleRate(fs, frame_rate); }

// ---------------------------------------------------------------------------------------------
namespace {
struct SslRun {
  SslState st; int frame = 0, fired = 0, max_frames = 0; bool analysis_only = false;
  int *fired_frame; double *doa_deg, *prob, *power, *energy, *corr_scaled;
};
void ssl_hook(void *user, double *frames, int, int) {
  SslRun &r = *static_cast<SslRun *>(user);
  const int D = r.st.steer.D, P = r.st.steer.P, S = r.st.S;
  std::vector<double> energy(static_cast<size_t>(D)), corr(size_t(P) * D);
  std::vector<int> idx(static_cast<size_t>(S));
  FrameReport rep = ssl_localise(r.st, frames, energy.data(), idx.data(), corr.data());
  if (rep.fired && r.fired < r.max_frames) {
    const int f = r.fired;
    r.fired_frame[f] = r.frame;
    for (int s = 0; s < S; ++s) {
      r.doa_deg[size_t(f) * S + s] = r.st.cur_doa[size_t(s)] * (180 / M_PI);       // toDegrees: microhponeArrayHelpers.cpp:91-98
      r.prob[size_t(f) * S + s] = r.st.prob[size_t(s)];
    }
    r.power[f] = rep.power;
    std::copy(energy.begin(), energy.end(), r.energy + size_t(f) * D);
    if (r.corr_scaled) {
      const double b = double(1 - 0.8f);
      for (size_t i = 0; i < corr.size(); ++i) r.corr_scaled[size_t(f) * P * D + i] = b * corr[i];
    }
  }
  if (rep.fired) ++r.fired;
  if (!r.analysis_only) ssl_separate(r.st, frames);
  ++r.frame;
}
template <class Proc>
int feed(Proc &proc, int M, const double *in, int n, int chunk, double *out, int out_cap, bool synth) {
  int written = 0;
  if (chunk <= 0) chunk = n;
  std::vector<double> obuf(size_t(M) * size_t(chunk + proc.getMaxLatency()));
  for (int pos = 0; pos < n; pos += chunk) {
    const int len = std::min(chunk, n - pos);
    std::vector<double *> pin, pout;
    for (int c = 0; c < M; ++c) { pin.push_back(const_cast<double *>(in) + size_t(c) * n + pos); pout.push_back(&obuf[size_t(c) * size_t(chunk + proc.getMaxLatency())]); }
    if (synth) {
      int got = proc.process(pin, len, pout, chunk + proc.getMaxLatency());
      if (written + got > out_cap) return -1;
      for (int c = 0; c < M; ++c) std::copy(pout[size_t(c)], pout[size_t(c)] + got, out + size_t(c) * out_cap + written);
      written += got;
    } else {
      proc.process(pin, len);
    }
  }
  return written;
}
}  // namespace

int orc_ssl_run(int fs, int M, const double *mic_xyz, int S, int use_floor, int analysis_only,
                const double *in, int n, int chunk, double *out, int out_cap, int *n_out,
                int max_frames, int *n_frames, int *n_fired, int *fired_frame,
                double *doa_deg, double *prob, double *power, double *energy, double *corr_scaled) {
  const int order = dsp::ShortTimeProcess::calculateOrderFromSampleRate(fs, 0.025f);     // _frameRate, SourceSeparationAndLocalisation.h:60
  SslRun r;
  r.st.init(fs, mic_xyz, M, 1 << order, S, use_floor != 0);
  r.max_frames = max_frames; r.analysis_only = analysis_only != 0;
  r.fired_frame = fired_frame; r.doa_deg = doa_deg; r.prob = prob; r.power = power; r.energy = energy; r.corr_scaled = corr_scaled;
  FrameTap tap(M, order, !analysis_only, ssl_hook, &r);
  int w = feed(tap, M, in, n, chunk, out, out_cap, !analysis_only);
  if (w < 0) return -1;
  if (n_out) *n_out = w;
  if (n_frames) *n_frames = r.frame;
  if (n_fired) *n_fired = r.fired;
  return 1 << order;
}

// ---------------------------------------------------------------------------------------------
namespace {
struct GccRun { FreqGccState st; int frame = 0, fired = 0, max_frames = 0; int *fired_frame, *idx; double *curves, *power; };
void gcc_hook(void *user, double *frames, int, int) {
  GccRun &r = *static_cast<GccRun *>(user);
  std::vector<double> curve(static_cast<size_t>(r.st.D)); int idx = 0;
  FrameReport rep = freqgcc_frame(r.st, frames, curve.data(), &idx);
  if (rep.fired && r.fired < r.max_frames) {
    r.fired_frame[r.fired] = r.frame; r.idx[r.fired] = idx; r.power[r.fired] = rep.power;
    std::copy(curve.begin(), curve.end(), r.curves + size_t(r.fired) * r.st.D);
  }
  if (rep.fired) ++r.fired;
  ++r.frame;
}
}  // namespace

int orc_freqgcc_run(int fs, double mic_dist, int use_floor, int noise_preestimated, const double *in, int n, int chunk,
                    int max_frames, int *n_frames, int *n_fired, int *fired_frame, double *curves, int *idx, double *power) {
  const int order = dsp::ShortTimeProcess::calculateOrderFromSampleRate(fs, 0.075f);     // BinauralLocalisation.h:196
  GccRun r;
  r.st.init(fs, mic_dist, 1 << order, use_floor != 0);
  if (noise_preestimated) r.st.noise_estimated = true;
  r.max_frames = max_frames; r.fired_frame = fired_frame; r.idx = idx; r.curves = curves; r.power = power;
  FrameTap tap(2, order, false, gcc_hook, &r);
  feed(tap, 2, in, n, chunk, nullptr, 0, false);
  if (n_frames) *n_frames = r.frame;
  if (n_fired) *n_fired = r.fired;
  return 1 << order;
}

// ---------------------------------------------------------------------------------------------
namespace {
struct MbRun { MultibandState st; int frame = 0, fired = 0, max_frames = 0; int *fired_frame, *cell, *band_cells; double *prob, *power, *doa_deg, *hist; };
void mb_hook(void *user, double *frames, int, int) {
  MbRun &r = *static_cast<MbRun *>(user);
  std::vector<double> hist(static_cast<size_t>(r.st.D)); std::vector<int> bc(static_cast<size_t>(r.st.nb)); int cell = 0;
  FrameReport rep = multiband_frame(r.st, frames, hist.data(), bc.data(), &cell);
  if (rep.fired && r.fired < r.max_frames) {
    const int f = r.fired;
    r.fired_frame[f] = r.frame; r.cell[f] = cell; r.prob[f] = r.st.prob; r.power[f] = rep.power;
    r.doa_deg[f] = r.st.cur_doa * (180 / M_PI);
    std::copy(hist.begin(), hist.end(), r.hist + size_t(f) * r.st.D);
    std::copy(bc.begin(), bc.end(), r.band_cells + size_t(f) * r.st.nb);
  }
  if (rep.fired) ++r.fired;
  ++r.frame;
}
}  // namespace

int orc_multiband_run(int fs, double mic_dist, int nbins, int use_floor, int noise_preestimated, const double *in, int n, int chunk,
                      int max_frames, int *n_frames, int *n_fired, int *fired_frame, int *n_dirs,
                      int *cell, double *prob, double *power, double *doa_deg, double *hist, int *band_cells) {
  const int order = dsp::ShortTimeProcess::calculateOrderFromSampleRate(fs, 0.025f);     // _frameRate, MultibandBinarualLocalisation.h:41
  MbRun r;
  r.st.init(fs, mic_dist, 1 << order, nbins, use_floor != 0);
  if (n_dirs) *n_dirs = r.st.D;
  if (!in) return 1 << order;
  if (noise_preestimated) r.st.noise_estimated = true;
  r.max_frames = max_frames; r.fired_frame = fired_frame; r.cell = cell; r.band_cells = band_cells; r.prob = prob; r.power = power; r.doa_deg = doa_deg; r.hist = hist;
  FrameTap tap(2, order, false, mb_hook, &r);
  feed(tap, 2, in, n, chunk, nullptr, 0, false);
  if (n_frames) *n_frames = r.frame;
  if (n_fired) *n_fired = r.fired;
  return 1 << order;
}

// frame-level entry for the GPU tests: T spectra [T][2][N+2] from a fresh state (floor pre-estimated, gate off)
extern "C" int orc_multiband_frames(const double *spec, int T, int N, int fs, double mic_dist, int nbins, int *cell, double *prob, double *hist,
                                    int *band_cells, double *H_out /*[nbins][N/2+1] or NULL*/) {
  MultibandState st;
  st.init(fs, mic_dist, N, nbins, false);
  st.noise_estimated = true;
  if (H_out) std::copy(st.H.begin(), st.H.end(), H_out);
  for (int t = 0; t < T && spec; ++t) {
    multiband_frame(st, spec + size_t(t) * 2 * (N + 2), hist + size_t(t) * st.D, band_cells + size_t(t) * nbins, cell + t);
    prob[t] = st.prob;
  }
  return st.D;
}

// FreqGCC with the deterministic DOA tracker (the `#else` branch of USE_PARTICLE_FILTER, BinauralLocalisation.cpp:501-504) and
// setProbability (:454,569-631): per frame t (every frame, fired or not) active[t], doa_rad[t] = _currentDOA, prob[t] = _prob,
// idx[t] / curves[t][61] valid on active frames (previous values are held otherwise).
namespace {
struct GccTrack { FreqGccState st; int frame = 0, max_frames = 0; int *active, *idx; double *doa, *prob, *curves, *power; int last_idx = 0; };
void gcc_track_hook(void *user, double *frames, int, int) {
  GccTrack &r = *static_cast<GccTrack *>(user);
  std::vector<double> curve(static_cast<size_t>(r.st.D)); int idx = r.last_idx;
  FrameReport rep = freqgcc_frame(r.st, frames, curve.data(), &idx);
  if (r.frame < r.max_frames) {
    const int t = r.frame;
    r.active[t] = rep.fired ? 1 : 0; r.doa[t] = r.st.cur_doa; r.prob[t] = r.st.prob; r.power[t] = rep.power;
    if (rep.fired) { r.last_idx = idx; std::copy(curve.begin(), curve.end(), r.curves + size_t(t) * r.st.D); }
    else std::copy(r.st.prev_corr.begin(), r.st.prev_corr.end(), r.curves + size_t(t) * r.st.D);
    r.idx[t] = r.last_idx;
  }
  ++r.frame;
}
}  // namespace
extern "C" int orc_freqgcc_track_run(int fs, double mic_dist, int N, int use_floor, int noise_preestimated, const double *in, int n, int chunk,
                                     int max_frames, int *n_frames, int *active, double *doa_rad, double *prob, int *idx, double *curves, double *power) {
  int order = 0;
  if (N > 0) { while ((1 << order) < N) ++order; } else order = dsp::ShortTimeProcess::calculateOrderFromSampleRate(fs, 0.075f);
  GccTrack r;
  r.st.init(fs, mic_dist, 1 << order, use_floor != 0);
  if (noise_preestimated) r.st.noise_estimated = true;
  r.max_frames = max_frames; r.active = active; r.doa = doa_rad; r.prob = prob; r.idx = idx; r.curves = curves; r.power = power;
  FrameTap tap(2, order, false, gcc_track_hook, &r);
  feed(tap, 2, in, n, chunk, nullptr, 0, false);
  if (n_frames) *n_frames = r.frame;
  return 1 << order;
}

void orc_freqgcc_probability(int, double, const double *curve, const double *doas, double *probs, int size) {
  const float step = float(3 * M_PI / 180);
  freqgcc_probability(curve, num_doa_steps(step), step, doas, probs, size);
}

// ---------------------------------------------------------------------------------------------
namespace {
struct MaskRun { MaskState st; int frame = 0, max_frames = 0; double *Q, *spectra; };
void mask_hook(void *user, double *frames, int, int ccs) {
  MaskRun &r = *static_cast<MaskRun *>(user);
  mask_frame(r.st, frames, nullptr);
  if (r.frame < r.max_frames) {
    if (r.Q) std::copy(r.st.Q.begin(), r.st.Q.end(), r.Q + size_t(r.frame) * r.st.n_bands);
    if (r.spectra) std::copy(frames, frames + 2 * ccs, r.spectra + size_t(r.frame) * 2 * ccs);
  }
  ++r.frame;
}
}  // namespace

extern "C" void orc_mel_bank(int N, int n_bands, int fs, float lo, float hi, double *H, double *fc_norm) {
  int order = 0; while ((1 << order) < N) ++order;
  dsp::FilterBankFFTWMelScale fb(order, n_bands, fs, lo, hi);
  fb.getFiltersCoeficients(H, n_bands * (N / 2 + 1));
  for (int b = 0; b < n_bands; ++b) fc_norm[b] = fb.getBinCenterFrequency(b);
}

int orc_mask_run(int fs, double mic_dist, float lo, float hi, int method, int alg,
                 const double *in, int n, int chunk, double *out, int out_cap, int *n_out,
                 int max_frames, int *n_frames, double *Q, double *spectra_out) {
  const int order = dsp::ShortTimeProcess::calculateOrderFromSampleRate(fs, 0.050f);     // FastBinauralMasking.h:112
  const int N = 1 << order, nb = 45;                                                      // .h:111
  std::vector<double> H(size_t(nb) * (N / 2 + 1)), fc(static_cast<size_t>(nb));
  orc_mel_bank(N, nb, fs, lo, hi, H.data(), fc.data());
  MaskRun r;
  r.st.init(fs, mic_dist, N, method, alg, nb, H.data(), fc.data());
  r.max_frames = max_frames; r.Q = Q; r.spectra = spectra_out;
  FrameTap tap(2, order, true, mask_hook, &r);
  int w = feed(tap, 2, in, n, chunk, out, out_cap, true);
  if (w < 0) return -1;
  if (n_out) *n_out = w;
  if (n_frames) *n_frames = r.frame;
  return N;
}

// ---------------------------------------------------------------------------------------------
int orc_beamformer_frame(int fs, int M, const double *mic_xyz, int ccs_len, const double *frames, double doa, double *out) {
  std::vector<double> x(static_cast<size_t>(M));
  for (int m = 0; m < M; ++m) x[size_t(m)] = mic_xyz[3 * m];
  ds_beamform(frames, M, ccs_len, fs, x.data(), doa, out);
  return 0;
}

int orc_steering_frames(int fs, int M, const double *mic_xyz, int ccs_len, int S, const double *frames, int T,
                        double *doa_rad, double *prob, double *energy) {
  SslState st;
  st.init(fs, mic_xyz, M, ccs_len - 2, S, false);
  const int D = st.steer.D;
  std::vector<double> corr(size_t(st.steer.P) * D), e(static_cast<size_t>(D));
  std::vector<int> idx(static_cast<size_t>(S));
  for (int t = 0; t < T; ++t) {
    steering_correlations(st.steer, frames + size_t(t) * M * ccs_len, corr.data());
    steering_energy(st.steer, corr.data(), e.data());
    steering_select(st.steer, e.data(), S, idx.data(), doa_rad + size_t(t) * S, prob + size_t(t) * S);
    std::copy(e.begin(), e.end(), energy + size_t(t) * D);
  }
  return D;
}

// =============================================================================================
// orc_-only: building blocks with explicit conventions (window / hop / grids as inputs) and the
// generalisations BASELINE.json's configs 2-4 need (SURVEY.md §8a "Generalisations").  Each one
// reduces to the reference functions above on a linear x-axis array with the 37-point grid
// (tests/test_oracle.py checks that).
// =============================================================================================

extern "C" void orc_sqrt_hann(int N, double *w) {
  for (int n = 0; n < N; ++n) w[n] = std::sqrt(0.5 * (1.0 - std::cos(2.0 * M_PI * double(n) / double(N))));
}

// A1: in [M][n] -> spec [T][M][N+2]; T = floor((n-N)/hop)+1.  Returns T (or the count it would need).
extern "C" int orc_stft(const double *in, int M, int n, int N, int hop, const double *win, double *spec, int max_frames) {
  if (n < N) return 0;
  const int T = (n - N) / hop + 1;
  if (!spec) return T;
  int order = 0; while ((1 << order) < N) ++order;
  dsp::FFT fft(order);
  std::vector<double> w(static_cast<size_t>(N)), frame(static_cast<size_t>(N));
  if (win) std::copy(win, win + N, w.begin()); else orc_sqrt_hann(N, w.data());
  const int ccs = N + 2;
  for (int t = 0; t < std::min(T, max_frames); ++t)
    for (int m = 0; m < M; ++m) {
      const double *src = in + size_t(m) * n + size_t(t) * hop;
      for (int i = 0; i < N; ++i) frame[size_t(i)] = src[i] * w[size_t(i)];
      fft.fwdTransform(frame.data(), spec + (size_t(t) * M + m) * ccs);
    }
  return T;
}

// A11: spec [T][C][N+2] -> out [C][T*hop]; `tail` [C][N-hop] carries the overlap in and out (may be NULL = zeros).
extern "C" void orc_istft(const double *spec, int T, int C, int N, int hop, const double *win, double *out, double *tail) {
  int order = 0; while ((1 << order) < N) ++order;
  dsp::FFT fft(order);
  std::vector<double> w(static_cast<size_t>(N)), frame(static_cast<size_t>(N));
  if (win) std::copy(win, win + N, w.begin()); else orc_sqrt_hann(N, w.data());
  const int ccs = N + 2, ov = N - hop;
  for (int c = 0; c < C; ++c) {
    std::vector<double> acc(size_t(T) * hop + size_t(ov), 0.0);
    if (tail) for (int i = 0; i < ov; ++i) acc[size_t(i)] = tail[size_t(c) * ov + i];
    for (int t = 0; t < T; ++t) {
      fft.invTransfrom(frame.data(), spec + (size_t(t) * C + c) * ccs);
      for (int i = 0; i < N; ++i) acc[size_t(t) * hop + i] += frame[size_t(i)] * w[size_t(i)];
    }
    std::copy(acc.begin(), acc.begin() + long(T) * hop, out + size_t(c) * T * hop);
    if (tail) for (int i = 0; i < ov; ++i) tail[size_t(c) * ov + i] = acc[size_t(T) * hop + i];
  }
}

// SignalPower::FFTLogPower per frame: spec [T][M][ccs] -> power_db [T]
extern "C" void orc_fft_log_power(const double *spec, int T, int M, int N, double *power_db) {
  const int ccs = N + 2;
  for (int t = 0; t < T; ++t) {
    std::vector<const double *> fv;
    for (int m = 0; m < M; ++m) fv.push_back(spec + (size_t(t) * M + m) * ccs);
    power_db[t] = dsp::SignalPower::FFTLogPower(fv, ccs);
  }
}

// A2/A4 with an explicit pair tau table [P][D]: corr [T][P][D] = Re GCC-PHAT, pairs i<j lexicographic.
extern "C" void orc_gcc_tau_frames(const double *spec, int T, int M, int N, const double *pair_tau, int D, double *corr) {
  const int ccs = N + 2, K = N / 2 + 1, P = M * (M - 1) / 2;
  for (int t = 0; t < T; ++t) {
    int p = 0;
    for (int i = 0; i < M; ++i)
      for (int j = i + 1; j < M; ++j, ++p)
        gcc_phat_tau(spec + (size_t(t) * M + i) * ccs, spec + (size_t(t) * M + j) * ccs, K, pair_tau + size_t(p) * D, D,
                     corr + (size_t(t) * P + p) * D);
  }
}

// config 2: integer-lag TDOA on all pairs.  curves [T][P][2L+1], lags [T][P].  Uses the precomputed tau matrix exactly as
// SteeringBeamforming drives dsp::GeneralisedCrossCorrelation (precomputeTauMatrix + calculateCorrelationsForPrecomputedTauMatrix,
// SteeringBeamforming.cpp:84-88,115-122) with tau = -L..L, then wipp::maxidx (first maximum).
extern "C" void orc_tdoa_lags(const double *spec, int T, int M, int N, int max_lag, double *curves, int *lags) {
  const int ccs = N + 2, K = N / 2 + 1, P = M * (M - 1) / 2, L = 2 * max_lag + 1;
  std::vector<double> tau(static_cast<size_t>(L)), re(static_cast<size_t>(L));
  for (int l = -max_lag; l <= max_lag; ++l) tau[size_t(l + max_lag)] = double(l);
  dsp::GeneralisedCrossCorrelation gcc(K, dsp::GeneralisedCrossCorrelation::ONESIDEDFFT);
  gcc.precomputeTauMatrix(tau.data(), L, K, dsp::GeneralisedCrossCorrelation::ONESIDEDFFT);
  std::vector<dsp::Complex> cc(static_cast<size_t>(L));
  for (int t = 0; t < T; ++t) {
    int p = 0;
    for (int i = 0; i < M; ++i)
      for (int j = i + 1; j < M; ++j, ++p) {
        gcc.calculateCorrelationsForPrecomputedTauMatrix(reinterpret_cast<const dsp::Complex *>(spec + (size_t(t) * M + i) * ccs),
                                                         reinterpret_cast<const dsp::Complex *>(spec + (size_t(t) * M + j) * ccs), cc.data(), K, L,
                                                         dsp::GeneralisedCrossCorrelation::ONESIDEDFFT);
        wipp::real(reinterpret_cast<const wipp::wipp_complex_t *>(cc.data()), re.data(), size_t(L));
        double mx; size_t idx;
        wipp::maxidx(re.data(), size_t(L), &mx, &idx);
        lags[size_t(t) * P + p] = int(idx) - max_lag;
        if (curves) std::copy(re.begin(), re.end(), curves + (size_t(t) * P + p) * L);
      }
  }
}

// A5 with explicit constants: corr [T][P][D] -> energy [T][D]; state [D] in/out; active [T] (NULL = all).
extern "C" void orc_energy_scan(const double *corr, int T, int P, int D, double a, double b, const unsigned char *active, double *state, double *energy) {
  for (int t = 0; t < T; ++t) {
    double *e = energy + size_t(t) * D;
    if (active && !active[t]) { std::copy(state, state + D, e); continue; }
    for (int d = 0; d < D; ++d) e[d] = a * state[d];
    for (int p = 0; p < P; ++p) for (int d = 0; d < D; ++d) { double c = b * corr[(size_t(t) * P + p) * D + d]; e[d] += c; }
    std::copy(e, e + D, state);
  }
}

// A6 on a general grid size: energy [T][D] -> idx [T][S] (cell = maxidx+1), prob [T][S]; n_pairs sets m = -15*P.
extern "C" void orc_select_doa(const double *energy, int T, int D, int n_pairs, int S, int *idx, double *prob) {
  SteeringState st; st.D = D; st.P = n_pairs; st.doa_step = 0;
  std::vector<double> doa(static_cast<size_t>(S));
  for (int t = 0; t < T; ++t) steering_select(st, energy + size_t(t) * D, S, idx + size_t(t) * S, doa.data(), prob + size_t(t) * S);
}

// A7 to a fan of D azimuths (config 3): out [T][D][ccs]
extern "C" void orc_ds_fan(const double *spec, int T, int M, int N, int fs, const double *mic_x, const double *doas, int D, double *out) {
  const int ccs = N + 2;
  for (int t = 0; t < T; ++t)
    for (int d = 0; d < D; ++d) ds_beamform(spec + size_t(t) * M * ccs, M, ccs, fs, mic_x, doas[d], out + (size_t(t) * D + d) * ccs);
}

// Filter-and-sum fan (BASELINE.json north_star; no reference class): Beamformer::processFrame (Beamformer.cpp:56-70) with the generated
// phasor exp(j k phi_c) replaced by a loaded complex weight W[d][c][k]: channels added in index order, then divided by M.
// spec [T][M][ccs], W [D][M][K] complex (re, im), out [T][D][ccs].
extern "C" void orc_fs_fan(const double *spec, int T, int M, int N, const double *W, int D, double *out) {
  const int ccs = N + 2, K = N / 2 + 1;
  for (int t = 0; t < T; ++t)
    for (int d = 0; d < D; ++d) {
      double *o = out + (size_t(t) * D + d) * ccs;
      std::fill(o, o + ccs, 0.0);
      for (int c = 0; c < M; ++c) {
        const double *x = spec + (size_t(t) * M + c) * ccs, *w = W + (size_t(d) * M + c) * 2 * K;
        for (int k = 0; k < K; ++k) {
          o[2 * k] += x[2 * k] * w[2 * k] - x[2 * k + 1] * w[2 * k + 1];
          o[2 * k + 1] += x[2 * k] * w[2 * k + 1] + x[2 * k + 1] * w[2 * k];
        }
      }
      for (int i = 0; i < ccs; ++i) o[i] /= M;
    }
}

// Generalised far-field geometry: per-mic advance tau_m(d) = (p_m . u_d)/c*fs, pair delay
// tau_ij(d) = tau_j(d) - tau_i(d).  On an ascending x-axis array with u = (sin th, cos th, 0) this is the
// reference's dist*sin(th)/c*fs up to its float rounding.
extern "C" void orc_mic_tau(const double *xyz, int M, int fs, const double *dirs, int D, double *mic_tau /*[M][D]*/) {
  for (int m = 0; m < M; ++m)
    for (int d = 0; d < D; ++d)
      mic_tau[size_t(m) * D + d] = (xyz[3 * m] * dirs[3 * d] + xyz[3 * m + 1] * dirs[3 * d + 1] + xyz[3 * m + 2] * dirs[3 * d + 2]) / speed_of_sound() * fs;
}
extern "C" void orc_pair_tau_from_mic_tau(const double *mic_tau, int M, int D, double *pair_tau /*[P][D]*/) {
  int p = 0;
  for (int i = 0; i < M; ++i) for (int j = i + 1; j < M; ++j, ++p)
    for (int d = 0; d < D; ++d) pair_tau[size_t(p) * D + d] = mic_tau[size_t(j) * D + d] - mic_tau[size_t(i) * D + d];
}
extern "C" void orc_reference_pair_tau(const double *xyz, int M, int fs, float doa_step, double *pair_tau, int *D_out) {
  std::vector<double> tau; int D;
  reference_pair_delays(xyz, M, fs, doa_step, tau, D);
  if (pair_tau) std::copy(tau.begin(), tau.end(), pair_tau);
  *D_out = D;
}

// Channel form of the pair sum: sum_{i<j} Re(G_ij e^{+j w tau_ij}) =
// 1/2 (|sum_m U_m e^{-j w tau_m}|^2 - M'), U = X/|X| (0 where |X| = 0), M' = number of non-zero channels
// in that bin.  srp [T][D].  Threads split the direction axis (pure data parallelism, for speed only).
extern "C" void orc_srp_channel(const double *spec, int T, int M, int N, const double *mic_tau, int D, double *srp, int n_threads) {
  const int ccs = N + 2, K = N / 2 + 1;
  std::vector<cd> U(size_t(M) * K);
  if (n_threads < 1) n_threads = 1;
  for (int t = 0; t < T; ++t) {
    std::vector<int> nz(static_cast<size_t>(K), 0);
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) {
        const double *x = spec + (size_t(t) * M + m) * ccs + 2 * k;
        double mag = std::sqrt(x[0] * x[0] + x[1] * x[1]);
        U[size_t(m) * K + k] = mag > 0 ? cd(x[0] / mag, x[1] / mag) : cd(0, 0);
        if (mag > 0) ++nz[size_t(k)];
      }
    auto work = [&](int d0, int d1) {
      for (int d = d0; d < d1; ++d) {
        double acc = 0;
        for (int k = 0; k < K; ++k) {
          cd y(0, 0);
          for (int m = 0; m < M; ++m) y += U[size_t(m) * K + k] * std::polar(1.0, -2.0 * M_PI * double(k) * mic_tau[size_t(m) * D + d] / double(N));
          acc += 0.5 * (std::norm(y) - nz[size_t(k)]);
        }
        srp[size_t(t) * D + d] = acc;
      }
    };
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work, int(int64_t(D) * i / n_threads), int(int64_t(D) * (i + 1) / n_threads));
    for (auto &x : th) x.join();
  }
}

// A10 on given spectra with an explicit filter bank (H [nb][K], fc_norm [nb]); spec [T][2][ccs] in place.
extern "C" void orc_mask_frames(double *spec, int T, int N, int fs, double mic_dist, int method, int alg, int n_bands, const double *H,
                     const double *fc_norm, double *Q_state, double *noise_state, int *first_call, int *decisions /*[T][nb] or NULL*/,
                     double *Q_trace /*[T][nb] or NULL*/) {
  MaskState st;
  st.init(fs, mic_dist, N, method, alg, n_bands, H, fc_norm);
  if (Q_state) st.Q.assign(Q_state, Q_state + n_bands);
  if (noise_state) st.noise_est.assign(noise_state, noise_state + n_bands);
  if (first_call) st.first_call = *first_call;
  for (int t = 0; t < T; ++t) {
    mask_frame(st, spec + size_t(t) * 2 * (N + 2), decisions ? decisions + size_t(t) * n_bands : nullptr);
    if (Q_trace) std::copy(st.Q.begin(), st.Q.end(), Q_trace + size_t(t) * n_bands);
  }
  if (Q_state) std::copy(st.Q.begin(), st.Q.end(), Q_state);
  if (noise_state) std::copy(st.noise_est.begin(), st.noise_est.end(), noise_state);
  if (first_call) *first_call = st.first_call;
}

// A9 on given spectra: spec [T][2][ccs] -> curves [T][D], idx [T]; use_floor = 0 path only.
extern "C" void orc_freqgcc_frames(const double *spec, int T, int N, int fs, double mic_dist, double *curves, int *idx) {
  FreqGccState st; st.init(fs, mic_dist, N, false);
  for (int t = 0; t < T; ++t) freqgcc_frame(st, spec + size_t(t) * 2 * (N + 2), curves + size_t(t) * st.D, idx + t);
}
extern "C" int orc_freqgcc_grid(int fs, double mic_dist, double *tau /*[61] or NULL*/) {
  FreqGccState st; st.init(fs, mic_dist, 512, false);
  if (tau) std::copy(st.tau.begin(), st.tau.end(), tau);
  return st.D;
}

// CPU-baseline helper for bench.py: STFT -> integer-lag GCC-PHAT on all pairs for B independent streams,
// one stream per thread (the reference itself is single-threaded).  in [B][M][n]; lags [B][T][P].
extern "C" int orc_tdoa_pipeline(const double *in, int B, int M, int n, int N, int hop, int max_lag, int *lags, int n_threads) {
  const int T = (n - N) / hop + 1, P = M * (M - 1) / 2;
  if (n_threads < 1) n_threads = 1;
  auto work = [&](int b0, int b1) {
    std::vector<double> spec(size_t(T) * M * (N + 2));
    for (int b = b0; b < b1; ++b) {
      orc_stft(in + size_t(b) * M * n, M, n, N, hop, nullptr, spec.data(), T);
      orc_tdoa_lags(spec.data(), T, M, N, max_lag, nullptr, lags + size_t(b) * T * P);
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) th.emplace_back(work, int(int64_t(B) * i / n_threads), int(int64_t(B) * (i + 1) / n_threads));
  for (auto &x : th) x.join();
  return T;
}


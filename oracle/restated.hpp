// TEST INFRASTRUCTURE ONLY (oracle). Never linked into, imported by or executed from the product
// path (mcarray_b200/, include/); only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
//
// CPU float64 restatement of the mcarray hot path.  Everything that
// mcarray itself owns follows the cited reference lines; everything that DSPONE / WIPP own
// (FFT, framing, window, GCC-PHAT, mel bank, SignalPower, vector primitives) comes from
// oracle/standin/ and is PARITY UNPINNED (oracle/CONVENTIONS.md).  The mcarray-owned part is
// pinned by `make ref`: the reference's own .cpp files, compiled where they lie against the same
// stand-in, must give identical numbers (tests/test_oracle_vs_ref.py).
//
// Layout notes: spectra are CCS buffers of N+2 doubles (K = N/2+1 complex bins); flat arrays are
// row-major with the sizes written next to each argument.
#ifndef ORACLE_RESTATED_HPP
#define ORACLE_RESTATED_HPP

#include <dspone/standin.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <vector>

namespace orc {

typedef std::complex<double> cd;

// ------------------------------------------------------------------------------------------------
// A3 helpers — microhponeArrayHelpers.cpp:38-72,110-120.  The float typing is part of the contract:
// the reference returns `float` from every one of these, so grids and delay tables are float-rounded.
// ------------------------------------------------------------------------------------------------
inline double speed_of_sound() { return 346.1; }                                  // :38-43

inline float doa_idx_to_angle(int idx, float doa_step) {                           // :117-120
  return float((static_cast<float>(idx) * doa_step) - M_PI_2);
}
inline float angle_to_doa_idx(float angle, float doa_step) {                       // :110-115
  angle = float(std::max(static_cast<double>(angle), -M_PI_2));
  angle = float(std::min(static_cast<double>(angle), M_PI_2));
  return float(static_cast<int>((angle + M_PI_2) / doa_step));
}
inline float doa_to_delay_far_field(float doa, float micro_dist) {                 // :46-67
  // `sin(float)` resolves to the float overload under <cmath>; float*float, then /double, then
  // the float return type rounds once more.
  float delay = float((micro_dist * std::sin(doa)) / speed_of_sound());
  return delay;
}
inline float doa_to_delay_far_field_samples(float doa, float micro_dist, int fs) { // :69-72
  return doa_to_delay_far_field(doa, micro_dist) * float(fs);
}
inline int num_doa_steps(float doa_step) { return int(std::round(M_PI / doa_step) + 1); }   // SteeringBeamforming.cpp:40

// ArrayDescription::distance — ArrayDescription.cpp:57-64
inline double mic_distance(const double *xyz, int i, int j) {
  return std::sqrt(std::pow(xyz[3 * j] - xyz[3 * i], 2) + std::pow(xyz[3 * j + 1] - xyz[3 * i + 1], 2) +
                   std::pow(xyz[3 * j + 2] - xyz[3 * i + 2], 2));
}
// ArrayDescription::maxDistance — ArrayDescription.cpp:80-94
inline double max_mic_distance(const double *xyz, int M) {
  double best = 0;
  for (int i = 0; i < M; ++i) for (int j = 0; j < M; ++j) if (i != j) best = std::max(best, mic_distance(xyz, i, j));
  return best;
}

// A3 — SteeringBeamforming::generateLookupTable, SteeringBeamforming.cpp:58-94.
// pair order i<j lexicographic; tau[p][d] in samples (double holding a float-rounded value).
inline void reference_pair_delays(const double *xyz, int M, int fs, float doa_step, std::vector<double> &tau, int &D) {
  D = num_doa_steps(doa_step);
  tau.clear();
  for (int i = 0; i < M; ++i)
    for (int j = i + 1; j < M; ++j) {
      double distance = mic_distance(xyz, i, j);
      for (int d = 0; d < D; ++d) tau.push_back(doa_to_delay_far_field_samples(doa_idx_to_angle(d, doa_step), float(distance), fs));
    }
}

// ------------------------------------------------------------------------------------------------
// A2 — GCC-PHAT on a tau grid (DSPONE; stand-in conventions C5).  Returns Re(corr[d]) only,
// which is all the reference consumes (SteeringBeamforming.cpp:122, BinauralLocalisation.cpp:444).
// ------------------------------------------------------------------------------------------------
inline void gcc_phat_tau(const double *x_ccs, const double *y_ccs, int K, const double *tau, int D, double *corr_re) {
  std::vector<cd> g;
  dsp::GeneralisedCrossCorrelation::phat(reinterpret_cast<const dsp::Complex *>(x_ccs), reinterpret_cast<const dsp::Complex *>(y_ccs), g, K);
  const double nfft = 2.0 * double(K - 1);
  for (int d = 0; d < D; ++d) {
    cd acc(0, 0);
    for (int k = 0; k < K; ++k) acc += g[size_t(k)] * std::polar(1.0, 2.0 * M_PI * double(k) * tau[d] / nfft);
    corr_re[d] = acc.real();
  }
}

// Integer-lag TDOA (BASELINE config 2): the tau-vector mode of A2 with tau = -L..L, then the
// first-maximum argmax (wipp::maxidx).  lag = argmax - L.
inline int gcc_phat_lags(const double *x_ccs, const double *y_ccs, int K, int max_lag, double *curve /*[2L+1]*/) {
  std::vector<double> tau(size_t(2 * max_lag + 1));
  for (int l = -max_lag; l <= max_lag; ++l) tau[size_t(l + max_lag)] = double(l);
  gcc_phat_tau(x_ccs, y_ccs, K, tau.data(), 2 * max_lag + 1, curve);
  double mx; size_t idx;
  wipp::maxidx(curve, size_t(2 * max_lag + 1), &mx, &idx);
  return int(idx) - max_lag;
}

// ------------------------------------------------------------------------------------------------
// A4-A6 — SteeringBeamforming::{computeCorrelations, computeEnergyInDOA, selectDOA},
// SteeringBeamforming.cpp:104-195.
// ------------------------------------------------------------------------------------------------
struct SteeringState {
  int M, P, D, K;
  float doa_step;
  std::vector<double> tau;          // [P][D]
  std::vector<double> prev_energy;  // [D], starts at 0 (SteeringBeamforming.cpp:51)
  std::vector<std::array<int, 2>> pairs;
  void init(int M_, int K_, int D_, float step, const std::vector<double> &tau_) {
    M = M_; K = K_; D = D_; doa_step = step; tau = tau_; P = M * (M - 1) / 2;
    prev_energy.assign(size_t(D), 0.0);
    pairs.clear();
    for (int i = 0; i < M; ++i) for (int j = i + 1; j < M; ++j) pairs.push_back({i, j});
  }
};

// frames: [M][N+2].  corr: [P][D] (raw Re of GCC, before the in-place 0.2 scaling).
inline void steering_correlations(const SteeringState &st, const double *frames, double *corr) {   // :104-130
  const int ccs = 2 * st.K;
  for (int p = 0; p < st.P; ++p)
    gcc_phat_tau(frames + size_t(st.pairs[size_t(p)][0]) * ccs, frames + size_t(st.pairs[size_t(p)][1]) * ccs, st.K,
                 &st.tau[size_t(p) * size_t(st.D)], st.D, corr + size_t(p) * st.D);
}

// energy (out): smoothed, un-normalised E_t [D]; state updated.
inline void steering_energy(SteeringState &st, const double *corr, double *energy) {               // :132-144
  const float mem = 0.8f;                                  // SteeringBeamforming.h:70 (constexpr float)
  const double a = double(mem), b = double(1 - mem);       // multC takes a double
  for (int d = 0; d < st.D; ++d) energy[d] = a * st.prev_energy[size_t(d)];
  for (int p = 0; p < st.P; ++p)
    for (int d = 0; d < st.D; ++d) energy[d] += b * corr[size_t(p) * st.D + d];
  for (int d = 0; d < st.D; ++d) st.prev_energy[size_t(d)] = energy[d];
}

// selectDOA :146-195.  `energy` is consumed by value (the reference normalises in place).
// idx_out[s] = maxIdx+1 (the grid cell), doa_out[s] = cell angle (rad, float-rounded), prob_out[s] = peak weight.
inline void steering_select(const SteeringState &st, const double *energy, int S, int *idx_out, double *doa_out, double *prob_out) {
  const int D = st.D;
  std::vector<double> e(energy, energy + D), f(size_t(D - 1)), ff(size_t(D - 1)), s2(size_t(D - 2));
  const double min_e = -15.0 * st.P;                                                       // :152
  for (int d = 0; d < D; ++d) e[size_t(d)] = (e[size_t(d)] - min_e) / (-2.0 * min_e);      // :155-156
  for (int i = 0; i < D - 1; ++i) f[size_t(i)] = e[size_t(i) + 1] - e[size_t(i)];          // :159
  for (int i = 0; i < D - 1; ++i) { if (f[size_t(i)] < 0.0) f[size_t(i)] = 1.0; else if (f[size_t(i)] > 0.0) f[size_t(i)] = 0.0; }  // :161
  wipp::median_filter(f.data(), ff.data(), size_t(D - 1), 3);                              // :164
  for (int i = 0; i < D - 2; ++i) s2[size_t(i)] = (ff[size_t(i) + 1] - ff[size_t(i)]) * e[size_t(i) + 1];  // :170,173
  for (int s = 0; s < S; ++s) {                                                            // :185-194
    double mx; size_t mi;
    wipp::maxidx(s2.data(), size_t(D - 2), &mx, &mi);
    s2[mi] = 0;
    idx_out[s] = int(mi) + 1;
    doa_out[s] = doa_idx_to_angle(int(mi) + 1, st.doa_step);
    prob_out[s] = mx;
  }
}

// ------------------------------------------------------------------------------------------------
// A7 — Beamformer::processFrame, Beamformer.cpp:51-71.  frames [M][N+2] -> out [N+2].
// ------------------------------------------------------------------------------------------------
inline void ds_beamform(const double *frames, int M, int ccs_len, int fs, const double *mic_x, double doa, double *out) {
  const int K = ccs_len / 2;
  for (int i = 0; i < ccs_len; ++i) out[i] = 0.0;
  for (int c = 0; c < M; ++c) {
    const double slope = 2 * M_PI * fs / (ccs_len - 2) / speed_of_sound() * mic_x[c] * std::cos(doa + M_PI / 2);   // :59
    const double *x = frames + size_t(c) * ccs_len;
    for (int k = 0; k < K; ++k) {
      const double ph = 0.0 + slope * double(k);
      const double cr = std::cos(ph), ci = std::sin(ph);
      out[2 * k] += x[2 * k] * cr - x[2 * k + 1] * ci;
      out[2 * k + 1] += x[2 * k] * ci + x[2 * k + 1] * cr;
    }
  }
  for (int i = 0; i < ccs_len; ++i) out[i] /= double(M);                                    // :70
}

// ------------------------------------------------------------------------------------------------
// A8 — BeamformingSeparationAndLocalisation, BeamformingSeparationAndLocalisation.cpp:29-119, plus
// the power-floor state of SoundLocalisationImpl (SoundLocalisationImpl.cpp:29-41, .h:77-86).
// ------------------------------------------------------------------------------------------------
struct SslState {
  SteeringState steer;
  int fs, ccs_len, S;
  bool use_floor;
  std::vector<double> mic_x;
  std::vector<double> cur_doa, prob;     // init 0 / -1 (:46-52)
  double power_floor; bool noise_estimated; int samples_for_noise;
  void init(int fs_, const double *xyz, int M, int N, int S_, bool use_floor_) {
    fs = fs_; ccs_len = N + 2; S = S_; use_floor = use_floor_;
    float step = float(5 * M_PI / 180);                     // SteeringBeamforming.cpp:39 into `const float`
    std::vector<double> tau; int D;
    reference_pair_delays(xyz, M, fs, step, tau, D);
    steer.init(M, ccs_len / 2, D, step, tau);
    mic_x.resize(size_t(M));
    for (int m = 0; m < M; ++m) mic_x[size_t(m)] = xyz[3 * m];
    cur_doa.assign(size_t(S), 0.0); prob.assign(size_t(S), -1.0);
    power_floor = 0; noise_estimated = false; samples_for_noise = 0;
  }
};

struct FrameReport { bool fired; double power; };

// frames [M][N+2] in; returns whether the callback fires.  energy_out [D] (raw smoothed E_t) and
// idx_out [S] are filled when fired.
inline FrameReport ssl_localise(SslState &st, const double *frames, double *energy_out, int *idx_out, double *corr_out /*[P][D] or null*/) {
  const int M = st.steer.M;
  std::vector<const double *> fv;
  for (int c = 0; c < M; ++c) fv.push_back(frames + size_t(c) * st.ccs_len);
  double power;
  if (!st.noise_estimated && st.use_floor) {                                               // :82-83 -> :55-72
    const int needed = int(3 * st.fs);                                                     // _durationToEstimatePowerFloor = 3
    double p = dsp::SignalPower::FFTPower(fv, st.ccs_len) * (st.ccs_len - 2);
    st.power_floor += p;
    st.samples_for_noise += (st.ccs_len - 2);
    if (st.samples_for_noise >= needed) {
      st.noise_estimated = true;
      st.power_floor /= st.samples_for_noise;
      st.power_floor = 10 * std::log10(st.power_floor) + 3.0;                              // _noiseMarginDB = 3 (.h:52)
    }
    power = st.power_floor;
  } else {
    power = dsp::SignalPower::FFTLogPower(fv, st.ccs_len);                                 // :85
  }
  FrameReport r; r.power = power; r.fired = false;
  if ((power > st.power_floor) || !st.use_floor) {                                         // :89
    std::vector<double> corr(size_t(st.steer.P) * st.steer.D), energy(size_t(st.steer.D));
    steering_correlations(st.steer, frames, corr.data());
    if (corr_out) std::copy(corr.begin(), corr.end(), corr_out);
    steering_energy(st.steer, corr.data(), energy.data());
    std::vector<int> idx(size_t(st.S));
    steering_select(st.steer, energy.data(), st.S, idx.data(), st.cur_doa.data(), st.prob.data());
    if (energy_out) std::copy(energy.begin(), energy.end(), energy_out);
    if (idx_out) std::copy(idx.begin(), idx.end(), idx_out);
    r.fired = true;
  }
  return r;
}

// :103-119 — frames [M][N+2] are overwritten: channels < min(M,S) get the beams, the rest zeros.
inline void ssl_separate(const SslState &st, double *frames) {
  const int M = st.steer.M;
  std::vector<double> copy(frames, frames + size_t(M) * st.ccs_len);
  int c = 0;
  for (; c < std::min(M, st.S); ++c) ds_beamform(copy.data(), M, st.ccs_len, st.fs, st.mic_x.data(), st.cur_doa[size_t(c)], frames + size_t(c) * st.ccs_len);
  for (; c < M; ++c) std::fill(frames + size_t(c) * st.ccs_len, frames + size_t(c + 1) * st.ccs_len, 0.0);
}

// ------------------------------------------------------------------------------------------------
// A9 — FreqGCCBinauralLocalisation, BinauralLocalisation.cpp:320-631 (deterministic part: the
// smoothed correlation curve and its argmax; the particle filter is out of scope).
// ------------------------------------------------------------------------------------------------
struct FreqGccState {
  int fs, K, D; float doa_step; bool use_floor;
  double mic_dist;
  std::vector<double> tau, prev_corr;
  float corr_mem;                         // 0 until the first voiced frame, then 0.8f (:323,523)
  float doa_mem;                          // _doaMemoryFactor: 0 (:324) -> _maxDoaMemoryFactor = 0.6f (.h:199) after a voiced frame (:524)
  double cur_doa, prob;                   // _currentDOA[0] = 0, _prob[0] = -1 (:338-339); deterministic tracker = the #else branch (:501-504)
  double power_floor; bool noise_estimated; int samples_for_noise; int silence_frames;
  void init(int fs_, double mic_dist_, int N, bool use_floor_) {
    fs = fs_; K = N / 2 + 1; use_floor = use_floor_; mic_dist = mic_dist_;
    doa_step = float(3 * M_PI / 180);                                                      // :328
    D = num_doa_steps(doa_step);                                                           // :329
    tau.resize(size_t(D));
    for (int i = 0; i < D; ++i) tau[size_t(i)] = doa_to_delay_far_field_samples(doa_idx_to_angle(i, doa_step), float(mic_dist), fs);  // :363-366
    prev_corr.assign(size_t(D), 0.0); corr_mem = 0; doa_mem = 0; cur_doa = 0; prob = -1;
    power_floor = 0; noise_estimated = false; samples_for_noise = 0; silence_frames = 0;
  }
};

inline void freqgcc_probability(const double *curve, int D, float doa_step, const double *doas, double *probs, int size);

// frames [2][N+2]; curve_out [D] = smoothed correlation; returns fired flag; *idx_out = argmax cell.
inline FrameReport freqgcc_frame(FreqGccState &st, const double *frames, double *curve_out, int *idx_out) {
  const int ccs = 2 * st.K;
  std::vector<const double *> fv{frames, frames + ccs};
  double power;
  if (!st.noise_estimated) {                                                               // :429-430 -> :387-404
    const int needed = int(3 * st.fs);
    double p = dsp::SignalPower::power(fv, ccs) * (ccs - 2);
    st.power_floor += p + 1e-10;
    st.samples_for_noise += (ccs - 2);
    if (st.samples_for_noise >= needed) {
      st.noise_estimated = true;
      if (st.samples_for_noise > 0) st.power_floor /= st.samples_for_noise;
      st.power_floor = 10 * std::log10(st.power_floor) + double(6.0f);                     // _noiseMarginDB = 6 (.h:197)
    }
    power = st.power_floor;
  } else {
    power = dsp::SignalPower::FFTLogPower(fv, ccs);                                        // :432
  }
  FrameReport r; r.power = power; r.fired = false;
  if (power > st.power_floor || !st.use_floor) {                                           // :434
    std::vector<double> c(size_t(st.D));
    gcc_phat_tau(frames, frames + ccs, st.K, st.tau.data(), st.D, c.data());               // :438-444
    const double keep = double(1 - st.corr_mem), mem = double(st.corr_mem);
    for (int d = 0; d < st.D; ++d) {                                                       // :445-448
      c[size_t(d)] *= keep;
      st.prev_corr[size_t(d)] *= mem;
      c[size_t(d)] += st.prev_corr[size_t(d)];
      st.prev_corr[size_t(d)] = c[size_t(d)];
    }
    freqgcc_probability(c.data(), st.D, st.doa_step, &st.cur_doa, &st.prob, 1);            // :454, with the PREVIOUS _currentDOA
    double mx; size_t mi;
    wipp::maxidx(c.data(), size_t(st.D), &mx, &mi);                                        // :459 / :502
    // deterministic tracker: the `#else` branch of USE_PARTICLE_FILTER (:501-504).  float * double and (1 - float) as written there.
    const double doa = doa_idx_to_angle(int(mi), st.doa_step);
    st.cur_doa = st.doa_mem * st.cur_doa + (1 - st.doa_mem) * doa;
    if (curve_out) std::copy(c.begin(), c.end(), curve_out);
    if (idx_out) *idx_out = int(mi);
    st.corr_mem = 0.8f;                                                                    // :523
    st.doa_mem = 0.6f;                                                                     // :524
    st.silence_frames = 0;
    r.fired = true;
  } else if (st.noise_estimated) {                                                         // :528-561
    const int windows_to_decay = 3 * st.fs / (ccs / 2 - 1);
    if (st.silence_frames < windows_to_decay) { st.corr_mem = 0.8f; st.doa_mem = 0.6f; }
    else { st.corr_mem = 0; st.doa_mem = 0; }
    ++st.silence_frames;
  }
  return r;
}

// setProbability — BinauralLocalisation.cpp:569-631, on a given smoothed curve.
inline void freqgcc_probability(const double *curve, int D, float doa_step, const double *doas, double *probs, int size) {
  double mn, sm;
  wipp::min(curve, size_t(D), &mn);
  wipp::sum(curve, size_t(D), &sm);
  sm -= mn * D;
  double prevcorr = 0, nextcorr = 0, prevdoa = 0, nextdoa = 0, p;
  for (int i = 0; i < size; ++i) {
    int idx = int(angle_to_doa_idx(float(doas[i]), doa_step));
    double angle = doa_idx_to_angle(idx, doa_step);
    if (0 < idx && idx < (D - 1)) {
      if (angle > doas[i] && idx > 0) {
        prevcorr = curve[idx - 1]; prevdoa = doa_idx_to_angle(idx - 1, doa_step); nextcorr = curve[idx]; nextdoa = angle;
      } else if (angle <= doas[i] && idx < (D - 1)) {
        prevcorr = curve[idx]; prevdoa = angle; nextcorr = curve[idx + 1]; nextdoa = doa_idx_to_angle(idx + 1, doa_step);
      }
      double slope = (nextcorr - prevcorr) / (nextdoa - prevdoa);
      p = slope * (doas[i] - prevdoa) + prevcorr;
    } else {
      p = curve[idx];
    }
    probs[i] = 0;
    if (sm > 0) probs[i] = (p - mn) / sm;
    probs[i] = (probs[i] < 0.01) ? 0 : probs[i];
  }
}

// ------------------------------------------------------------------------------------------------
// N2 — MultibandBinarualLocalisation, MultibandBinarualLocalisation.cpp:52-259: per linear sub-band a
// GCC-PHAT curve on the 5-degree grid with 0.4 memory and its arg-max, then an energy-weighted histogram of the band
// arg-maxima whose own arg-max is the published DOA.  The sub-band frames are X * H_b (stand-in C8).
// ------------------------------------------------------------------------------------------------
struct MultibandState {
  int fs, N, K, D, nb; float doa_step; bool use_floor;
  double mic_dist;
  std::vector<double> tau, H /*[nb][K]*/, prev_corr /*[nb][D]*/;
  double cur_doa, prob, power_floor; bool noise_estimated; int samples_for_noise;
  void init(int fs_, double mic_dist_, int N_, int nbins, bool use_floor_) {
    fs = fs_; N = N_; K = N / 2 + 1; nb = nbins; use_floor = use_floor_; mic_dist = mic_dist_;
    doa_step = float(5 * M_PI / 180);                                                      // :62
    D = int(std::floor(M_PI / doa_step) + 1);                                              // :63 (floor here, round in SteeringBeamforming)
    tau.resize(size_t(D));
    for (int i = 0; i < D; ++i) tau[size_t(i)] = doa_to_delay_far_field_samples(doa_idx_to_angle(i, doa_step), float(mic_dist), fs);   // :99
    int order = 0; while ((1 << order) < N) ++order;
    const float max_freq = float(speed_of_sound() / (2 * float(mic_dist)));                // maxFreqForSpatialAliasing, microhponeArrayHelpers.cpp:85-89
    dsp::FilterBankFFTWLinear bank(order, nb, fs, 100, max_freq);                          // :54-60
    H.assign(bank.band(0), bank.band(0) + size_t(nb) * K);
    prev_corr.assign(size_t(nb) * D, 0.0);
    cur_doa = 0; prob = -1; power_floor = 0; noise_estimated = false; samples_for_noise = 0;   // :76-79, SoundLocalisationImpl
  }
};

// frames [2][N+2].  hist_out [D] = _energyInDOA, band_cells_out [nb], *cell_out = arg-max of the histogram.
inline FrameReport multiband_frame(MultibandState &st, const double *frames, double *hist_out, int *band_cells_out, int *cell_out) {
  const int ccs = 2 * st.K;
  std::vector<double> hist(size_t(st.D), 0.0), band(size_t(2) * ccs), c(size_t(st.D));    // processSetup :145-151
  const float mem = 0.4f;                                                                  // _corrMemoryFactor, .h:44
  for (int b = 0; b < st.nb; ++b) {                                                        // processOneSubband :164-196
    const double *h = &st.H[size_t(b) * st.K];
    for (int ch = 0; ch < 2; ++ch)
      for (int i = 0; i < ccs; ++i) band[size_t(ch) * ccs + i] = frames[size_t(ch) * ccs + i] * h[i / 2];
    gcc_phat_tau(band.data(), band.data() + ccs, st.K, st.tau.data(), st.D, c.data());    // :175-179
    double *prev = &st.prev_corr[size_t(b) * st.D];
    for (int d = 0; d < st.D; ++d) {                                                       // :180-183
      c[size_t(d)] *= double(1 - mem);
      prev[d] *= double(mem);
      c[size_t(d)] += prev[d];
      prev[d] = c[size_t(d)];
    }
    double mx; size_t mi;
    wipp::maxidx(c.data(), size_t(st.D), &mx, &mi);                                        // :184
    std::vector<const double *> bv{band.data(), band.data() + ccs};
    hist[mi] += dsp::SignalPower::FFTPower(bv, ccs);                                       // :188-190
    if (band_cells_out) band_cells_out[b] = int(mi);
  }
  std::vector<const double *> fv{frames, frames + ccs};
  double power;
  if (!st.noise_estimated) {                                                               // :214-217 -> :127-143
    const int K = ccs / 2, needed = int(3 * st.fs);
    st.power_floor += dsp::SignalPower::FFTPower(fv, K) * (2 * K - 2);
    st.samples_for_noise += (2 * K - 2);
    if (st.samples_for_noise >= needed) {
      st.noise_estimated = true;
      st.power_floor /= st.samples_for_noise;
      st.power_floor = 10 * std::log10(st.power_floor) + double(3.0f);
    }
    power = st.power_floor;
  } else {
    power = dsp::SignalPower::FFTPower(fv, ccs);                                           // :221
  }
  FrameReport r; r.power = power; r.fired = false;
  if (power > st.power_floor || !st.use_floor) {                                           // :225
    double sum = 0, mx; size_t mi;
    wipp::sum(hist.data(), size_t(st.D), &sum);
    wipp::maxidx(hist.data(), size_t(st.D), &mx, &mi);
    st.prob = sum;
    if (st.prob != 0) st.prob = hist[mi] / st.prob;                                        // :230-233
    const double doa = doa_idx_to_angle(int(mi), st.doa_step);
    st.cur_doa = double(0.0f) * st.cur_doa + double(1 - 0.0f) * doa;                       // _doaMemoryFactor = 0 (:239)
    if (cell_out) *cell_out = int(mi);
    r.fired = true;
  } else {
    st.cur_doa = st.cur_doa * double(1.0f) + double(1 - 1.0f) * 0;                         // :252
    st.prob = -100000;
  }
  if (hist_out) std::copy(hist.begin(), hist.end(), hist_out);
  return r;
}

// ------------------------------------------------------------------------------------------------
// A10 — FastBinauralMasking, FastBinauralMasking.cpp:51-538; constants FastBinauralMasking.h:111-128.
// ------------------------------------------------------------------------------------------------
enum MaskMethod { M_FACTOR = 0, M_RELATIVE = 1, M_FULL = 3, M_NOISY = 4, M_NOTHING = 5 };   // ArrayModules.h:81
enum MaskAlg { A_BOTH = 0, A_SPATIAL = 1, A_TEMPORAL = 2 };                                  // ArrayModules.h:89

struct MaskState {
  int fs, N, K, n_bands, method, alg, first_call;
  double mic_dist;
  std::vector<double> H;            // [n_bands][K] real filter magnitudes
  std::vector<double> thresholds;   // [n_bands]
  std::vector<double> Q, noise_est; // [n_bands]
  void init(int fs_, double mic_dist_, int N_, int method_, int alg_, int n_bands_, const double *H_, const double *band_freq_norm) {
    fs = fs_; N = N_; K = N / 2 + 1; n_bands = n_bands_; method = method_; alg = alg_; mic_dist = mic_dist_; first_call = 0;
    H.assign(H_, H_ + size_t(n_bands) * K);
    thresholds.resize(size_t(n_bands));
    const double phi = 10 * M_PI / 180;                                                    // .h:113
    for (int b = 0; b < n_bands; ++b) {                                                    // :342-366
      double w = band_freq_norm[b] * fs * 2 * M_PI;
      thresholds[size_t(b)] = std::cos(w * mic_dist * std::sin(phi) / speed_of_sound());
    }
    Q.assign(size_t(n_bands), 0.0); noise_est.assign(size_t(n_bands), 0.0);
  }
};

namespace maskdetail {
// getPower :520-538 — `length` doubles = length/2 complex bins
inline double get_power(const double *frame, int length) {
  const int n = length / 2; double acc = 0;
  for (int k = 0; k < n; ++k) { double m = std::sqrt(frame[2 * k] * frame[2 * k] + frame[2 * k + 1] * frame[2 * k + 1]); acc += m * m; }
  return std::sqrt(acc / n);
}
// getFramePower :496-516
inline double frame_power(const double *l, const double *r, int length) {
  std::vector<double> mixed(static_cast<size_t>(length));
  for (int i = 0; i < length; ++i) mixed[size_t(i)] = l[i] / 2 + r[i] / 2;
  return get_power(mixed.data(), length);
}
// normaliseFFTCorrelation :410-460
inline double norm_fft_corr(const double *l, const double *r, int K) {
  double num = 0, el = 0, er = 0;
  for (int k = 0; k < K; ++k) num += r[2 * k] * l[2 * k] + r[2 * k + 1] * l[2 * k + 1];   // Re(R * conj(L))
  num /= K;
  if (num == 0) return 0;
  for (int k = 0; k < K; ++k) { double m = std::sqrt(l[2 * k] * l[2 * k] + l[2 * k + 1] * l[2 * k + 1]); el += m * m; }
  for (int k = 0; k < K; ++k) { double m = std::sqrt(r[2 * k] * r[2 * k] + r[2 * k + 1] * r[2 * k + 1]); er += m * m; }
  double den = std::sqrt((el / K) * (er / K));
  return den == 0 ? 1 : num / den;
}
}  // namespace maskdetail

// frames [2][N+2], overwritten with the masked spectra.  decisions_out [n_bands]: 2 = spatial mask,
// 1 = temporal mask, 0 = pass (optional).
inline void mask_frame(MaskState &st, double *frames, int *decisions_out) {
  using namespace maskdetail;
  if (st.method == M_NOTHING) return;                                                      // :130-134
  const int ccs = st.N + 2, K = st.K, W = st.N;                                            // W = _windowSize doubles
  std::vector<double> outL(size_t(ccs), 0.0), outR(size_t(ccs), 0.0), L(static_cast<size_t>(ccs)), R(static_cast<size_t>(ccs));
  double *xl = frames, *xr = frames + ccs;
  for (int b = 0; b < st.n_bands; ++b) {                                                   // :146-191
    const double *h = &st.H[size_t(b) * K];
    for (int k = 0; k < K; ++k) { L[2 * k] = xl[2 * k] * h[k]; L[2 * k + 1] = xl[2 * k + 1] * h[k]; R[2 * k] = xr[2 * k] * h[k]; R[2 * k + 1] = xr[2 * k + 1] * h[k]; }
    // temportalMasking :477-493
    const float lam = 0.04f, reject = 0.999f;
    double pw = frame_power(L.data(), R.data(), W);
    st.Q[size_t(b)] = st.Q[size_t(b)] * lam + (1 - lam) * pw;
    bool temp = pw < reject * st.Q[size_t(b)];
    bool spat = false;
    if (st.alg == A_BOTH || st.alg == A_SPATIAL) {                                         // :159-166
      spat = norm_fft_corr(L.data(), R.data(), K) < st.thresholds[size_t(b)];             // :369-376
      if (st.alg == A_SPATIAL) temp = false;
    }
    int decision = spat ? 2 : (temp ? 1 : 0);
    if (decisions_out) decisions_out[b] = decision;
    if (decision) {                                                                        // :168-182 -> maskFrame :289-309
      const float factor = spat ? 10.0f : 3.0f;
      double *ch[2] = {L.data(), R.data()};
      for (int c = 0; c < 2; ++c) {
        double *f = ch[c];
        switch (st.method) {
          case M_FULL: for (int i = 0; i < W; ++i) f[i] /= 1000; break;                    // :214-217
          case M_RELATIVE: {                                                               // :246-282
            double mp = 0; for (int k = 0; k < K; ++k) { double m = std::sqrt(f[2 * k] * f[2 * k] + f[2 * k + 1] * f[2 * k + 1]); mp += m * m; }
            double fac = (mp / K) * 0.01f;
            if (st.Q[size_t(b)] < 1e-10) fac = 0.01f; else fac /= st.Q[size_t(b)];
            fac = std::sqrt(fac);
            for (int i = 0; i < W; ++i) f[i] *= fac;
          } break;
          case M_FACTOR: for (int i = 0; i < W; ++i) f[i] /= factor; break;                // :284-287
          case M_NOISY: {                                                                  // :220-243
            double pb = get_power(f, W), fac = 1;
            if (pb > 0) fac = st.noise_est[size_t(b)] / pb;
            if (st.first_call >= 2) for (int i = 0; i < W; ++i) f[i] *= fac;
          } break;
          default: break;
        }
      }
    } else {
      const float enh = 1;                                                                 // :184-187, _enhanceFactor = 1
      for (int i = 0; i < W; ++i) { L[size_t(i)] *= enh; R[size_t(i)] *= enh; }
    }
    for (int i = 0; i < ccs; ++i) { outL[size_t(i)] += L[size_t(i)]; outR[size_t(i)] += R[size_t(i)]; }   // :189-190
  }
  ++st.first_call;                                                                         // :193-197
  if (st.first_call < 2) st.noise_est = st.Q;
  std::copy(outL.begin(), outL.end(), xl);                                                 // :199-200
  std::copy(outR.begin(), outR.end(), xr);
}

// ------------------------------------------------------------------------------------------------
// A1 / A11 — framing, window, FFT / IFFT, overlap-add: thin processors over the DSPONE stand-in so
// that the restated per-frame functions above see exactly the frames the reference classes would.
// ------------------------------------------------------------------------------------------------
class FrameTap : public dsp::STFT {
 public:
  typedef void (*Hook)(void *user, double *frames /*[M][N+2] contiguous*/, int M, int ccs);
  FrameTap(int M, int order, bool synth, Hook hook, void *user) : dsp::STFT(M, order), _hook(hook), _user(user), _synth(synth) {
    if (!synth) _mode = ANALYSIS;
  }
 protected:
  virtual void processParametrisation(std::vector<double *> &af, int analysisLength, std::vector<double *> &, int) {
    const int M = int(af.size());
    _flat.resize(size_t(M) * analysisLength);
    for (int c = 0; c < M; ++c) std::copy(af[size_t(c)], af[size_t(c)] + analysisLength, &_flat[size_t(c) * analysisLength]);
    _hook(_user, _flat.data(), M, analysisLength);
    for (int c = 0; c < M; ++c) std::copy(&_flat[size_t(c) * analysisLength], &_flat[size_t(c + 1) * analysisLength], af[size_t(c)]);
  }
 private:
  Hook _hook; void *_user; bool _synth; std::vector<double> _flat;
};

}  // namespace orc

#endif

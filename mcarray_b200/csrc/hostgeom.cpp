// Host-side constructor math of the reference processors (no GPU work): delay tables, steering slopes, frame-size rule,
// mel filter bank.  Float typing follows the reference helpers exactly, because their float rounding is part of the
// contract: src/mcarray/microhponeArrayHelpers.cpp:38-72,110-120,
// SteeringBeamforming.cpp:34-94, Beamformer.cpp:59, BinauralLocalisation.cpp:328-366, FastBinauralMasking.cpp:95-98,342-366.
#include "../../include/mcarray_b200.h"

#include <cmath>
#include <vector>

namespace {
const double kSpeedOfSound = 346.1;                                                    // microhponeArrayHelpers.cpp:38-43
inline float doa_idx_to_angle(int idx, float step) { return float((static_cast<float>(idx) * step) - M_PI_2); }   // :117-120
inline float delay_samples(float doa, float dist, int fs) {                            // :46-72
  float delay = float((dist * std::sin(doa)) / kSpeedOfSound);
  return delay * float(fs);
}
inline double dist3(const double *xyz, int i, int j) {                                 // ArrayDescription.cpp:57-64
  return std::sqrt(std::pow(xyz[3 * j] - xyz[3 * i], 2) + std::pow(xyz[3 * j + 1] - xyz[3 * i + 1], 2) + std::pow(xyz[3 * j + 2] - xyz[3 * i + 2], 2));
}
}  // namespace

extern "C" {

/* N = 2^calculateOrderFromSampleRate(fs, frameRate): smallest power of two >= frameRate*fs (oracle/CONVENTIONS.md C1) */
int mcag_geom_frame_size(int fs, double frame_rate) {
  double want = frame_rate * double(fs);
  int order = 0;
  while (double(1 << order) < want) ++order;
  return 1 << order;
}

/* number of grid cells for a step: round(pi/step)+1 (SteeringBeamforming.cpp:40, BinauralLocalisation.cpp:329) */
int mcag_geom_grid_size(float doa_step) { return int(std::round(M_PI / doa_step) + 1); }

double mcag_geom_cell_angle(int idx, float doa_step) { return doa_idx_to_angle(idx, doa_step); }

/* SteeringBeamforming::generateLookupTable: tau [P][D] samples, pairs i<j lexicographic, scalar pair distance */
int mcag_geom_pair_tau_reference(const double *mic_xyz, int M, int fs, float doa_step, double *tau) {
  const int D = mcag_geom_grid_size(doa_step);
  int p = 0;
  for (int i = 0; i < M; ++i)
    for (int j = i + 1; j < M; ++j, ++p) {
      const double distance = dist3(mic_xyz, i, j);
      for (int d = 0; d < D; ++d) tau[(long long)p * D + d] = delay_samples(doa_idx_to_angle(d, doa_step), float(distance), fs);
    }
  return D;
}

/* Beamformer::processFrame slope / (2 pi): turns per bin for every grid cell, plus row D for the initial DOA of 0 rad.
 * turns [(D+1)][M]; uses the x coordinate only, like Beamformer.cpp:59. */
int mcag_geom_steer_turns_reference(const double *mic_xyz, int M, int fs, int N, float doa_step, double *turns) {
  const int D = mcag_geom_grid_size(doa_step);
  for (int d = 0; d <= D; ++d) {
    const double doa = (d < D) ? double(doa_idx_to_angle(d, doa_step)) : 0.0;
    for (int c = 0; c < M; ++c) turns[(long long)d * M + c] = double(fs) / double(N) / kSpeedOfSound * mic_xyz[3 * c] * std::cos(doa + M_PI / 2);
  }
  return D;
}
/* the same for arbitrary steering angles (Beamformer::processFrame(frames, out, DOA) with a free DOA; BASELINE config 3) */
void mcag_geom_steer_turns(const double *mic_xyz, int M, int fs, int N, const double *doas, int D, double *turns) {
  for (int d = 0; d < D; ++d)
    for (int c = 0; c < M; ++c) turns[(long long)d * M + c] = double(fs) / double(N) / kSpeedOfSound * mic_xyz[3 * c] * std::cos(doas[d] + M_PI / 2);
}

/* generalised far-field geometry: per-microphone advance (p_m . u_d)/c*fs for unit vectors dirs [D][3]; mic_tau [M][D] */
void mcag_geom_mic_tau(const double *mic_xyz, int M, int fs, const double *dirs, int D, double *mic_tau) {
  for (int m = 0; m < M; ++m)
    for (int d = 0; d < D; ++d)
      mic_tau[(long long)m * D + d] = (mic_xyz[3 * m] * dirs[3 * d] + mic_xyz[3 * m + 1] * dirs[3 * d + 1] + mic_xyz[3 * m + 2] * dirs[3 * d + 2]) / kSpeedOfSound * fs;
}
void mcag_geom_pair_tau_from_mic_tau(const double *mic_tau, int M, int D, double *pair_tau) {
  int p = 0;
  for (int i = 0; i < M; ++i)
    for (int j = i + 1; j < M; ++j, ++p)
      for (int d = 0; d < D; ++d) pair_tau[(long long)p * D + d] = mic_tau[(long long)j * D + d] - mic_tau[(long long)i * D + d];
}

/* mel filter bank in the FFT domain (DSPONE FilterBankFFTWMelScale stand-in, oracle/CONVENTIONS.md C7) and the spatial
 * thresholds of FastBinauralMasking::calculateThresholds (FastBinauralMasking.cpp:342-366, phi = 10 degrees). */
void mcag_geom_mel_bank(int N, int n_bands, int fs, float lo, float hi, double mic_dist, double *H, double *fc_norm, double *thresholds) {
  const int K = N / 2 + 1;
  auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
  auto mel2hz = [](double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); };
  const double mlo = hz2mel(lo), mhi = hz2mel(hi);
  std::vector<double> edges(size_t(n_bands) + 2);
  for (int i = 0; i < n_bands + 2; ++i) edges[size_t(i)] = mel2hz(mlo + (mhi - mlo) * double(i) / double(n_bands + 1));
  const double phi = 10 * M_PI / 180;
  for (int b = 0; b < n_bands; ++b) {
    const double l = edges[size_t(b)], mid = edges[size_t(b) + 1], h = edges[size_t(b) + 2];
    if (fc_norm) fc_norm[b] = mid / double(fs);
    if (thresholds) thresholds[b] = std::cos((mid / double(fs)) * fs * 2 * M_PI * mic_dist * std::sin(phi) / kSpeedOfSound);
    for (int k = 0; k < K; ++k) {
      const double f = double(k) * double(fs) / double(N);
      double v = 0.0;
      if (f > l && f <= mid) v = (f - l) / (mid - l);
      else if (f > mid && f < h) v = (h - f) / (h - mid);
      H[(long long)b * K + k] = v;
    }
  }
}

/* MultibandBinarualLocalisation.cpp:52-101; the linear bank is the DSPONE SubBandSTFTAnalysis stand-in (oracle/CONVENTIONS.md C8):
 * triangular unit-peak responses whose centres are equally spaced in Hz between 100 Hz and maxFreqForSpatialAliasing. */
int mcag_geom_multiband(int fs, double mic_dist, int N, int n_bands, double *tau, double *H) {
  const float step = float(5 * M_PI / 180);
  const int D = int(std::floor(M_PI / step) + 1);                                      // :63
  if (tau)
    for (int d = 0; d < D; ++d) tau[d] = delay_samples(doa_idx_to_angle(d, step), float(mic_dist), fs);   // :99
  if (H) {
    const int K = N / 2 + 1;
    const float lo_f = 100, hi_f = float(kSpeedOfSound / (2 * float(mic_dist)));        // microhponeArrayHelpers.cpp:85-89
    std::vector<double> edges(size_t(n_bands) + 2);
    for (int i = 0; i < n_bands + 2; ++i) edges[size_t(i)] = double(lo_f) + (double(hi_f) - double(lo_f)) * double(i) / double(n_bands + 1);
    for (int b = 0; b < n_bands; ++b) {
      const double l = edges[size_t(b)], mid = edges[size_t(b) + 1], h = edges[size_t(b) + 2];
      for (int k = 0; k < K; ++k) {
        const double f = double(k) * double(fs) / double(N);
        double v = 0.0;
        if (f > l && f <= mid) v = (f - l) / (mid - l);
        else if (f > mid && f < h) v = (h - f) / (h - mid);
        H[(long long)b * K + k] = v;
      }
    }
  }
  return D;
}

}  // extern "C"

// TEST INFRASTRUCTURE ONLY (oracle).  extern "C" drivers around the REFERENCE'S OWN classes.
// Built only where /root/reference exists (`make -C oracle ref`): the reference .cpp files are
// compiled from where they lie against oracle/standin/ (DSPONE / WIPP / Boost are absent from the
// image) into oracle/_ref/libmcarray_ref.so.  No reference source is copied into this repository.
// Compiled with -Dprivate=public -Dprotected=public so the drivers can read per-frame internals
// (_prevEnergyInDOA, _correlations, _correlationsReal, _shortTimePower) without touching the sources.
#define ORC_PREFIX ref_
#include "capi.h"

#include <mcarray/ArrayDescription.h>
#include <mcarray/Beamformer.h>
#include <mcarray/BinauralLocalisation.h>
#include <mcarray/FastBinauralMasking.h>
#include <mcarray/MultibandBinarualLocalisation.h>
#include <mcarray/SoundLocalisationCallback.h>
#include <mcarray/SourceLocalisation.h>
#include <mcarray/SourceSeparationAndLocalisation.h>
#include <mcarray/SteeringBeamforming.h>
#include <mcarray/microhponeArrayHelpers.h>

#include <algorithm>
#include <vector>

using namespace mca;

namespace {

ArrayDescription make_array(const double *xyz, int M) {
  ArrayDescription a;
  for (int m = 0; m < M; ++m) a.pushPosition(xyz[3 * m], xyz[3 * m + 1], xyz[3 * m + 2]);
  return a;
}

template <class Proc>
int feed(Proc &proc, int M, const double *in, int n, int chunk, double *out, int out_cap, bool synth) {
  int written = 0;
  if (chunk <= 0) chunk = n;
  const int ocap = chunk + proc.getMaxLatency();
  std::vector<double> obuf(size_t(M) * size_t(ocap));
  for (int pos = 0; pos < n; pos += chunk) {
    const int len = std::min(chunk, n - pos);
    std::vector<double *> pin, pout;
    for (int c = 0; c < M; ++c) { pin.push_back(const_cast<double *>(in) + size_t(c) * n + pos); pout.push_back(&obuf[size_t(c) * size_t(ocap)]); }
    if (synth) {
      int got = proc.process(pin, len, pout, ocap);
      if (written + got > out_cap) return -1;
      for (int c = 0; c < M; ++c) std::copy(pout[size_t(c)], pout[size_t(c)] + got, out + size_t(c) * out_cap + written);
      written += got;
    } else {
      proc.process(pin, len);
    }
  }
  return written;
}

// frame counters: wrap the (now public) virtual hook of each processor
struct CountingSSL : public SourceSeparationAndLocalisation {
  CountingSSL(int fs, ArrayDescription a, unsigned S, bool floor) : SourceSeparationAndLocalisation(fs, a, S, floor), frames(0) {}
  virtual void processParametrisation(std::vector<double *> &af, int al, std::vector<double *> &dc, int dl) {
    SourceSeparationAndLocalisation::processParametrisation(af, al, dc, dl);
    ++frames;
  }
  int frames;
};
struct CountingSL : public SourceLocalisation {
  CountingSL(int fs, ArrayDescription a, unsigned S, bool floor) : SourceLocalisation(fs, a, S, floor), frames(0) {}
  virtual void processParametrisation(std::vector<double *> &af, int al, std::vector<double *> &dc, int dl) {
    SourceLocalisation::processParametrisation(af, al, dc, dl);
    ++frames;
  }
  int frames;
};

struct SslCapture : public LocalisationCallback {
  BeamformingSeparationAndLocalisation *impl; const int *frame_counter;
  int fired, max_frames, S; int *fired_frame; double *doa_deg, *prob, *power, *energy, *corr_scaled;
  virtual void setDOA(SignalPtr doa, SignalPtr p, double pw, int n) {
    if (fired < max_frames) {
      SteeringBeamforming &sb = impl->_steeringBeamforming;
      const int D = sb._numSteps, P = int(sb._correlations.size());
      fired_frame[fired] = *frame_counter;
      for (int s = 0; s < n; ++s) { doa_deg[size_t(fired) * S + s] = doa[s]; prob[size_t(fired) * S + s] = p[s]; }
      power[fired] = pw;
      std::copy(sb._prevEnergyInDOA.get(), sb._prevEnergyInDOA.get() + D, energy + size_t(fired) * D);
      if (corr_scaled)
        for (int q = 0; q < P; ++q) std::copy(sb._correlations[size_t(q)].get(), sb._correlations[size_t(q)].get() + D, corr_scaled + (size_t(fired) * P + q) * D);
    }
    ++fired;
  }
};

}  // namespace

double ref_doa_idx_to_angle(int idx, float doa_step) { return doaIdx2angle(idx, doa_step); }
double ref_angle_to_doa_idx(float angle, float doa_step) { return angle2DOAidx(angle, doa_step); }
double ref_doa_to_delay_samples(float doa, float mic_dist, int fs) { return doaToDelayFarFieldSamples(doa, mic_dist, fs); }
double ref_array_distance(const double *xyz, int M, int i, int j) { return make_array(xyz, M).distance(i, j); }
double ref_array_max_distance(const double *xyz, int M) { return make_array(xyz, M).maxDistance(); }
int ref_frame_size(int fs, double frame_rate) { return 1 << dsp::ShortTimeProcess::calculateOrderFromSampleRate(fs, frame_rate); }

int ref_ssl_run(int fs, int M, const double *mic_xyz, int S, int use_floor, int analysis_only,
                const double *in, int n, int chunk, double *out, int out_cap, int *n_out,
                int max_frames, int *n_frames, int *n_fired, int *fired_frame,
                double *doa_deg, double *prob, double *power, double *energy, double *corr_scaled) {
  ArrayDescription a = make_array(mic_xyz, M);
  SslCapture cb;
  cb.fired = 0; cb.max_frames = max_frames; cb.S = S; cb.fired_frame = fired_frame; cb.doa_deg = doa_deg; cb.prob = prob;
  cb.power = power; cb.energy = energy; cb.corr_scaled = corr_scaled;
  int N, frames, w = 0;
  if (analysis_only) {
    CountingSL p(fs, a, unsigned(S), use_floor != 0);
    cb.impl = p._impl.get(); cb.frame_counter = &p.frames;
    p.setCallback(&cb);
    feed(p, M, in, n, chunk, nullptr, 0, false);
    N = p.getWindowSize(); frames = p.frames;
  } else {
    CountingSSL p(fs, a, unsigned(S), use_floor != 0);
    cb.impl = p._impl.get(); cb.frame_counter = &p.frames;
    p.setCallback(&cb);
    w = feed(p, M, in, n, chunk, out, out_cap, true);
    if (w < 0) return -1;
    N = p.getWindowSize(); frames = p.frames;
  }
  if (n_out) *n_out = w;
  if (n_frames) *n_frames = frames;
  if (n_fired) *n_fired = cb.fired;
  return N;
}

namespace {
struct CountingGcc : public FreqGCCBinauralLocalisation {
  CountingGcc(int fs, ArrayDescription a, bool floor) : FreqGCCBinauralLocalisation(fs, a, floor), frames(0) {}
  virtual void processParametrisation(std::vector<double *> &af, int al, std::vector<double *> &dc, int dl) {
    FreqGCCBinauralLocalisation::processParametrisation(af, al, dc, dl);
    ++frames;
  }
  int frames;
};
struct GccCapture : public LocalisationCallback {
  CountingGcc *proc; int fired, max_frames; int *fired_frame, *idx; double *curves, *power;
  virtual void setDOA(SignalPtr, SignalPtr, double pw, int) {
    if (fired < max_frames) {
      const int D = proc->_numSteps;
      fired_frame[fired] = proc->frames; power[fired] = pw;
      std::copy(proc->_correlationsReal.get(), proc->_correlationsReal.get() + D, curves + size_t(fired) * D);
      double mx; size_t mi;
      wipp::maxidx(proc->_correlationsReal.get(), size_t(D), &mx, &mi);       // :459
      idx[fired] = int(mi);
    }
    ++fired;
  }
};
}  // namespace

int ref_freqgcc_run(int fs, double mic_dist, int use_floor, int noise_preestimated, const double *in, int n, int chunk,
                    int max_frames, int *n_frames, int *n_fired, int *fired_frame, double *curves, int *idx, double *power) {
  std::vector<double> x = {0.0, mic_dist};
  ArrayDescription a = ArrayDescription::make_linear_array_description(x);
  CountingGcc p(fs, a, use_floor != 0);
  if (noise_preestimated) p._noiseEstimated = true;
  GccCapture cb; cb.proc = &p; cb.fired = 0; cb.max_frames = max_frames; cb.fired_frame = fired_frame; cb.idx = idx; cb.curves = curves; cb.power = power;
  p.setCallback(&cb);
  feed(p, 2, in, n, chunk, nullptr, 0, false);
  if (n_frames) *n_frames = p.frames;
  if (n_fired) *n_fired = cb.fired;
  return p.getWindowSize();
}

void ref_freqgcc_probability(int fs, double mic_dist, const double *curve, const double *doas, double *probs, int size) {
  std::vector<double> x = {0.0, mic_dist};
  FreqGCCBinauralLocalisation p(fs, ArrayDescription::make_linear_array_description(x), false);
  std::copy(curve, curve + p._numSteps, p._correlationsReal.get());
  p.setProbability(doas, probs, size);
}

namespace {
struct CountingMultiband : public MultibandBinarualLocalisation {
  CountingMultiband(int fs, ArrayDescription a, int nbins, bool floor) : MultibandBinarualLocalisation(fs, a, nbins, floor), frames(0) {}
  virtual void processParametrisation(std::vector<double *> &af, int al, std::vector<double *> &dc, int dl) {
    MultibandBinarualLocalisation::processParametrisation(af, al, dc, dl);
    ++frames;
  }
  int frames;
};
struct MultibandCapture : public LocalisationCallback {
  CountingMultiband *proc; int fired, max_frames; int *fired_frame, *cell, *band_cells; double *prob, *power, *doa_deg, *hist;
  virtual void setDOA(SignalPtr doa, SignalPtr p, double pw, int) {
    if (fired < max_frames) {
      const int D = proc->_numSteps, nb = proc->_numberOfBins;
      fired_frame[fired] = proc->frames; power[fired] = pw; prob[fired] = p[0]; doa_deg[fired] = doa[0];
      std::copy(proc->_energyInDOA.get(), proc->_energyInDOA.get() + D, hist + size_t(fired) * D);
      double mx; size_t mi;
      wipp::maxidx(proc->_energyInDOA.get(), size_t(D), &mx, &mi);                                   // :223
      cell[fired] = int(mi);
      for (int b = 0; b < nb; ++b)   // _binDOAs holds doaIdx2angle(idx) (:182): back to the cell
        band_cells[size_t(fired) * nb + b] = int(std::lround((proc->_binDOAs[b] + M_PI_2) / double(proc->_doaStep)));
    }
    ++fired;
  }
};
}  // namespace

int ref_multiband_run(int fs, double mic_dist, int nbins, int use_floor, int noise_preestimated, const double *in, int n, int chunk,
                      int max_frames, int *n_frames, int *n_fired, int *fired_frame, int *n_dirs,
                      int *cell, double *prob, double *power, double *doa_deg, double *hist, int *band_cells) {
  std::vector<double> x = {0.0, mic_dist};
  CountingMultiband p(fs, ArrayDescription::make_linear_array_description(x), nbins, use_floor != 0);
  if (n_dirs) *n_dirs = p._numSteps;
  if (!in) return p.getWindowSize();
  if (noise_preestimated) p._noiseEstimated = true;
  MultibandCapture cb; cb.proc = &p; cb.fired = 0; cb.max_frames = max_frames; cb.fired_frame = fired_frame; cb.cell = cell; cb.band_cells = band_cells;
  cb.prob = prob; cb.power = power; cb.doa_deg = doa_deg; cb.hist = hist;
  p.setCallback(&cb);
  feed(p, 2, in, n, chunk, nullptr, 0, false);
  if (n_frames) *n_frames = p.frames;
  if (n_fired) *n_fired = cb.fired;
  return p.getWindowSize();
}

namespace {
struct TracingMask : public FastBinauralMasking {
  TracingMask(int fs, double d, float lo, float hi, MaskingMethod m, MaskingAlg a) : FastBinauralMasking(fs, d, lo, hi, m, a), frames(0), max_frames(0), Q(nullptr), spectra(nullptr) {}
  virtual void processParametrisation(std::vector<double *> &af, int al, std::vector<double *> &dc, int dl) {
    FastBinauralMasking::processParametrisation(af, al, dc, dl);
    if (frames < max_frames) {
      if (Q) std::copy(_shortTimePower.get(), _shortTimePower.get() + _nBins, Q + size_t(frames) * _nBins);
      if (spectra) for (int c = 0; c < 2; ++c) std::copy(af[size_t(c)], af[size_t(c)] + al, spectra + (size_t(frames) * 2 + c) * al);
    }
    ++frames;
  }
  int frames, max_frames; double *Q, *spectra;
};
}  // namespace

int ref_mask_run(int fs, double mic_dist, float lo, float hi, int method, int alg,
                 const double *in, int n, int chunk, double *out, int out_cap, int *n_out,
                 int max_frames, int *n_frames, double *Q, double *spectra_out) {
  TracingMask p(fs, mic_dist, lo, hi, BinauralMasking::MaskingMethod(method), BinauralMasking::MaskingAlg(alg));
  p.max_frames = max_frames; p.Q = Q; p.spectra = spectra_out;
  int w = feed(p, 2, in, n, chunk, out, out_cap, true);
  if (w < 0) return -1;
  if (n_out) *n_out = w;
  if (n_frames) *n_frames = p.frames;
  return p.getWindowSize();
}

int ref_beamformer_frame(int fs, int M, const double *mic_xyz, int ccs_len, const double *frames, double doa, double *out) {
  Beamformer bf(fs, make_array(mic_xyz, M), ccs_len, unsigned(M));
  SignalVector fv;
  for (int m = 0; m < M; ++m) { fv.push_back(SignalPtr(new BaseType[ccs_len])); std::copy(frames + size_t(m) * ccs_len, frames + size_t(m + 1) * ccs_len, fv.back().get()); }
  SignalPtr o(new BaseType[ccs_len]);
  bf.processFrame(fv, o, doa);
  std::copy(o.get(), o.get() + ccs_len, out);
  return 0;
}

int ref_steering_frames(int fs, int M, const double *mic_xyz, int ccs_len, int S, const double *frames, int T,
                        double *doa_rad, double *prob, double *energy) {
  SteeringBeamforming sb(fs, make_array(mic_xyz, M), ccs_len, unsigned(M));
  SignalVector fv, wiener;
  for (int m = 0; m < M; ++m) fv.push_back(SignalPtr(new BaseType[ccs_len]));
  SignalPtr doa(new BaseType[S]), pr(new BaseType[S]);
  for (int t = 0; t < T; ++t) {
    for (int m = 0; m < M; ++m) std::copy(frames + (size_t(t) * M + m) * ccs_len, frames + (size_t(t) * M + m + 1) * ccs_len, fv[size_t(m)].get());
    sb.processFrame(fv, doa, pr, S, wiener);
    for (int s = 0; s < S; ++s) { doa_rad[size_t(t) * S + s] = doa[s]; prob[size_t(t) * S + s] = pr[s]; }
    std::copy(sb._prevEnergyInDOA.get(), sb._prevEnergyInDOA.get() + sb._numSteps, energy + size_t(t) * sb._numSteps);
  }
  return sb._numSteps;
}


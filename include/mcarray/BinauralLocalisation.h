/* mca::FreqGCCBinauralLocalisation — 2-microphone GCC-PHAT localiser on the 61-cell / 3 degree grid
 * (include/mcarray/BinauralLocalisation.h:188-247, src/mcarray/BinauralLocalisation.cpp:320-631): correlation curve with
 * 0.8 memory (the _corrMemoryFactor silence state machine of :523-561 included), first-maximum cell, DOA delivered in degrees.
 * The stochastic particle-filter tracker the reference layers on top (BinauralLocalisation.cpp:456-473) is out of scope.
 * By default the callback carries the arg-max cell of each frame with probability 1; with deterministicTracker = true it carries
 * what the reference's `#else` branch publishes (:502-504,521): _currentDOA smoothed by _doaMemoryFactor (0 -> 0.6 -> 0 after 3 s
 * of silence) and the setProbability value (:454,569-631), both computed on the device.
 * Frame length from _frameRate = 0.075 s (BinauralLocalisation.h:196); noise margin 6 dB (:197). */
#ifndef MCARRAY_B200_BINAURALLOCALISATION_H
#define MCARRAY_B200_BINAURALLOCALISATION_H

#include <mcarray/ArrayDescription.h>
#include <mcarray/SoundLocalisationImpl.h>

namespace mca {

class FreqGCCBinauralLocalisation : public LocalisingProcessor {
 public:
  FreqGCCBinauralLocalisation(int sampleRate, ArrayDescription microphonePositions, bool usePowerFloor = true, int streams = 1,
                              int maxFramesPerCall = 256, int device = 0, int frameSize = 0, bool deterministicTracker = false)
      : _tracker(deterministicTracker) {
    if (microphonePositions.size() != 2) throw MCArrayException("Binaural localisation is only working for 2 channels.");
    _doaStep = float(3 * M_PI / 180);   // BinauralLocalisation.cpp:328
    _cellsPerFrame = 1;
    _hasProb = false;
    const double dist = microphonePositions.distance(0, 1);
    const double xyz[6] = {0, 0, 0, dist, 0, 0};   // scalar microphone distance, as :363-366
    const int D = mcag_geom_grid_size(_doaStep);
    std::vector<double> tau(size_t(D), 0.0);
    mcag_geom_pair_tau_reference(xyz, 2, sampleRate, _doaStep, tau.data());
    mcag_config c;
    mcag_config_init(&c);
    c.kind = MCAG_KIND_FREQGCC; c.device = device; c.sample_rate = sampleRate;
    c.frame_size = frameSize ? frameSize : mcag_geom_frame_size(sampleRate, 0.075f);
    c.hop = c.frame_size / 2; c.n_channels = 2; c.n_streams = streams; c.max_frames_per_call = maxFramesPerCall;
    c.n_dirs = D; c.pair_tau = tau.data(); c.use_power_floor = usePowerFloor ? 1 : 0; c.noise_margin_db = 6.0f; c.floor_ccs_power = 1;
    c.corr_memory = 0.8f; c.doa_tracker = deterministicTracker ? 1 : 0; c.doa_memory = 0.6f;   // _maxCorrMemoryFactor / _maxDoaMemoryFactor, .h:198-199
    create(c);
  }
  virtual ~FreqGCCBinauralLocalisation() {}

  /** smoothed correlation curves of the last call: [streams][frames][61] */
  std::vector<float> curves() const {
    std::vector<float> v(size_t(_info.n_streams) * mcag_frames_done(_handle) * _info.n_dirs);
    if (!v.empty()) check(mcag_fetch(_handle, MCAG_OUT_CURVES, v.data(), (long long)v.size() * 4));
    return v;
  }
  /** deterministicTracker only: _currentDOA in radians after every frame of the last call, [streams][frames] */
  std::vector<double> trackedDOA() const {
    std::vector<double> v(size_t(_info.n_streams) * mcag_frames_done(_handle));
    if (!v.empty()) check(mcag_fetch(_handle, MCAG_OUT_TRACK_DOA, v.data(), (long long)v.size() * 8));
    return v;
  }

 protected:
  virtual void deliver(int frames) {
    if (!_tracker) { LocalisingProcessor::deliver(frames); return; }
    if (!_callback || frames <= 0) return;
    const size_t n = size_t(_info.n_streams) * frames;
    std::vector<double> doa(n); std::vector<float> prob(n), power(n); std::vector<unsigned char> active(n);
    check(mcag_fetch(_handle, MCAG_OUT_TRACK_DOA, doa.data(), (long long)n * 8));
    check(mcag_fetch(_handle, MCAG_OUT_PROB, prob.data(), (long long)n * 4));
    check(mcag_fetch(_handle, MCAG_OUT_POWER_DB, power.data(), (long long)n * 4));
    check(mcag_fetch(_handle, MCAG_OUT_ACTIVE, active.data(), (long long)n));
    for (size_t i = 0; i < n; ++i) {
      if (!active[i]) continue;
      SignalPtr d(new BaseType[1]), p(new BaseType[1]);
      d[0] = doa[i] * (180 / M_PI);   // toDegrees, microhponeArrayHelpers.cpp:91-98
      p[0] = prob[i];
      _callback->setDOA(d, p, power[i], 1);   // BinauralLocalisation.cpp:521
    }
  }

 private:
  bool _tracker;
};

}  // namespace mca

#endif

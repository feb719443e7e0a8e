/* mca::FreqGCCBinauralLocalisation — 2-microphone GCC-PHAT localiser on the 61-cell / 3 degree grid
 * (include/mcarray/BinauralLocalisation.h:188-247, src/mcarray/BinauralLocalisation.cpp:320-631): correlation curve with
 * 0.8 memory, first-maximum cell, DOA delivered in degrees.  The stochastic particle-filter tracker the reference layers on
 * top (BinauralLocalisation.cpp:456-473) is out of scope; the callback carries the arg-max cell of each frame,
 * which is what the reference's #else branch (:502-504) starts from, with probability 1.
 * Frame length from _frameRate = 0.075 s (BinauralLocalisation.h:196); noise margin 6 dB (:197). */
#ifndef MCARRAY_B200_BINAURALLOCALISATION_H
#define MCARRAY_B200_BINAURALLOCALISATION_H

#include <mcarray/ArrayDescription.h>
#include <mcarray/SoundLocalisationImpl.h>

namespace mca {

class FreqGCCBinauralLocalisation : public LocalisingProcessor {
 public:
  FreqGCCBinauralLocalisation(int sampleRate, ArrayDescription microphonePositions, bool usePowerFloor = true, int streams = 1,
                              int maxFramesPerCall = 256, int device = 0, int frameSize = 0) {
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
    c.corr_memory = 0.8f;
    create(c);
  }
  virtual ~FreqGCCBinauralLocalisation() {}

  /** smoothed correlation curves of the last call: [streams][frames][61] */
  std::vector<float> curves() const {
    std::vector<float> v(size_t(_info.n_streams) * mcag_frames_done(_handle) * _info.n_dirs);
    if (!v.empty()) check(mcag_fetch(_handle, MCAG_OUT_CURVES, v.data(), (long long)v.size() * 4));
    return v;
  }
};

}  // namespace mca

#endif

/* mca::MultibandBinarualLocalisation — 2-microphone multiband localiser (include/mcarray/MultibandBinarualLocalisation.h:36-99,
 * src/mcarray/MultibandBinarualLocalisation.cpp:52-259): nbins linear sub-bands between 100 Hz and the spatial-aliasing limit
 * c/(2d), per band a GCC-PHAT curve on the 5 degree grid with 0.4 memory and its arg-max, an energy-weighted histogram of
 * the band arg-maxima, and the histogram's arg-max as the published DOA (degrees) with prob = its share of the energy.
 * Frame length from _frameRate = 0.025 s (.h:41); noise margin 3 dB (.h:45).  The class name keeps the reference's spelling. */
#ifndef MCARRAY_B200_MULTIBANDBINARUALLOCALISATION_H
#define MCARRAY_B200_MULTIBANDBINARUALLOCALISATION_H

#include <mcarray/ArrayDescription.h>
#include <mcarray/SoundLocalisationImpl.h>

namespace mca {

class MultibandBinarualLocalisation : public LocalisingProcessor {
 public:
  MultibandBinarualLocalisation(int sampleRate, ArrayDescription microphonePositions, int nbins = 15, bool usePowerFloor = true, int streams = 1,
                                int maxFramesPerCall = 256, int device = 0, int frameSize = 0)
      : _nbins(nbins) {
    if (microphonePositions.size() != 2) throw MCArrayException("Multiband binaural localisation is only working for 2 channels.");
    _doaStep = float(5 * M_PI / 180);   // MultibandBinarualLocalisation.cpp:62
    _cellsPerFrame = 1;
    const double dist = microphonePositions.distance(0, 1);
    const int N = frameSize ? frameSize : mcag_geom_frame_size(sampleRate, 0.025f);
    const int D = mcag_geom_multiband(sampleRate, dist, N, nbins, NULL, NULL);
    std::vector<double> tau(size_t(D), 0.0), H(size_t(nbins) * (N / 2 + 1), 0.0);
    mcag_geom_multiband(sampleRate, dist, N, nbins, tau.data(), H.data());
    mcag_config c;
    mcag_config_init(&c);
    c.kind = MCAG_KIND_MULTIBAND; c.device = device; c.sample_rate = sampleRate; c.frame_size = N; c.hop = N / 2; c.n_channels = 2;
    c.n_streams = streams; c.max_frames_per_call = maxFramesPerCall; c.n_dirs = D; c.pair_tau = tau.data(); c.n_bands = nbins;
    c.band_coefs = H.data(); c.use_power_floor = usePowerFloor ? 1 : 0; c.noise_margin_db = 3.0f; c.corr_memory = 0.4f;   // .h:44-45
    create(c);
  }
  virtual ~MultibandBinarualLocalisation() {}

  int getNumberOfBins() const { return _nbins; }

  /** arg-max cell of every sub-band curve of the last call: [streams][frames][nbins] */
  std::vector<int32_t> bandCells() const {
    std::vector<int32_t> v(size_t(_info.n_streams) * mcag_frames_done(_handle) * _nbins);
    if (!v.empty()) check(mcag_fetch(_handle, MCAG_OUT_BAND_CELL, v.data(), (long long)v.size() * 4));
    return v;
  }
  /** energy-weighted DOA histogram (_energyInDOA) of the last call: [streams][frames][37] */
  std::vector<float> histogram() const {
    std::vector<float> v(size_t(_info.n_streams) * mcag_frames_done(_handle) * _info.n_dirs);
    if (!v.empty()) check(mcag_fetch(_handle, MCAG_OUT_ENERGY, v.data(), (long long)v.size() * 4));
    return v;
  }

 private:
  int _nbins;
};

}  // namespace mca

#endif

/* mca::SourceSeparationAndLocalisation — the processor `mcbeam` runs (src/programs/mcabeamf.cpp:194): STFT -> SRP/GCC-PHAT
 * localisation on the reference's 37-cell azimuth grid -> delay-and-sum separation of numOfSources sources -> overlap-add.
 * Constructor and setCallback signatures as include/mcarray/SourceSeparationAndLocalisation.h:42-56; frame length from
 * _frameRate = 0.025 s (:60).  Extra trailing arguments (all defaulted) are GPU batching knobs: `streams` independent
 * arrays processed in lock step by one object, and the most frames one internal device call may complete. */
#ifndef MCARRAY_B200_SOURCESEPARATIONANDLOCALISATION_H
#define MCARRAY_B200_SOURCESEPARATIONANDLOCALISATION_H

#include <mcarray/ArrayDescription.h>
#include <mcarray/SoundLocalisationImpl.h>

namespace mca {

namespace detail {
/** shared constructor of the two SteeringBeamforming-based processors (SteeringBeamforming.cpp:34-94, Beamformer.cpp:59) */
inline mcag_config steering_config(int kind, int sampleRate, const ArrayDescription &mics, unsigned numOfSources, bool usePowerFloor, int streams,
                                   int maxFramesPerCall, int device, std::vector<double> &tau, std::vector<double> &turns, float doaStep) {
  if (mics.size() < 2) throw MCArrayException("Localisation needs at least two microphones.");
  const std::vector<double> xyz = mics.xyz();
  const int M = int(mics.size()), P = M * (M - 1) / 2, D = mcag_geom_grid_size(doaStep);
  const int N = mcag_geom_frame_size(sampleRate, 0.025f);
  tau.resize(size_t(P) * D);
  turns.resize(size_t(D + 1) * M);
  mcag_geom_pair_tau_reference(xyz.data(), M, sampleRate, doaStep, tau.data());
  mcag_geom_steer_turns_reference(xyz.data(), M, sampleRate, N, doaStep, turns.data());
  mcag_config c;
  mcag_config_init(&c);
  c.kind = kind; c.device = device; c.sample_rate = sampleRate; c.frame_size = N; c.hop = N / 2; c.n_channels = M; c.n_streams = streams;
  c.max_frames_per_call = maxFramesPerCall; c.n_dirs = D; c.pair_tau = tau.data(); c.steer_turns = turns.data();
  c.n_sources = int(numOfSources); c.use_power_floor = usePowerFloor ? 1 : 0; c.noise_margin_db = 3.0f;   // BeamformingSeparationAndLocalistaion.h:52
  return c;
}
}  // namespace detail

class SourceSeparationAndLocalisation : public LocalisingProcessor {
 public:
  SourceSeparationAndLocalisation(int sampleRate, ArrayDescription microphonePositions, unsigned int numOfSources, bool usePowerFloor = true,
                                  int streams = 1, int maxFramesPerCall = 256, int device = 0) {
    _doaStep = float(5 * M_PI / 180);   // SteeringBeamforming.cpp:39
    _cellsPerFrame = int(numOfSources);
    std::vector<double> tau, turns;
    create(detail::steering_config(MCAG_KIND_SSL, sampleRate, microphonePositions, numOfSources, usePowerFloor, streams, maxFramesPerCall, device, tau, turns, _doaStep));
  }
  virtual ~SourceSeparationAndLocalisation() {}
};

}  // namespace mca

#endif

/* mca::SourceLocalisation — analysis-only twin of SourceSeparationAndLocalisation (include/mcarray/SourceLocalisation.h:38-52,
 * src/mcarray/SourceLocalisation.cpp:51-81): STFT -> SRP/GCC-PHAT -> selectDOA, results through the callback. */
#ifndef MCARRAY_B200_SOURCELOCALISATION_H
#define MCARRAY_B200_SOURCELOCALISATION_H

#include <mcarray/SourceSeparationAndLocalisation.h>

namespace mca {

class SourceLocalisation : public LocalisingProcessor {
 public:
  SourceLocalisation(int sampleRate, ArrayDescription microphonePositions, unsigned int numOfSources, bool usePowerFloor = true, int streams = 1,
                     int maxFramesPerCall = 256, int device = 0) {
    _doaStep = float(5 * M_PI / 180);
    _cellsPerFrame = int(numOfSources);
    std::vector<double> tau, turns;
    create(detail::steering_config(MCAG_KIND_SL, sampleRate, microphonePositions, numOfSources, usePowerFloor, streams, maxFramesPerCall, device, tau, turns, _doaStep));
  }
  virtual ~SourceLocalisation() {}
};

}  // namespace mca

#endif

/* Facade classes and the masking enums of include/mcarray/ArrayModules.h:41-106 (src/mcarray/ArrayModules.cpp:30-89).
 * SoundLocalisation picks FreqGCCBinauralLocalisation for two microphones (with the deterministic DOA tracker: the reference facade
 * publishes the particle filter's estimate, the `#else` branch of BinauralLocalisation.cpp:501-504 is its deterministic counterpart)
 * and the SRP localiser otherwise (the reference's
 * >2-microphone branch casts an unrelated type and hard-codes 512, ArrayModules.cpp:46-53: not reproduced);
 * BinauralMasking wraps FastBinauralMasking with the distance between the first two microphones (:77). */
#ifndef MCARRAY_B200_ARRAYMODULES_H
#define MCARRAY_B200_ARRAYMODULES_H

#include <mcarray/ArrayDescription.h>
#include <mcarray/SoundLocalisationCallback.h>

#include <memory>

namespace mca {

class ShortTimeProcessor;

class BinauralMasking {
 public:
  typedef enum { FACTOR = 0, RELATIVE = 1, FULL = 3, NOISY = 4, NOTHING = 5 } MaskingMethod;
  typedef enum { BOTH = 0, SPATIAL = 1, TEMPORAL = 2 } MaskingAlg;

  inline BinauralMasking(int samplerate, ArrayDescription microphones, float lowFreq = 400, float highFreq = 4000, MaskingMethod mmethod = RELATIVE,
                         MaskingAlg algorithm = BOTH);
  inline virtual ~BinauralMasking();
  inline int process(const std::vector<double *> &in, int nsamples, const std::vector<double *> &out, int outbuffersize);
  inline int process(const std::vector<int16_t *> &in, int nsamples, const std::vector<int16_t *> &out, int outbuffersize);
  inline int getMaxLatency() const;
  inline int getFrameSize() const;

 private:
  std::unique_ptr<ShortTimeProcessor> _impl;
};

class SoundLocalisation {
 public:
  inline SoundLocalisation(int sampleRate, ArrayDescription microphonePositions, LocalisationCallback *callback = NULL);
  inline virtual ~SoundLocalisation();
  inline int process(const std::vector<double *> &in, int nsamples);
  inline int process(const std::vector<int16_t *> &in, int nsamples);
  inline int getFrameSize() const;

 private:
  std::unique_ptr<ShortTimeProcessor> _impl;
};

}  // namespace mca

#include <mcarray/BinauralLocalisation.h>
#include <mcarray/FastBinauralMasking.h>
#include <mcarray/SourceLocalisation.h>

namespace mca {

BinauralMasking::BinauralMasking(int samplerate, ArrayDescription microphones, float lowFreq, float highFreq, MaskingMethod mmethod, MaskingAlg algorithm) {
  const double microDist = microphones.size() > 1 ? microphones.distance(0, 1) : 0;
  _impl.reset(new FastBinauralMasking(samplerate, microDist, lowFreq, highFreq, mmethod, algorithm));
}
BinauralMasking::~BinauralMasking() {}
int BinauralMasking::process(const std::vector<double *> &in, int n, const std::vector<double *> &out, int cap) { return _impl->process(in, n, out, cap); }
int BinauralMasking::process(const std::vector<int16_t *> &in, int n, const std::vector<int16_t *> &out, int cap) { return _impl->process(in, n, out, cap); }
int BinauralMasking::getMaxLatency() const { return _impl->getMaxLatency(); }
int BinauralMasking::getFrameSize() const { return _impl->getFrameSize(); }

SoundLocalisation::SoundLocalisation(int sampleRate, ArrayDescription microphonePositions, LocalisationCallback *callback) {
  const bool usePowerFloor = true;
  LocalisingProcessor *loc;
  if (microphonePositions.size() == 2) loc = new FreqGCCBinauralLocalisation(sampleRate, microphonePositions, usePowerFloor, 1, 256, 0, 0, true);
  else loc = new SourceLocalisation(sampleRate, microphonePositions, 1, usePowerFloor);
  if (callback != NULL) loc->setCallback(callback);
  _impl.reset(loc);
}
SoundLocalisation::~SoundLocalisation() {}
int SoundLocalisation::process(const std::vector<double *> &in, int n) { return _impl->process(in, n); }
int SoundLocalisation::process(const std::vector<int16_t *> &in, int n) { return _impl->process(in, n); }
int SoundLocalisation::getFrameSize() const { return _impl->getFrameSize(); }

}  // namespace mca

#endif

/* LocalisationCallback — result delivery of the localisation processors, same contract as the reference
 * (include/mcarray/SoundLocalisationCallback.h:53): called synchronously on the caller's thread, once per processed
 * (above-floor) frame, in frame order, with the DOAs in DEGREES.  The GPU build runs a whole batch of frames, copies the
 * [T][S] results back and then fires the callbacks before process() returns. */
#ifndef MCARRAY_B200_SOUNDLOCALISATIONCALLBACK_H
#define MCARRAY_B200_SOUNDLOCALISATIONCALLBACK_H

#include <mcarray/mcadefs.h>

#include <iostream>

namespace mca {

class LocalisationCallback {
 public:
  LocalisationCallback() {}
  virtual ~LocalisationCallback() {}
  virtual void setDOA(SignalPtr doa, SignalPtr prob, double power, int numOfSources) = 0;
};

/** prints every DOA it receives (SoundLocalisationCallback.cpp: "[DOA: x, p=y, P=z]") */
class DummyLocalisationCallback : public LocalisationCallback {
 public:
  using LocalisationCallback::setDOA;
  virtual void setDOA(SignalPtr doa, SignalPtr prob, double power, int numOfSources) {
    for (int i = 0; i < numOfSources; ++i) setDOA(doa[i], prob[i], power);
  }
  virtual void setDOA(double doa, double prob, double power) { std::cout << "[DOA: " << doa << ", p=" << prob << ", P=" << power << "] " << std::endl; }
};

}  // namespace mca

#endif

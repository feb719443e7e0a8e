/* mca::FastBinauralMasking — frequency-domain binaural (spatial + temporal) masking on a 45-band mel bank
 * (include/mcarray/FastBinauralMasking.h:71-128, src/mcarray/FastBinauralMasking.cpp:51-538): STFT -> per-band power
 * tracker / interaural correlation -> FULL / RELATIVE / FACTOR / NOISY attenuation -> overlap-add, two channels in, two out.
 * Frame length from 0.050 s (FastBinauralMasking.h:112).  Configuration errors throw MCArrayException like :88-104. */
#ifndef MCARRAY_B200_FASTBINAURALMASKING_H
#define MCARRAY_B200_FASTBINAURALMASKING_H

#include <mcarray/ArrayModules.h>
#include <mcarray/ShortTimeProcessor.h>

namespace mca {

class FastBinauralMasking : public ShortTimeProcessor {
 public:
  typedef BinauralMasking::MaskingMethod MaskingMethod;
  typedef BinauralMasking::MaskingAlg MaskingAlg;

  FastBinauralMasking(int samplerate, double microDistance, float lowFreq, float highFreq, MaskingMethod mmethod = BinauralMasking::RELATIVE,
                      MaskingAlg algorithm = BinauralMasking::BOTH, int streams = 1, int maxFramesPerCall = 256, int device = 0, int frameSize = 0)
      : _microDistance(microDistance) {
    const int N = frameSize ? frameSize : mcag_geom_frame_size(samplerate, 0.050f);
    std::vector<double> H(size_t(_nbins) * (N / 2 + 1)), fc(_nbins), thr(_nbins);
    mcag_geom_mel_bank(N, _nbins, samplerate, lowFreq, highFreq, microDistance, H.data(), fc.data(), thr.data());
    mcag_config c;
    mcag_config_init(&c);
    c.kind = MCAG_KIND_MASK; c.device = device; c.sample_rate = samplerate; c.frame_size = N; c.hop = N / 2; c.n_channels = 2; c.n_streams = streams;
    c.max_frames_per_call = maxFramesPerCall; c.mask_method = int(mmethod); c.mask_alg = int(algorithm); c.n_bands = _nbins;
    c.band_coefs = H.data(); c.band_thresholds = thr.data();
    create(c);
  }
  virtual ~FastBinauralMasking() {}

  inline int getNonMaskingAngle() { return int((10 * M_PI / 180) * 180 * M_1_PI); }   // _phi, FastBinauralMasking.h:89,113
  inline float getMicroPhoneDistance() { return float(_microDistance); }
  inline float getSpatialMaskingFactor() { return 1 / 10.0f; }         // 1/_spatialMaskingFactor, FastBinauralMasking.h:99,118
  inline float getTemporalMaskingFactor() { return 1 / 3.0f; }         // 1/_temporalMaskingFactor, FastBinauralMasking.h:104,117

 private:
  static const int _nbins = 45;                                         // FastBinauralMasking.h:111
  const double _microDistance;
};

}  // namespace mca

#endif

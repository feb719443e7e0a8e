/* mca::BeamformingSeparationAndLocalisation — per-frame driver (file name keeps the reference's spelling):
 * power-floor gate, SteeringBeamforming localisation, callback in degrees, then one Beamformer pass per source
 * (include/mcarray/BeamformingSeparationAndLocalistaion.h:40-44, src/mcarray/BeamformingSeparationAndLocalisation.cpp:29-119).
 * Power is dsp::SignalPower::FFTLogPower of the frame (oracle/CONVENTIONS.md C6), the floor is estimated over the first 3 s
 * with a 3 dB margin (:55-72, SoundLocalisationImpl.h:77). */
#ifndef MCARRAY_B200_BEAMFORMINGSEPARATIONANDLOCALISATION_H
#define MCARRAY_B200_BEAMFORMINGSEPARATIONANDLOCALISATION_H

#include <mcarray/SoundLocalisationCallback.h>
#include <mcarray/SteeringBeamforming.h>

#include <algorithm>
#include <cmath>

namespace mca {

class BeamformingSeparationAndLocalisation {
 public:
  BeamformingSeparationAndLocalisation(int sampleRate, int fftCCSLength, ArrayDescription microphonePositions, unsigned int numOfSources, bool usePowerFloor)
      : _nchannels(unsigned(microphonePositions.size())), _sampleRate(sampleRate), _fftCCSLength(fftCCSLength), _usePowerFloor(usePowerFloor),
        _numOfSources(numOfSources), _steeringBeamforming(sampleRate, microphonePositions, fftCCSLength, _nchannels),
        _beamformer(sampleRate, microphonePositions, fftCCSLength, _nchannels), _ptrCallback(NULL), _powerFloor(0), _samplesConsumedForNoise(0),
        _noiseEstimated(false) {
    for (unsigned c = 0; c < _nchannels; ++c) _inputFrames.push_back(SignalPtr(new BaseType[_fftCCSLength]));
    _currentDOA.reset(new BaseType[_numOfSources]);
    _prob.reset(new BaseType[_numOfSources]);
    std::fill(_currentDOA.get(), _currentDOA.get() + _numOfSources, 0.0);
    std::fill(_prob.get(), _prob.get() + _numOfSources, -1.0);
  }
  virtual ~BeamformingSeparationAndLocalisation() {}

  void setCallback(LocalisationCallback &callback) { _ptrCallback = &callback; }
  void setCallback(LocalisationCallback *callback) { _ptrCallback = callback; }

  void processFrameLocalisation(SignalVector &analysisFrames, SignalVector &wienerCoefs) {
    BaseType power;
    const double lin = fftPower(analysisFrames);
    if (!_noiseEstimated && _usePowerFloor) {
      _powerFloor += lin * (_fftCCSLength - 2);
      _samplesConsumedForNoise += _fftCCSLength - 2;
      if (_samplesConsumedForNoise >= int(3.0f * float(_sampleRate))) {
        _noiseEstimated = true;
        _powerFloor /= _samplesConsumedForNoise;
        _powerFloor = 10 * std::log10(_powerFloor) + 3.0;
      }
      power = _powerFloor;
    } else {
      power = 10 * std::log10(lin);
    }
    if ((power > _powerFloor) || !_usePowerFloor) {
      _steeringBeamforming.processFrame(analysisFrames, _currentDOA, _prob, int(_numOfSources), wienerCoefs);
      if (_ptrCallback) {
        SignalPtr deg(new BaseType[_numOfSources]);
        for (unsigned s = 0; s < _numOfSources; ++s) deg[s] = _currentDOA[s] * 180.0 / M_PI;
        _ptrCallback->setDOA(deg, _prob, power, int(_numOfSources));
      }
    }
  }

  void processFrameSeparation(SignalVector &inputFrames, SignalVector &outputFrames) {
    unsigned c;
    for (c = 0; c < _nchannels; ++c) std::copy(inputFrames[c].get(), inputFrames[c].get() + _fftCCSLength, _inputFrames[c].get());   // in == out is allowed
    for (c = 0; c < std::min(_nchannels, _numOfSources); ++c) _beamformer.processFrame(_inputFrames, outputFrames[c], _currentDOA[c]);
    for (; c < _nchannels; ++c) std::fill(outputFrames[c].get(), outputFrames[c].get() + _fftCCSLength, 0.0);
  }

 private:
  /** mean over channels of the time-domain mean square, by Parseval on the one-sided spectrum */
  double fftPower(const SignalVector &frames) const {
    const int N = _fftCCSLength - 2, K = N / 2 + 1;
    double acc = 0;
    for (unsigned c = 0; c < _nchannels; ++c) {
      double s = 0;
      for (int k = 0; k < K; ++k) {
        const double re = frames[c][2 * k], im = frames[c][2 * k + 1];
        s += ((k == 0 || k == K - 1) ? 1.0 : 2.0) * (re * re + im * im);
      }
      acc += s / (double(N) * double(N));
    }
    return acc / _nchannels;
  }

  const unsigned int _nchannels;
  int _sampleRate, _fftCCSLength;
  bool _usePowerFloor;
  unsigned int _numOfSources;
  SignalVector _inputFrames;
  SteeringBeamforming _steeringBeamforming;
  Beamformer _beamformer;
  LocalisationCallback *_ptrCallback;
  SignalPtr _currentDOA, _prob;
  double _powerFloor;
  int _samplesConsumedForNoise;
  bool _noiseEstimated;
};

}  // namespace mca

#endif

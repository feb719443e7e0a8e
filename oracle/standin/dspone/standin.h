// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// Stand-in for the DSPONE library (jordi-adell/dspone, un-vendored, un-pinned:
// /root/reference/CMakeLists.txt:30, /root/reference/cmake/FindDSPONE.cmake:9-10,
// /root/reference/.travis.yml:21-22).  DSPONE owns the STFT framework, the FFT, the
// generalised cross-correlation, the mel filter bank and the SignalPower family that the
// reference calls; its source is absent from /root/reference, so the classes below restate
// the *published* algorithms behind the API the reference's call sites use.
// PARITY UNPINNED at this boundary; every choice is listed in oracle/CONVENTIONS.md and each
// is an explicit input of the CUDA library (window, hop, tau tables, filter bank, ...).
//
// The API (names, argument order) follows the reference's call sites so that
//   (a) oracle/restated.hpp can be written against it, and
//   (b) the reference's OWN .cpp files compile against it unmodified (oracle/Makefile, `make ref`),
//       which pins the restatement of the mcarray-owned logic to the literal reference code.
#ifndef ORACLE_STANDIN_DSPONE_H
#define ORACLE_STANDIN_DSPONE_H

#include <wipp/wipp.h>
#include <boost/shared_array.hpp>
#include <boost/scoped_ptr.hpp>

#include <cmath>
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

// DSPONE ships its own stream-logging macros; some reference sources use them without including
// mcalogger.h (BeamformingSeparationAndLocalisation.cpp:68,95,99).  No-ops here.
#ifndef INFO_STREAM
#define INFO_STREAM(x)
#endif
#ifndef WARN_STREAM
#define WARN_STREAM(x)
#endif
#ifndef ERROR_STREAM
#define ERROR_STREAM(x)
#endif
#ifndef DEBUG_STREAM
#define DEBUG_STREAM(x)
#endif
#ifndef TRACE_STREAM
#define TRACE_STREAM(x)
#endif
#ifndef WARN_STREAM_ONCE
#define WARN_STREAM_ONCE(x)
#endif

namespace dsp {

typedef double BaseType;
typedef struct { double re; double im; } Complex;
typedef boost::shared_array<double> SignalPtr;
typedef std::vector<SignalPtr> SignalVector;

// ------------------------------------------------------------------------------------------------
// FFT: forward unnormalised, inverse scaled by 1/N; real data <-> CCS buffer of N+2 doubles
// (K = N/2+1 interleaved re,im).  Call sites: test_mcarray.cpp:672,723,747 (fwdTransform,
// invTransfrom [sic]).  Iterative radix-2, double precision.
// ------------------------------------------------------------------------------------------------
class FFT {
 public:
  explicit FFT(int order) : _order(order), _n(1 << order), _tw(size_t(_n / 2)), _rev(size_t(_n)) {
    for (int i = 0; i < _n / 2; ++i) _tw[size_t(i)] = std::polar(1.0, -2.0 * M_PI * double(i) / double(_n));
    for (int i = 0; i < _n; ++i) {
      int r = 0;
      for (int b = 0; b < order; ++b) if (i & (1 << b)) r |= 1 << (order - 1 - b);
      _rev[size_t(i)] = r;
    }
    _work.resize(size_t(_n));
  }
  int getFFTLength() const { return _n; }
  int getOneSidedFFTLength() const { return _n / 2 + 1; }
  int getAnalysisLength() const { return _n + 2; }

  void fwdTransform(const double *time, double *ccs) {
    for (int i = 0; i < _n; ++i) _work[size_t(_rev[size_t(i)])] = std::complex<double>(time[i], 0.0);
    butterflies(false);
    for (int k = 0; k <= _n / 2; ++k) { ccs[2 * k] = _work[size_t(k)].real(); ccs[2 * k + 1] = _work[size_t(k)].imag(); }
  }
  void invTransfrom(double *time, const double *ccs) {
    std::vector<std::complex<double>> full(static_cast<size_t>(_n));
    for (int k = 0; k <= _n / 2; ++k) full[size_t(k)] = std::complex<double>(ccs[2 * k], ccs[2 * k + 1]);
    for (int k = _n / 2 + 1; k < _n; ++k) full[size_t(k)] = std::conj(full[size_t(_n - k)]);
    // a real signal has real DC and Nyquist terms; like FFTW's c2r the imaginary parts are ignored
    full[0] = std::complex<double>(full[0].real(), 0.0);
    full[size_t(_n / 2)] = std::complex<double>(full[size_t(_n / 2)].real(), 0.0);
    for (int i = 0; i < _n; ++i) _work[size_t(_rev[size_t(i)])] = full[size_t(i)];
    butterflies(true);
    for (int i = 0; i < _n; ++i) time[i] = _work[size_t(i)].real() / double(_n);
  }
  void invTransform(double *time, const double *ccs) { invTransfrom(time, ccs); }

 private:
  void butterflies(bool inverse) {
    for (int len = 2; len <= _n; len <<= 1) {
      const int half = len / 2, step = _n / len;
      for (int s = 0; s < _n; s += len)
        for (int j = 0; j < half; ++j) {
          std::complex<double> w = _tw[size_t(j * step)];
          if (inverse) w = std::conj(w);
          std::complex<double> a = _work[size_t(s + j)], b = _work[size_t(s + j + half)] * w;
          _work[size_t(s + j)] = a + b;
          _work[size_t(s + j + half)] = a - b;
        }
    }
  }
  int _order, _n;
  std::vector<std::complex<double>> _tw;
  std::vector<int> _rev;
  std::vector<std::complex<double>> _work;
};

// ------------------------------------------------------------------------------------------------
// ShortTimeProcess: framing / analysis window / per-frame hook / synthesis window / overlap-add.
// Conventions (CONVENTIONS.md C1-C4): window = sqrt of the periodic Hann, hop = N/2, FIFO starts
// empty, a frame is processed whenever N samples are buffered, each processed frame emits `hop`
// finished output samples, getMaxLatency() = N.
// API follows the call sites: mcabeamf.cpp:85,112; test_mcarray.cpp:596,618,622,843-844,896,937.
// ------------------------------------------------------------------------------------------------
class ShortTimeProcess {
 public:
  typedef enum { ANALYSIS_SYNTHESIS = 0, ANALYSIS = 1 } Mode;

  ShortTimeProcess(int windowSize, int analysisLength, int nchannels, Mode mode = ANALYSIS_SYNTHESIS)
      : _windowSize(windowSize), _windowShift(windowSize / 2), _analysisLength(analysisLength),
        _nchannels(nchannels), _mode(mode), _framesProcessed(0) {
    _window.resize(size_t(_windowSize));
    for (int n = 0; n < _windowSize; ++n)
      _window[size_t(n)] = std::sqrt(0.5 * (1.0 - std::cos(2.0 * M_PI * double(n) / double(_windowSize))));
    _fifo.assign(size_t(_nchannels), std::vector<double>());
    _olaTail.assign(size_t(_nchannels), std::vector<double>(size_t(_windowSize - _windowShift), 0.0));
    for (int c = 0; c < _nchannels; ++c) {
      _analysisStore.push_back(std::vector<double>(size_t(_analysisLength), 0.0));
      _frameStore.push_back(std::vector<double>(size_t(_windowSize), 0.0));
    }
  }
  virtual ~ShortTimeProcess() {}

  // N = 2^ceil(log2(frameRate*fs)); consistent with every data point in the reference
  // (0.025 s: 16 kHz -> 512, 48 kHz -> 2048 = test_mcarray.cpp:659-662).
  static int calculateOrderFromSampleRate(int sampleRate, double frameRate) {
    double want = frameRate * double(sampleRate);
    int order = 0;
    while (double(1 << order) < want) ++order;
    return order;
  }

  int getFrameSize() const { return _windowShift; }
  int getWindowSize() const { return _windowSize; }
  int getWindowShift() const { return _windowShift; }
  int getAnalysisLength() const { return _analysisLength; }
  int getNumberOfChannels() const { return _nchannels; }
  int getMaxLatency() const { return _windowSize; }
  int getLatency() const { return _fifo.empty() ? 0 : int(_fifo[0].size()); }
  long getFramesProcessed() const { return _framesProcessed; }
  // inverse-window helper used only by the out-of-scope TemporalGCC path (BinauralLocalisation.cpp:107)
  void unwindowFrame(const double *in, double *out, int length) const {
    for (int n = 0; n < length; ++n) out[n] = _window[size_t(n)] > 1e-12 ? in[n] / _window[size_t(n)] : 0.0;
  }
  void setWindow(const double *analysisWindow) { for (int n = 0; n < _windowSize; ++n) _window[size_t(n)] = analysisWindow[n]; }
  const std::vector<double> &getWindow() const { return _window; }

  // analysis + synthesis; returns the number of output samples written per channel
  int process(const std::vector<double *> &in, int nsamples, const std::vector<double *> &out, int outbuffersize) {
    return run(in, nsamples, &out, outbuffersize);
  }
  int process(const std::vector<int16_t *> &in, int nsamples, const std::vector<int16_t *> &out, int outbuffersize) {
    std::vector<std::vector<double>> din(in.size(), std::vector<double>(size_t(nsamples)));
    std::vector<std::vector<double>> dout(out.size(), std::vector<double>(size_t(outbuffersize)));
    std::vector<double *> pin, pout;
    for (size_t c = 0; c < in.size(); ++c) { for (int i = 0; i < nsamples; ++i) din[c][size_t(i)] = in[c][i]; pin.push_back(din[c].data()); }
    for (size_t c = 0; c < out.size(); ++c) pout.push_back(dout[c].data());
    int n = run(pin, nsamples, &pout, outbuffersize);
    for (size_t c = 0; c < out.size(); ++c)
      for (int i = 0; i < n; ++i) {
        double v = std::nearbyint(dout[c][size_t(i)]);
        out[c][i] = int16_t(std::min(32767.0, std::max(-32768.0, v)));
      }
    return n;
  }
  // analysis only (test_mcarray.cpp:618,622)
  int process(const SignalVector &in, int nsamples) {
    std::vector<double *> pin;
    for (size_t c = 0; c < in.size(); ++c) pin.push_back(in[c].get());
    return run(pin, nsamples, nullptr, 0);
  }
  int process(const std::vector<double *> &in, int nsamples) { return run(in, nsamples, nullptr, 0); }

 protected:
  virtual void frameAnalysis(double *inFrame, double *analysis, int frameLength, int analysisLength, int channel) = 0;
  virtual void processParametrisation(std::vector<double *> &analysisFrames, int analysisLength,
                                      std::vector<double *> &dataChannels, int dataLength) = 0;
  virtual void frameSynthesis(double *outFrame, double *analysis, int frameLength, int analysisLength, int channel) = 0;

  int _windowSize, _windowShift, _analysisLength, _nchannels;
  Mode _mode;

 private:
  int run(const std::vector<double *> &in, int nsamples, const std::vector<double *> *out, int outcap) {
    if (int(in.size()) < _nchannels) throw std::runtime_error("ShortTimeProcess: too few input channels");
    for (int c = 0; c < _nchannels; ++c) _fifo[size_t(c)].insert(_fifo[size_t(c)].end(), in[size_t(c)], in[size_t(c)] + nsamples);
    const bool synth = (_mode == ANALYSIS_SYNTHESIS) && out != nullptr;
    int written = 0;
    size_t consumed = 0;
    std::vector<double *> analysisPtrs, dataPtrs;
    for (int c = 0; c < _nchannels; ++c) analysisPtrs.push_back(_analysisStore[size_t(c)].data());
    while (_fifo[0].size() - consumed >= size_t(_windowSize)) {
      if (synth && written + _windowShift > outcap) break;
      for (int c = 0; c < _nchannels; ++c) {
        double *frame = _frameStore[size_t(c)].data();
        const double *src = _fifo[size_t(c)].data() + consumed;
        for (int n = 0; n < _windowSize; ++n) frame[n] = src[n] * _window[size_t(n)];
        frameAnalysis(frame, analysisPtrs[size_t(c)], _windowSize, _analysisLength, c);
      }
      processParametrisation(analysisPtrs, _analysisLength, dataPtrs, 0);
      if (synth) {
        for (int c = 0; c < _nchannels; ++c) {
          double *frame = _frameStore[size_t(c)].data();
          frameSynthesis(frame, analysisPtrs[size_t(c)], _windowSize, _analysisLength, c);
          std::vector<double> &tail = _olaTail[size_t(c)];
          double *dst = (*out)[size_t(c)] + written;
          const int ov = _windowSize - _windowShift;
          for (int n = 0; n < _windowShift; ++n) dst[n] = frame[n] * _window[size_t(n)] + (n < ov ? tail[size_t(n)] : 0.0);
          // shift the tail by hop and add the remainder of this frame
          for (int n = 0; n < ov; ++n) {
            double carry = (n + _windowShift < ov) ? tail[size_t(n + _windowShift)] : 0.0;
            tail[size_t(n)] = carry + frame[n + _windowShift] * _window[size_t(n + _windowShift)];
          }
        }
        written += _windowShift;
      }
      consumed += size_t(_windowShift);
      ++_framesProcessed;
    }
    for (int c = 0; c < _nchannels; ++c) _fifo[size_t(c)].erase(_fifo[size_t(c)].begin(), _fifo[size_t(c)].begin() + long(consumed));
    return written;
  }

  std::vector<double> _window;
  std::vector<std::vector<double>> _fifo, _olaTail, _analysisStore, _frameStore;
  long _framesProcessed;
};

class ShortTimeAnalysis : public ShortTimeProcess {
 public:
  ShortTimeAnalysis(int windowSize, int analysisLength, int nchannels) : ShortTimeProcess(windowSize, analysisLength, nchannels, ANALYSIS) {}
 protected:
  virtual void frameSynthesis(double *, double *, int, int, int) {}
};

// STFT: frameAnalysis = real FFT into CCS, frameSynthesis = inverse.  Ctor (nchannels, order):
// SourceSeparationAndLocalisation.cpp:51-52, FastBinauralMasking.cpp:57.
class STFT : public ShortTimeProcess {
 public:
  static const int _defaultFFTOrder = 9;
  STFT(int nchannels, int order) : ShortTimeProcess(1 << order, (1 << order) + 2, nchannels, ANALYSIS_SYNTHESIS), _fft(order) {}
  int getOneSidedFFTLength() const { return _windowSize / 2 + 1; }
 protected:
  virtual void frameAnalysis(double *inFrame, double *analysis, int, int, int) { _fft.fwdTransform(inFrame, analysis); }
  virtual void frameSynthesis(double *outFrame, double *analysis, int, int, int) { _fft.invTransfrom(outFrame, analysis); }
  FFT _fft;
};

class STFTAnalysis : public ShortTimeAnalysis {
 public:
  STFTAnalysis(int nchannels, int order) : ShortTimeAnalysis(1 << order, (1 << order) + 2, nchannels), _fft(order) {}
  int getOneSidedFFTLength() const { return _windowSize / 2 + 1; }
 protected:
  virtual void frameAnalysis(double *inFrame, double *analysis, int, int, int) { _fft.fwdTransform(inFrame, analysis); }
  FFT _fft;
};

// ------------------------------------------------------------------------------------------------
// GeneralisedCrossCorrelation (GCC-PHAT evaluated at fractional delays).  Call sites:
// SteeringBeamforming.cpp:84-88,115-119; BinauralLocalisation.cpp:331,371,438-442;
// MultibandBinarualLocalisation.cpp:69,175-176.   Conventions C5:
//   G[k]    = x[k]*conj(y[k]) / |x[k]*conj(y[k])|   (0 when the magnitude is 0)
//   corr[d] = sum_{k<length} G[k] * exp(+j*2*pi*k*tau_d/Nfft),  Nfft = 2*(length-1) one-sided
// (unnormalised sum over all one-sided bins, DC and Nyquist included).
// ------------------------------------------------------------------------------------------------
class GeneralisedCrossCorrelation {
 public:
  typedef enum { ONESIDEDFFT = 0, TWOSIDEDFFT = 1 } SpectralRepresentation;
  GeneralisedCrossCorrelation(int length, SpectralRepresentation rep) : _length(length), _rep(rep), _ntau(0) {}

  static int fftLength(int length, SpectralRepresentation rep) { return rep == ONESIDEDFFT ? 2 * (length - 1) : length; }

  void precomputeTauMatrix(const double *tau, int ntau, int length, SpectralRepresentation rep) {
    _ntau = ntau;
    _table.assign(size_t(ntau) * size_t(length), std::complex<double>());
    const double nfft = double(fftLength(length, rep));
    for (int d = 0; d < ntau; ++d)
      for (int k = 0; k < length; ++k)
        _table[size_t(d) * size_t(length) + size_t(k)] = std::polar(1.0, 2.0 * M_PI * double(k) * tau[d] / nfft);
  }
  static void phat(const Complex *x, const Complex *y, std::vector<std::complex<double>> &g, int length) {
    g.resize(size_t(length));
    for (int k = 0; k < length; ++k) {
      std::complex<double> c = std::complex<double>(x[k].re, x[k].im) * std::conj(std::complex<double>(y[k].re, y[k].im));
      double m = std::abs(c);
      g[size_t(k)] = m > 0.0 ? c / m : std::complex<double>(0.0, 0.0);
    }
  }
  void calculateCorrelationsForPrecomputedTauMatrix(const Complex *x, const Complex *y, Complex *corr, int length, int ntau,
                                                    SpectralRepresentation) {
    phat(x, y, _g, length);
    for (int d = 0; d < ntau; ++d) {
      std::complex<double> acc(0.0, 0.0);
      const std::complex<double> *row = &_table[size_t(d) * size_t(length)];
      for (int k = 0; k < length; ++k) acc += _g[size_t(k)] * row[k];
      corr[d].re = acc.real(); corr[d].im = acc.imag();
    }
  }
  void calculateCorrelationsForTauVector(const Complex *x, const Complex *y, Complex *corr, int length, const double *tau, int ntau,
                                         SpectralRepresentation rep) {
    phat(x, y, _g, length);
    const double nfft = double(fftLength(length, rep));
    for (int d = 0; d < ntau; ++d) {
      std::complex<double> acc(0.0, 0.0);
      for (int k = 0; k < length; ++k) acc += _g[size_t(k)] * std::polar(1.0, 2.0 * M_PI * double(k) * tau[d] / nfft);
      corr[d].re = acc.real(); corr[d].im = acc.imag();
    }
  }

 private:
  int _length; SpectralRepresentation _rep; int _ntau;
  std::vector<std::complex<double>> _table, _g;
};

// ------------------------------------------------------------------------------------------------
// SignalPower (C6).  FFTPower = mean over channels of the time-domain mean square recovered from
// the one-sided spectrum by Parseval; power() is the plain mean square of the buffer it is given
// (the reference also feeds it CCS buffers: BinauralLocalisation.cpp:390).
// Call sites: BeamformingSeparationAndLocalisation.cpp:58,83; BinauralLocalisation.cpp:390,432.
// ------------------------------------------------------------------------------------------------
class SignalPower {
 public:
  static double FFTPowerOne(const double *ccs, int ccsLength) {
    const int n = ccsLength - 2, K = ccsLength / 2;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) {
      double p = ccs[2 * k] * ccs[2 * k] + ccs[2 * k + 1] * ccs[2 * k + 1];
      acc += (k == 0 || k == K - 1) ? p : 2.0 * p;
    }
    return acc / (double(n) * double(n));
  }
  template <class V> static double FFTPower(const V &frames, int ccsLength) {
    double acc = 0.0;
    for (size_t c = 0; c < frames.size(); ++c) acc += FFTPowerOne(ptr(frames[c]), ccsLength);
    return frames.empty() ? 0.0 : acc / double(frames.size());
  }
  template <class V> static double FFTLogPower(const V &frames, int ccsLength) { return 10.0 * std::log10(FFTPower(frames, ccsLength)); }
  template <class V> static double power(const V &frames, int length) {
    double acc = 0.0;
    for (size_t c = 0; c < frames.size(); ++c) { const double *p = ptr(frames[c]); double a = 0; for (int i = 0; i < length; ++i) a += p[i] * p[i]; acc += a / double(length); }
    return frames.empty() ? 0.0 : acc / double(frames.size());
  }
  template <class V> static double logPower(const V &frames, int length) { return 10.0 * std::log10(power(frames, length)); }
  static double logPower(const int16_t *x, int length) { double a = 0; for (int i = 0; i < length; ++i) a += double(x[i]) * double(x[i]); return 10.0 * std::log10(a / double(length)); }
  static double logPower(const double *x, int length) { double a = 0; for (int i = 0; i < length; ++i) a += x[i] * x[i]; return 10.0 * std::log10(a / double(length)); }
 private:
  static const double *ptr(const SignalPtr &p) { return p.get(); }
  static const double *ptr(const double *p) { return p; }
  static const double *ptr(double *p) { return p; }
};

// ------------------------------------------------------------------------------------------------
// Mel filter bank in the FFT domain (C7): nBins triangular magnitude responses with unit peak,
// centres equally spaced on the mel scale (2595*log10(1+f/700)) between minFreq and maxFreq,
// evaluated on the one-sided bin grid f_k = k*fs/N.  Call sites: FastBinauralMasking.cpp:95-98,361.
// ------------------------------------------------------------------------------------------------
class FilterBank {
 public:
  virtual ~FilterBank() {}
  virtual int getFiltersCoeficients(double *coefs, int length) const = 0;
  virtual double getBinCenterFrequency(int bin) const = 0;  // normalised by the sample rate
};

class FilterBankFFTWMelScale : public FilterBank {
 public:
  FilterBankFFTWMelScale(int order, int nBins, int sampleRate, float minFreq, float maxFreq)
      : _nBins(nBins), _K((1 << (order - 1)) + 1), _fs(sampleRate) {
    const int N = 1 << order;
    const double mlo = hz2mel(minFreq), mhi = hz2mel(maxFreq);
    std::vector<double> edges(size_t(nBins) + 2);
    for (int i = 0; i < nBins + 2; ++i) edges[size_t(i)] = mel2hz(mlo + (mhi - mlo) * double(i) / double(nBins + 1));
    _coefs.assign(size_t(nBins) * size_t(_K), 0.0);
    _centres.resize(size_t(nBins));
    for (int b = 0; b < nBins; ++b) {
      const double lo = edges[size_t(b)], mid = edges[size_t(b) + 1], hi = edges[size_t(b) + 2];
      _centres[size_t(b)] = mid / double(sampleRate);
      for (int k = 0; k < _K; ++k) {
        const double f = double(k) * double(sampleRate) / double(N);
        double v = 0.0;
        if (f > lo && f <= mid) v = (f - lo) / (mid - lo);
        else if (f > mid && f < hi) v = (hi - f) / (hi - mid);
        _coefs[size_t(b) * size_t(_K) + size_t(k)] = v;
      }
    }
  }
  virtual int getFiltersCoeficients(double *coefs, int length) const {
    int n = std::min(length, int(_coefs.size()));
    for (int i = 0; i < n; ++i) coefs[i] = _coefs[size_t(i)];
    return int(_coefs.size());
  }
  virtual double getBinCenterFrequency(int bin) const { return _centres[size_t(bin)]; }
  static double hz2mel(double f) { return 2595.0 * std::log10(1.0 + f / 700.0); }
  static double mel2hz(double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); }
 private:
  int _nBins, _K, _fs;
  std::vector<double> _coefs, _centres;
};

// ------------------------------------------------------------------------------------------------
// Linear filter bank + SubBandSTFTAnalysis (C8).  Call site: MultibandBinarualLocalisation.cpp:52-60
// (ctor `SubBandSTFTAnalysis(nbins, fs, order, 2, 100, maxFreq, SubBandSTFT::LINEAR)`), hooks
// processSetup / processOneSubband(frame, length, bin) / processSumamry (:145-259).
// Stand-in semantics: the bank has nbins triangular magnitude responses with unit peak whose centres
// are equally spaced in Hz between minFreq and maxFreq (same shape rule as the mel bank, linear axis);
// per frame the hooks see, for every band, the CCS spectra of all channels multiplied by that band's
// real response (a frequency-domain filter bank, like FastBinauralMasking.cpp:148-153 applies its own).
// ------------------------------------------------------------------------------------------------
class FilterBankFFTWLinear : public FilterBank {
 public:
  FilterBankFFTWLinear(int order, int nBins, int sampleRate, float minFreq, float maxFreq)
      : _nBins(nBins), _K((1 << (order - 1)) + 1) {
    const int N = 1 << order;
    std::vector<double> edges(size_t(nBins) + 2);
    for (int i = 0; i < nBins + 2; ++i) edges[size_t(i)] = double(minFreq) + (double(maxFreq) - double(minFreq)) * double(i) / double(nBins + 1);
    _coefs.assign(size_t(nBins) * size_t(_K), 0.0);
    _centres.resize(size_t(nBins));
    for (int b = 0; b < nBins; ++b) {
      const double lo = edges[size_t(b)], mid = edges[size_t(b) + 1], hi = edges[size_t(b) + 2];
      _centres[size_t(b)] = mid / double(sampleRate);
      for (int k = 0; k < _K; ++k) {
        const double f = double(k) * double(sampleRate) / double(N);
        double v = 0.0;
        if (f > lo && f <= mid) v = (f - lo) / (mid - lo);
        else if (f > mid && f < hi) v = (hi - f) / (hi - mid);
        _coefs[size_t(b) * size_t(_K) + size_t(k)] = v;
      }
    }
  }
  virtual int getFiltersCoeficients(double *coefs, int length) const {
    int n = std::min(length, int(_coefs.size()));
    for (int i = 0; i < n; ++i) coefs[i] = _coefs[size_t(i)];
    return int(_coefs.size());
  }
  virtual double getBinCenterFrequency(int bin) const { return _centres[size_t(bin)]; }
  int getNBins() const { return _nBins; }
  const double *band(int b) const { return &_coefs[size_t(b) * size_t(_K)]; }
 private:
  int _nBins, _K;
  std::vector<double> _coefs, _centres;
};

class SubBandSTFT { public: typedef enum { LINEAR = 0, MEL = 1 } BandScale; };

class SubBandSTFTAnalysis : public STFTAnalysis {
 public:
  SubBandSTFTAnalysis(int nbins, int sampleRate, int order, int nchannels, float minFreq, float maxFreq, SubBandSTFT::BandScale)
      : STFTAnalysis(nchannels, order), _filterBank(new FilterBankFFTWLinear(order, nbins, sampleRate, minFreq, maxFreq)), _nSubBands(nbins) {
    for (int c = 0; c < nchannels; ++c) _bandStore.push_back(std::vector<double>(size_t(_analysisLength), 0.0));
  }
  int getNumberOfBins() const { return _nSubBands; }
 protected:
  virtual void processSetup(std::vector<double *> &analysisFrames, int analysisLength, std::vector<double *> &dataChannels, int dataLength) = 0;
  virtual void processOneSubband(std::vector<double *> &analysisFrame, int length, int bin) = 0;
  virtual void processSumamry(std::vector<double *> &analysisFrames, int analysisLength, std::vector<double *> &dataChannels, int dataLength) = 0;
  virtual void processParametrisation(std::vector<double *> &analysisFrames, int analysisLength, std::vector<double *> &dataChannels, int dataLength) {
    processSetup(analysisFrames, analysisLength, dataChannels, dataLength);
    std::vector<double *> band;
    for (size_t c = 0; c < analysisFrames.size(); ++c) band.push_back(_bandStore[c].data());
    for (int b = 0; b < _nSubBands; ++b) {
      const double *h = _filterBank->band(b);
      for (size_t c = 0; c < analysisFrames.size(); ++c)
        for (int i = 0; i < analysisLength; ++i) band[c][i] = analysisFrames[c][i] * h[i / 2];
      processOneSubband(band, analysisLength, b);
    }
    processSumamry(analysisFrames, analysisLength, dataChannels, dataLength);
  }
  boost::scoped_ptr<FilterBankFFTWLinear> _filterBank;
 private:
  int _nSubBands;
  std::vector<std::vector<double>> _bandStore;
};

// ------------------------------------------------------------------------------------------------
// Particle-filter scaffolding: OUT OF SCOPE (stochastic).  Only what
// the reference's headers need to compile; the "filter" is a deterministic pass-through that
// keeps returning the state it was seeded with.
// ------------------------------------------------------------------------------------------------
template <class T> class ParticleSet {
 public:
  ParticleSet() {}
  explicit ParticleSet(int n) : _v(size_t(n)) {}
  size_t size() const { return _v.size(); }
  T *get() { return _v.data(); }
  const T *get() const { return _v.data(); }
  T &at(size_t i) { return _v.at(i); }
  const T &at(size_t i) const { return _v.at(i); }
 private:
  std::vector<T> _v;
};
template <class T> using BasicParticleSet = ParticleSet<T>;
template <class PS> class IObservationModel {
 public:
  virtual ~IObservationModel() {}
  virtual PS getWeights(const PS &particles) const = 0;
  virtual void updateModel() = 0;
};
template <class T> class PredictionModel {
 public:
  PredictionModel(T initialState, T initialVelocity) : _state(initialState), _velocity(initialVelocity) {}
  virtual ~PredictionModel() {}
  virtual void update(ParticleSet<T> &) {}
 protected:
  T _state, _velocity;
};
template <class T, class W> class ResamplingModel { public: virtual ~ResamplingModel() {} };
template <class T, class W> inline ResamplingModel<T, W> *make_resampling_model() { return new ResamplingModel<T, W>(); }
template <class T, class Id, class PS, class WS> class ParticleFilter {
 public:
  ParticleFilter(T initial, int nparticles, Id id, std::pair<T, T>, IObservationModel<PS> *, PredictionModel<T> *, ResamplingModel<T, T> *)
      : _state(initial), _id(id), _particles(nparticles), _weights(nparticles) {
    for (int i = 0; i < nparticles; ++i) { _particles.at(size_t(i)) = initial; _weights.at(size_t(i)) = T(1) / T(nparticles); }
  }
  T updateFilter() { return _state; }
  Id getId() const { return _id; }
  PS getParticles() const { return _particles; }
  WS getWeights() const { return _weights; }
 private:
  T _state; Id _id; PS _particles; WS _weights;
};

// ArrayModules.h bases (facades; out of scope, declared so the header parses)
class SignalAnalyser { public: virtual ~SignalAnalyser() {} };
class SignalProcessor { public: virtual ~SignalProcessor() {} };

}  // namespace dsp

#endif

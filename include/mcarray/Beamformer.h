/* mca::Beamformer — frame-level delay-and-sum beamformer with the reference's interface
 * (include/mcarray/Beamformer.h:39-49, src/mcarray/Beamformer.cpp:51-71):
 *   Y[k] = (1/M) sum_c X_c[k] exp(j k phi_c),  phi_c = 2 pi fs/N/c x_c cos(DOA + pi/2)   (x coordinate only, :59)
 * Frames are CCS buffers of fftCCSLength = N + 2 doubles (K interleaved re/im pairs).  One frame per call goes to the device
 * and back, so this class is for API parity and tests; the streaming processors keep the spectra on the GPU. */
#ifndef MCARRAY_B200_BEAMFORMER_H
#define MCARRAY_B200_BEAMFORMER_H

#include <mcarray/ArrayDescription.h>
#include <mcarray/mcadefs.h>
#include <mcarray/mcarray_exception.h>
#include <mcarray_b200.h>

#include <string>
#include <vector>

namespace mca {

namespace detail {
/** RAII device buffer over the C ABI */
class DeviceBuffer {
 public:
  DeviceBuffer() : _p(NULL) {}
  explicit DeviceBuffer(long long bytes) : _p(mcag_dev_alloc(bytes)) { if (!_p) throw MCArrayException(std::string("mcarray_b200: ") + mcag_last_error()); }
  ~DeviceBuffer() { mcag_dev_free(_p); }
  void alloc(long long bytes) { mcag_dev_free(_p); _p = mcag_dev_alloc(bytes); if (!_p) throw MCArrayException(std::string("mcarray_b200: ") + mcag_last_error()); }
  void *get() const { return _p; }
 private:
  DeviceBuffer(const DeviceBuffer &);
  DeviceBuffer &operator=(const DeviceBuffer &);
  void *_p;
};
inline void ok(int rc) { if (rc != MCAG_OK) throw MCArrayException(std::string("mcarray_b200: ") + mcag_last_error()); }

/** CCS doubles [M][N+2] -> device float2 rows of pitch N/2+2 */
inline void upload_frames(const SignalVector &frames, int M, int N, std::vector<float> &stage, DeviceBuffer &d_spec) {
  const int KP = N / 2 + 2;
  stage.assign(size_t(M) * KP * 2, 0.f);
  for (int c = 0; c < M; ++c)
    for (int i = 0; i < N + 2; ++i) stage[size_t(c) * KP * 2 + i] = float(frames[c][i]);
  ok(mcag_dev_upload(d_spec.get(), stage.data(), (long long)stage.size() * 4));
}
}  // namespace detail

class Beamformer {
 public:
  Beamformer(int sampleRate, ArrayDescription microphonePositions, int fftCCSLength, unsigned int nchannels)
      : _sampleRate(sampleRate), _N(fftCCSLength - 2), _M(int(nchannels)), _xyz(microphonePositions.xyz()),
        _d_spec((long long)nchannels * (fftCCSLength / 2 + 1) * 8), _d_out((long long)(fftCCSLength / 2 + 1) * 8), _d_fx((long long)nchannels * 8) {
    if (int(microphonePositions.size()) < _M) throw MCArrayException("Beamformer: fewer microphone positions than channels");
  }
  virtual ~Beamformer() {}

  void processFrame(SignalVector &inputAnalysisFrames, SignalPtr outputFrame, double DOA) {
    const int KP = _N / 2 + 2;
    detail::upload_frames(inputAnalysisFrames, _M, _N, _stage, _d_spec);
    std::vector<double> turns(_M);
    mcag_geom_steer_turns(_xyz.data(), _M, _sampleRate, _N, &DOA, 1, turns.data());
    detail::ok(mcag_k_phase_fx(turns.data(), _M, static_cast<uint64_t *>(_d_fx.get()), NULL));
    detail::ok(mcag_k_ds_fan(_d_spec.get(), 1, 1, _M, _N, static_cast<const uint64_t *>(_d_fx.get()), 1, _d_out.get(), NULL));
    std::vector<float> y(size_t(KP) * 2);
    detail::ok(mcag_dev_download(y.data(), _d_out.get(), (long long)y.size() * 4));
    for (int i = 0; i < _N + 2; ++i) outputFrame[i] = y[i];
  }

 private:
  int _sampleRate, _N, _M;
  std::vector<double> _xyz;
  std::vector<float> _stage;
  detail::DeviceBuffer _d_spec, _d_out, _d_fx;
};

}  // namespace mca

#endif

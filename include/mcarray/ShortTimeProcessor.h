/* mca::ShortTimeProcessor — the DSPONE-shaped base of every GPU processor: process() overloads, frame / latency getters
 * and callback delivery over one C-ABI handle (include/mcarray_b200.h).  It stands where dsp::ShortTimeProcess /
 * dsp::STFT / dsp::STFTAnalysis stand in the reference class hierarchy
 * (include/mcarray/SourceSeparationAndLocalisation.h:42, SourceLocalisation.h:38, BinauralLocalisation.h:188,
 * FastBinauralMasking.h:60) and keeps their calling convention:
 *   int n_out = processor.process(in, nsamples, out, outbuffersize)      src/programs/mcabeamf.cpp:112, test_mcarray.cpp:937
 *   processor.process(in, nsamples)                                      analysis only, test_mcarray.cpp:618,622
 * planar buffers (one pointer per channel), any chunk length, leftover samples buffered inside, outputs sized by the caller
 * as nsamples + getMaxLatency() (mcabeamf.cpp:85).  In the GPU build processParametrisation() is not a per-frame host hook:
 * the whole frame chain of a call runs on the device and results come back once per call. */
#ifndef MCARRAY_B200_SHORTTIMEPROCESSOR_H
#define MCARRAY_B200_SHORTTIMEPROCESSOR_H

#include <mcarray/mcadefs.h>
#include <mcarray/mcarray_exception.h>
#include <mcarray_b200.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

namespace mca {

class ShortTimeProcessor {
 public:
  virtual ~ShortTimeProcessor() { if (_handle) mcag_destroy(_handle); }

  /** dsp::ShortTimeProcess::calculateOrderFromSampleRate: log2 of the frame length for a frame duration in seconds */
  static int calculateOrderFromSampleRate(int sampleRate, double frameRate) {
    int n = mcag_geom_frame_size(sampleRate, frameRate), order = 0;
    while ((1 << order) < n) ++order;
    return order;
  }

  int getFrameSize() const { return _info.frame_size; }
  int getWindowSize() const { return _info.window_size; }
  int getAnalysisLength() const { return _info.analysis_length; }
  int getOneSidedFFTLength() const { return _info.one_sided_length; }
  int getNumberOfChannels() const { return _info.n_channels; }
  int getMaxLatency() const { return _info.max_latency; }
  int getNumberOfStreams() const { return _info.n_streams; }

  // ---- process(): planar host buffers, B*M input pointers (stream-major), B*C output pointers ---------------------
  int process(const std::vector<double *> &in, int nsamples, const std::vector<double *> &out, int outbuffersize) { return run<double>(in, nsamples, &out, outbuffersize, &mcag_process_f64); }
  int process(const std::vector<float *> &in, int nsamples, const std::vector<float *> &out, int outbuffersize) { return run<float>(in, nsamples, &out, outbuffersize, &mcag_process_f32); }
  int process(const std::vector<int16_t *> &in, int nsamples, const std::vector<int16_t *> &out, int outbuffersize) { return run<int16_t>(in, nsamples, &out, outbuffersize, &mcag_process_s16); }
  int process(const std::vector<double *> &in, int nsamples) { return run<double>(in, nsamples, NULL, 0, &mcag_process_f64); }
  int process(const std::vector<float *> &in, int nsamples) { return run<float>(in, nsamples, NULL, 0, &mcag_process_f32); }
  int process(const std::vector<int16_t *> &in, int nsamples) { return run<int16_t>(in, nsamples, NULL, 0, &mcag_process_s16); }
  int process(const SignalVector &in, int nsamples) { return process(raw(in), nsamples); }
  int process(const SignalVector &in, int nsamples, const SignalVector &out, int outbuffersize) { return process(raw(in), nsamples, raw(out), outbuffersize); }

  /** C handle for device-resident use (mcag_process_device_f32, mcag_device_ptr, mcag_fetch) */
  mcag_proc handle() const { return _handle; }
  void reset() { check(mcag_reset(_handle)); }

 protected:
  ShortTimeProcessor() : _handle(NULL) { _info = mcag_info(); }
  ShortTimeProcessor(const ShortTimeProcessor &);              // one handle per object, like the reference: not copyable
  ShortTimeProcessor &operator=(const ShortTimeProcessor &);

  static void check(int rc) { if (rc != MCAG_OK) throw MCArrayException(std::string("mcarray_b200: ") + mcag_last_error()); }
  void create(const mcag_config &cfg) {
    check(mcag_create(&cfg, &_handle));
    check(mcag_get_info(_handle, &_info));
  }
  /** results of the frames the last C call completed; processors with callbacks override it */
  virtual void deliver(int /*frames*/) {}

  mcag_proc _handle;
  mcag_info _info;

 private:
  static std::vector<double *> raw(const SignalVector &v) {
    std::vector<double *> r(v.size());
    for (size_t i = 0; i < v.size(); ++i) r[i] = v[i].get();
    return r;
  }
  template <class T, class Fn>
  int run(const std::vector<T *> &in, int nsamples, const std::vector<T *> *out, int outbuffersize, Fn fn) {
    const size_t rows = size_t(_info.n_streams) * _info.n_channels, orows = size_t(_info.n_streams) * _info.n_out_channels;
    if (in.size() < rows) throw MCArrayException("process: expected one input pointer per channel");
    if (out && orows && out->size() < orows) throw MCArrayException("process: expected one output pointer per channel");
    const bool audio = out && orows;
    // any chunk length: feed at most max_frames_per_call frames' worth of samples per C call
    const int step = std::max(_info.hop, (_info.max_frames_per_call - 1) * _info.hop);
    int done = 0, written = 0;
    std::vector<const T *> ip(rows);
    std::vector<T *> op(audio ? orows : 0);
    while (done < nsamples || (nsamples == 0 && done == 0)) {
      const int n = std::min(step, nsamples - done);
      for (size_t r = 0; r < rows; ++r) ip[r] = in[r] + done;
      for (size_t r = 0; r < op.size(); ++r) op[r] = (*out)[r] + written;
      int nout = 0;
      check(fn(_handle, ip.data(), n, audio ? op.data() : NULL, audio ? outbuffersize - written : 0, &nout));
      deliver(mcag_frames_done(_handle));
      written += nout;
      done += n;
      if (nsamples == 0) break;
    }
    return written;
  }
};

}  // namespace mca

#endif

/* mca::SteeringBeamforming — frame-level SRP / GCC-PHAT localiser with the reference's interface
 * (include/mcarray/SteeringBeamforming.h:43-54, src/mcarray/SteeringBeamforming.cpp:34-195): per pair the PHAT correlation at
 * the 37 grid delays, pair sum with 0.8 / 0.2 temporal smoothing (state carried between calls), derivative + median peak
 * pick of numOfSources DOAs (radians) and their weights.  wienerCoefs is accepted and ignored, as in the reference
 * (computeCorrelations never reads it).  One frame per call crosses PCIe: API parity and tests only. */
#ifndef MCARRAY_B200_STEERINGBEAMFORMING_H
#define MCARRAY_B200_STEERINGBEAMFORMING_H

#include <mcarray/Beamformer.h>

namespace mca {

class SteeringBeamforming {
 public:
  SteeringBeamforming(int sampleRate, ArrayDescription microphonePositions, int fftCCSLength, unsigned int nchannels)
      : _sampleRate(sampleRate), _N(fftCCSLength - 2), _M(int(nchannels)), _P(_M * (_M - 1) / 2), _doaStep(float(5 * M_PI / 180)),
        _D(mcag_geom_grid_size(_doaStep)) {
    if (int(microphonePositions.size()) < _M || _M < 2) throw MCArrayException("SteeringBeamforming: need one position per channel and at least two channels");
    const std::vector<double> xyz = microphonePositions.xyz();
    std::vector<double> tau(size_t(_P) * _D);
    mcag_geom_pair_tau_reference(xyz.data(), _M, sampleRate, _doaStep, tau.data());
    for (size_t i = 0; i < tau.size(); ++i) tau[i] /= double(_N);   // turns per bin: exp(+j 2 pi k tau / N)
    _d_fx.alloc((long long)tau.size() * 8);
    detail::ok(mcag_k_phase_fx(tau.data(), (long long)tau.size(), static_cast<uint64_t *>(_d_fx.get()), NULL));
    _d_spec.alloc((long long)_M * (_N / 2 + 2) * 8);
    _d_corr.alloc((long long)_P * _D * 4); _d_esum.alloc(_D * 4); _d_energy.alloc(_D * 4); _d_state.alloc(_D * 4);
  }
  virtual ~SteeringBeamforming() {}

  void processFrame(const SignalVector &analysisFrames, SignalPtr DOA, SignalPtr prob, int numOfSources, SignalVector & /*wienerCoefs*/) {
    detail::upload_frames(analysisFrames, _M, _N, _stage, _d_spec);
    const float a = 0.8f, b = 1.0f - 0.8f;   // _energyMemoryFactor, SteeringBeamforming.h:70
    if (_d_sel_cap < numOfSources) { _d_idx.alloc(numOfSources * 4); _d_prob.alloc(numOfSources * 4); _d_sel_cap = numOfSources; }
    float *corr = static_cast<float *>(_d_corr.get()), *esum = static_cast<float *>(_d_esum.get()), *energy = static_cast<float *>(_d_energy.get());
    detail::ok(mcag_k_gcc_tau(_d_spec.get(), 1, 1, _M, _N, static_cast<const uint64_t *>(_d_fx.get()), _D, corr, NULL));
    detail::ok(mcag_k_pair_sum(corr, 1, _P, _D, b, esum, NULL));
    detail::ok(mcag_k_energy_scan(esum, 1, 1, _D, a, NULL, static_cast<float *>(_d_state.get()), energy, NULL));
    detail::ok(mcag_k_select_doa(energy, 1, _D, _P, numOfSources, static_cast<int32_t *>(_d_idx.get()), static_cast<float *>(_d_prob.get()), NULL));
    std::vector<int32_t> idx(numOfSources);
    std::vector<float> pr(numOfSources);
    detail::ok(mcag_dev_download(idx.data(), _d_idx.get(), numOfSources * 4));
    detail::ok(mcag_dev_download(pr.data(), _d_prob.get(), numOfSources * 4));
    for (int s = 0; s < numOfSources; ++s) { DOA[s] = mcag_geom_cell_angle(idx[s], _doaStep); prob[s] = pr[s]; }
  }

  /** smoothed energy map of the last frame, [37] */
  std::vector<float> energyInDOA() const {
    std::vector<float> e(_D);
    detail::ok(mcag_dev_download(e.data(), _d_energy.get(), _D * 4));
    return e;
  }

 private:
  int _sampleRate, _N, _M, _P;
  const float _doaStep;
  const int _D;
  int _d_sel_cap = 0;
  std::vector<float> _stage;
  detail::DeviceBuffer _d_fx, _d_spec, _d_corr, _d_esum, _d_energy, _d_state, _d_idx, _d_prob;
};

}  // namespace mca

#endif

/* mca::LocalisingProcessor — callback plumbing shared by the localisation processors: the role SoundLocalisationImpl plays
 * in the reference (include/mcarray/SoundLocalisationImpl.h:45-60, src/mcarray/SoundLocalisationImpl.cpp:45-53).  The
 * callback pointer is borrowed, never freed.  After every C call the selected grid cells, peak weights, frame powers and
 * gate flags are fetched and setDOA() fires once per above-floor frame, in frame order, with angles in degrees
 * (BeamformingSeparationAndLocalisation.cpp:89-95, toDegrees microhponeArrayHelpers.cpp:91-98). */
#ifndef MCARRAY_B200_SOUNDLOCALISATIONIMPL_H
#define MCARRAY_B200_SOUNDLOCALISATIONIMPL_H

#include <mcarray/ShortTimeProcessor.h>
#include <mcarray/SoundLocalisationCallback.h>

namespace mca {

class LocalisingProcessor : public ShortTimeProcessor {
 public:
  void setCallback(LocalisationCallback &callback) { _callback = &callback; }
  void setCallback(LocalisationCallback *callback) { _callback = callback; }

  /** grid cell -> DOA in radians (doaIdx2angle, microhponeArrayHelpers.cpp:117-120); cell n_dirs is the initial DOA of 0 */
  double cellAngle(int cell) const { return (cell < 0 || cell >= _info.n_dirs) ? 0.0 : mcag_geom_cell_angle(cell, _doaStep); }

 protected:
  LocalisingProcessor() : _callback(NULL), _doaStep(0), _cellsPerFrame(1) {}

  virtual void deliver(int frames) {
    if (!_callback || frames <= 0) return;
    const int B = _info.n_streams, S = _cellsPerFrame;
    _cells.resize(size_t(B) * frames * S); _prob.assign(size_t(B) * frames * S, 1.0f); _power.resize(size_t(B) * frames); _active.resize(size_t(B) * frames);
    check(mcag_fetch(_handle, MCAG_OUT_CELL, _cells.data(), (long long)_cells.size() * 4));
    if (_hasProb) check(mcag_fetch(_handle, MCAG_OUT_PROB, _prob.data(), (long long)_prob.size() * 4));
    check(mcag_fetch(_handle, MCAG_OUT_POWER_DB, _power.data(), (long long)_power.size() * 4));
    check(mcag_fetch(_handle, MCAG_OUT_ACTIVE, _active.data(), (long long)_active.size()));
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < frames; ++t) {
        const size_t ft = size_t(b) * frames + t;
        if (!_active[ft]) continue;
        SignalPtr doa(new BaseType[S]), prob(new BaseType[S]);
        for (int s = 0; s < S; ++s) {
          doa[s] = cellAngle(_cells[ft * S + s]) * 180.0 / M_PI;
          prob[s] = _prob[ft * S + s];
        }
        _callback->setDOA(doa, prob, _power[ft], S);
      }
  }

  LocalisationCallback *_callback;
  float _doaStep;
  int _cellsPerFrame;
  bool _hasProb = true;

 private:
  std::vector<int32_t> _cells;
  std::vector<float> _prob, _power;
  std::vector<unsigned char> _active;
};

}  // namespace mca

#endif

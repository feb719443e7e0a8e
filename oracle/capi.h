/* TEST INFRASTRUCTURE ONLY (oracle).  C entry points shared by
 *   oracle/_build/liboracle.so   (prefix orc_: the restatement, oracle/restated.hpp) and
 *   oracle/_ref/libmcarray_ref.so (prefix ref_: the reference's own sources + the DSPONE/WIPP stand-in).
 * Both libraries export the SAME signatures for the first block so tests can diff them 1:1.
 * All arrays are row-major doubles unless noted; "ccs" = N+2 doubles per spectrum.
 */
#ifndef ORACLE_CAPI_H
#define ORACLE_CAPI_H

#ifndef ORC_PREFIX
#define ORC_PREFIX orc_
#endif
#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC_FN(name) ORC_CAT(ORC_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

/* helpers — microhponeArrayHelpers.cpp:46-72,110-120 */
double ORC_FN(doa_idx_to_angle)(int idx, float doa_step);
double ORC_FN(angle_to_doa_idx)(float angle, float doa_step);
double ORC_FN(doa_to_delay_samples)(float doa, float mic_dist, int fs);
/* ArrayDescription — ArrayDescription.cpp:57-94; xyz is [M][3] */
double ORC_FN(array_distance)(const double *xyz, int M, int i, int j);
double ORC_FN(array_max_distance)(const double *xyz, int M);
/* frame size the processors pick: N = 2^calculateOrderFromSampleRate(fs, frame_rate) */
int ORC_FN(frame_size)(int fs, double frame_rate);

/* SourceSeparationAndLocalisation (mcbeam's processor): whole-signal run, fed in `chunk`-sample calls.
 * Returns N (>0) or <0 on error.  Per fired (callback) frame f: fired_frame[f] = frame index,
 * doa_deg/prob [f][S], power[f], energy[f][37] = smoothed un-normalised E, corr_scaled[f][P][37] =
 * (1-0.8f)*Re(GCC) per pair (optional, may be NULL). */
int ORC_FN(ssl_run)(int fs, int M, const double *mic_xyz, int S, int use_floor, int analysis_only,
                    const double *in, int n, int chunk, double *out, int out_cap, int *n_out,
                    int max_frames, int *n_frames, int *n_fired, int *fired_frame,
                    double *doa_deg, double *prob, double *power, double *energy, double *corr_scaled);

/* FreqGCCBinauralLocalisation: curves[f][61] smoothed correlation, idx[f] = its first argmax.
 * noise_preestimated != 0 starts with _noiseEstimated = true / _powerFloor = 0: the only way to run the
 * reference build, because the std::vector<double*> setPowerFloor wrapper it would otherwise call during
 * the first 3 s falls off the end of a non-void function (BinauralLocalisation.cpp:376-385; g++ traps). */
int ORC_FN(freqgcc_run)(int fs, double mic_dist, int use_floor, int noise_preestimated, const double *in, int n, int chunk,
                        int max_frames, int *n_frames, int *n_fired, int *fired_frame,
                        double *curves, int *idx, double *power);
/* setProbability on a given curve — BinauralLocalisation.cpp:569-631 */
void ORC_FN(freqgcc_probability)(int fs, double mic_dist, const double *curve, const double *doas, double *probs, int size);

/* MultibandBinarualLocalisation (MultibandBinarualLocalisation.cpp:52-259): per fired frame f the cell = arg-max of the energy-weighted
 * DOA histogram, prob, power, doa_deg (the published _currentDOA), hist[f][D] = _energyInDOA, band_cells[f][nbins] = per-band arg-max
 * cells.  noise_preestimated as for freqgcc_run (the same falling-off-the-end setPowerFloor wrapper, :115-124). Returns N; *n_dirs = D. */
int ORC_FN(multiband_run)(int fs, double mic_dist, int nbins, int use_floor, int noise_preestimated, const double *in, int n, int chunk,
                          int max_frames, int *n_frames, int *n_fired, int *fired_frame, int *n_dirs,
                          int *cell, double *prob, double *power, double *doa_deg, double *hist, int *band_cells);

/* FastBinauralMasking: Q[f][45] = short-time power after frame f; spectra_out[f][2][ccs] masked spectra (optional). */
int ORC_FN(mask_run)(int fs, double mic_dist, float lo, float hi, int method, int alg,
                     const double *in, int n, int chunk, double *out, int out_cap, int *n_out,
                     int max_frames, int *n_frames, double *Q, double *spectra_out);

/* Beamformer::processFrame on one frame: frames[M][ccs] -> out[ccs] */
int ORC_FN(beamformer_frame)(int fs, int M, const double *mic_xyz, int ccs_len, const double *frames, double doa, double *out);
/* SteeringBeamforming::processFrame on T consecutive frames [T][M][ccs] */
int ORC_FN(steering_frames)(int fs, int M, const double *mic_xyz, int ccs_len, int S, const double *frames, int T,
                            double *doa_rad, double *prob, double *energy);

#ifdef __cplusplus
}
#endif
#endif

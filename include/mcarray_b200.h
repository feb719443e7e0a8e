/* mcarray_b200 — C ABI of the B200 (sm_100a) implementation of mcarray's frame-based multichannel hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, status codes instead of exceptions.  The C++
 * classes under include/mcarray/ (same names and constructor signatures as the reference's include/mcarray/) are
 * thin wrappers over it; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * One processor handle <-> one CUDA device <-> one CUDA stream; handles are stateful and not re-entrant, exactly like
 * the reference objects they replace (one object per array stream, no locks: SURVEY.md §8b "Threading").  A handle can
 * carry B independent array streams that are processed in lock step (B = 1 reproduces one reference object).
 *
 * What each entry point replaces in the reference (/root/reference):
 *   mcag_create / mcag_destroy        constructors / destructors of SourceSeparationAndLocalisation
 *                                     (include/mcarray/SourceSeparationAndLocalisation.h:42-56), SourceLocalisation
 *                                     (SourceLocalisation.h:38-52), FreqGCCBinauralLocalisation (BinauralLocalisation.h:188-192),
 *                                     FastBinauralMasking (FastBinauralMasking.h:71-104)
 *   mcag_process_f32/_f64/_s16        dsp::ShortTimeProcess::process(in, n, out, outsize) as called at
 *                                     src/programs/mcabeamf.cpp:112 and test/test_mcarray.cpp:618,622,937,1023
 *   mcag_get_info                     getFrameSize / getWindowSize / getAnalysisLength / getMaxLatency /
 *                                     getNumberOfChannels (mcabeamf.cpp:85, test_mcarray.cpp:596,843-844,896)
 *   mcag_fetch_*                      LocalisationCallback::setDOA deliveries (SoundLocalisationCallback.h:53;
 *                                     BeamformingSeparationAndLocalisation.cpp:91-94, BinauralLocalisation.cpp:521)
 *   mcag_k_*                          the per-frame algorithms themselves on device buffers:
 *                                     SteeringBeamforming.cpp:96-195, Beamformer.cpp:51-71,
 *                                     FastBinauralMasking.cpp:126-538, DSPONE's STFT and GeneralisedCrossCorrelation
 */
#ifndef MCARRAY_B200_H
#define MCARRAY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcag_proc_s *mcag_proc;

enum {
  MCAG_OK = 0,
  MCAG_ERR_INVALID = 1,   /* bad argument / unsupported configuration (the C++ wrappers rethrow MCArrayException) */
  MCAG_ERR_CUDA = 2,      /* CUDA runtime error; see mcag_last_error() */
  MCAG_ERR_NOMEM = 3,
  MCAG_ERR_CAPACITY = 4   /* more frames in one call than max_frames_per_call, or output buffer too small */
};

/* processor kinds */
enum {
  MCAG_KIND_SSL = 0,      /* SourceSeparationAndLocalisation: STFT -> SRP(GCC-PHAT, tau grid) -> selectDOA -> DS -> OLA */
  MCAG_KIND_SL = 1,       /* SourceLocalisation: the same without separation / synthesis */
  MCAG_KIND_FREQGCC = 2,  /* FreqGCCBinauralLocalisation: 2-mic GCC-PHAT curve, smoothing, argmax cell */
  MCAG_KIND_MASK = 3,     /* FastBinauralMasking: STFT -> 45-band spatial/temporal mask -> OLA */
  MCAG_KIND_TDOA = 4,     /* integer-lag GCC-PHAT on all pairs (BASELINE config 2) */
  MCAG_KIND_DSFAN = 5,    /* delay-and-sum to a fan of D azimuths, spectra out (BASELINE config 3) */
  MCAG_KIND_SRP = 6,      /* SRP-PHAT map over a large grid, channel form (BASELINE config 4) */
  MCAG_KIND_MULTIBAND = 7 /* MultibandBinarualLocalisation: per sub-band GCC-PHAT curves (0.4 memory) + energy-weighted DOA histogram */
};

/* result arrays (mcag_fetch_* / mcag_device_ptr); shapes use T = frames of the last process call */
enum {
  MCAG_OUT_SPECTRA = 0,   /* float2 [B][T][M][N/2+2]  analysis spectra (pad bin zero)            */
  MCAG_OUT_POWER_DB = 1,  /* float  [B][T]            dsp::SignalPower::FFTLogPower of the frame  */
  MCAG_OUT_CORR = 2,      /* float  [B][T][P][D]      Re GCC-PHAT per pair on the tau grid        */
  MCAG_OUT_ENERGY = 3,    /* float  [B][T][D]         smoothed energy map (SSL/SL/SRP)            */
  MCAG_OUT_CELL = 4,      /* int32  [B][T][S]         selected grid cells (SSL/SL), [B][T] (FREQGCC argmax, SRP argmax) */
  MCAG_OUT_PROB = 5,      /* float  [B][T][S]         peak weights ("prob")                       */
  MCAG_OUT_LAGS = 6,      /* int32  [B][T][P]         integer TDOA lags (TDOA)                    */
  MCAG_OUT_CURVES = 7,    /* float  [B][T][P][2L+1] (TDOA) or [B][T][D] smoothed curve (FREQGCC)  */
  MCAG_OUT_ACTIVE = 8,    /* uint8  [B][T]            1 where the power gate let the frame through */
  MCAG_OUT_BEAMS = 9,     /* float2 [B][T][C][mcag_info::beams_pitch]  beamformed / masked spectra (SSL, MASK: pitch N/2+2; DSFAN: C = D, rows padded) */
  MCAG_OUT_MASK_Q = 10,   /* float  [B][T][nb]        short-time band power after each frame (MASK) */
  MCAG_OUT_MASK_DEC = 11, /* uint8  [B][T][nb]        2 = spatial mask, 1 = temporal mask, 0 = pass (MASK) */
  MCAG_OUT_BAND_CELL = 12,/* int32  [B][T][nb]        arg-max cell of each sub-band curve (MULTIBAND); its MCAG_OUT_CURVES is [B][T][nb][D],
                             MCAG_OUT_ENERGY the energy-weighted histogram [B][T][D], MCAG_OUT_CELL / _PROB [B][T] */
  MCAG_OUT_TRACK_DOA = 13 /* double [B][T]            FREQGCC with doa_tracker: _currentDOA (rad) after each frame, the deterministic
                             `#else` branch of BinauralLocalisation.cpp:501-504; its MCAG_OUT_PROB [B][T] is setProbability (:454,569-631) */
};

enum {                    /* emit flags: keep optional intermediates of the last call fetchable */
  MCAG_EMIT_CORR = 1, MCAG_EMIT_CURVES = 2,
  MCAG_EMIT_SPECTRA = 4,  /* TDOA, MASK: their fused kernels keep the spectra on chip unless this is set (every other kind always has them); MASK
                             then runs the staged stft / stats / scan / apply / istft kernels and MCAG_OUT_SPECTRA / _BEAMS become fetchable */
  MCAG_EMIT_MASK_TRACE = 8 /* MASK: keep MCAG_OUT_MASK_Q / MCAG_OUT_MASK_DEC of the last call (always kept by the staged path) */
};

typedef struct {
  int kind;               /* MCAG_KIND_* */
  int device;             /* CUDA device ordinal */
  int sample_rate;
  int frame_size;         /* N in {256, 512, 1024, 2048} */
  int hop;                /* window shift; N/2 is the DSPONE convention */
  int n_channels;         /* M microphones per array stream */
  int n_streams;          /* B independent array streams processed in lock step */
  int max_frames_per_call;/* workspace sizing: a process call may complete at most this many frames per stream */
  const double *window;   /* [N] analysis = synthesis window, NULL = sqrt of the periodic Hann */
  int emit;               /* MCAG_EMIT_* */

  /* direction / delay grid */
  int n_dirs;             /* D */
  const double *pair_tau; /* [P][D] pair delays in samples, pairs i<j lexicographic (SSL, SL, FREQGCC) */
  const double *mic_tau;  /* [M][D] per-microphone advances in samples (SRP channel form)              */
  const double *steer_turns; /* [D][M] phase increment per bin in turns, Beamformer.cpp:59 / (2 pi) (SSL, DSFAN) */
  int n_sources;          /* S */
  float energy_memory;    /* 0.8f: SteeringBeamforming.h:70 */
  float corr_memory;      /* 0.8f: BinauralLocalisation.h:198 */
  int use_power_floor;    /* SoundLocalisationImpl power gate */
  float noise_margin_db;  /* 3 (BeamformingSeparationAndLocalistaion.h:52) or 6 (BinauralLocalisation.h:197) */
  float floor_seconds;    /* 3: SoundLocalisationImpl.h:77 */
  int floor_ccs_power;    /* 0: FFTPower*(N) accumulation (BSAL.cpp:58); 1: mean-square of the CCS buffer + 1e-10 (BinauralLocalisation.cpp:390-391) */
  int noise_preestimated; /* start with the floor already "estimated" at 0 (see oracle/CONVENTIONS.md) */

  int max_lag;            /* TDOA: lags -max_lag..max_lag */

  /* masking */
  int mask_method;        /* FACTOR=0 RELATIVE=1 FULL=3 NOISY=4 NOTHING=5 (ArrayModules.h:81) */
  int mask_alg;           /* BOTH=0 SPATIAL=1 TEMPORAL=2 (ArrayModules.h:89) */
  int n_bands;            /* 45 (FastBinauralMasking.h:111) */
  const double *band_coefs;      /* [n_bands][N/2+1] real filter-bank magnitudes */
  const double *band_thresholds; /* [n_bands] cos(w_b d sin(phi)/c) (FastBinauralMasking.cpp:342-366) */

  /* SSL / SL: how the pair sum of SteeringBeamforming::computeCorrelations (SteeringBeamforming.cpp:104-144) is evaluated.
   * 0 auto: the channel form on the tensor cores (sum_{i<j} Re G_ij e^{jw tau_ij} = (|sum_m U_m e^{-jw tau_m}|^2 - nz)/2)
   *         when the pair delays are consistent with per-microphone delays (tau_ij = tau_j - tau_i within 2e-5 samples, true for
   *         the reference's ascending linear arrays), M is 16/32/48/64 and MCAG_EMIT_CORR is not set; the pair form otherwise.
   * 1 pair form always (pair by pair in the reference's order; MCAG_OUT_CORR is only produced by this form).
   * 2 channel form required: mcag_create fails if the geometry / channel count does not allow it. */
  int srp_form;

  /* FREQGCC: deterministic DOA tracker replacing the (stochastic, out-of-scope) particle filter: the `#else` branch of
   * USE_PARTICLE_FILTER, _currentDOA = m _currentDOA + (1-m) DOA with m = _doaMemoryFactor (0 -> doa_memory after a voiced frame, 0 after
   * 3 s of silence; BinauralLocalisation.cpp:502-504,523-524,528-561) and setProbability (:569-631).  0 = off (arg-max cell only). */
  int doa_tracker;
  float doa_memory;       /* 0.6f: _maxDoaMemoryFactor, BinauralLocalisation.h:199 */

  /* DSFAN as a filter-and-sum beamformer: per-bin complex weights [D][M][N/2+1] (re, im interleaved) used instead of the delay phasors
   * of steer_turns, Y[d][k] = 1/M sum_c X_c[k] W[d][c][k].  NULL = delay-and-sum (Beamformer.cpp:51-71). */
  const double *fs_weights;
} mcag_config;

typedef struct {
  int frame_size, window_size, hop, analysis_length, one_sided_length, n_channels, n_streams, max_latency;
  int n_dirs, n_pairs, n_sources, n_out_channels, spectrum_pitch, max_frames_per_call;
  int srp_form;           /* SSL / SL: 1 = pair form, 2 = channel form (what mcag_create chose); 0 for other kinds */
  int beams_pitch;        /* complex bins per row of MCAG_OUT_BEAMS: spectrum_pitch, except MCAG_KIND_DSFAN whose rows are padded to a multiple
                             of 4 bins (32-byte aligned rows: the fan kernel writes them with 256-bit stores); the pad bins are zero */
} mcag_info;

const char *mcag_last_error(void);
int mcag_version(void);

void mcag_config_init(mcag_config *cfg);                 /* zero + reference defaults */
int mcag_create(const mcag_config *cfg, mcag_proc *out);
void mcag_destroy(mcag_proc p);
int mcag_reset(mcag_proc p);                             /* back to the freshly-constructed state */
int mcag_flush_input(mcag_proc p);                       /* drop buffered input samples (start of a new segment); host-only, no sync */
int mcag_get_info(mcag_proc p, mcag_info *info);
int mcag_synchronize(mcag_proc p);

/* Streaming process.  `in` holds B*M planar host pointers (stream-major), nsamples each; `out` holds B*C planar host
 * pointers with room for out_capacity samples each (C = n_out_channels; NULL for analysis-only kinds).  Returns the
 * number of samples written per output channel in *nsamples_out (hop per completed frame); callers size the output as
 * nsamples + max_latency like mcabeamf.cpp:85.  Synchronous: results are fetchable when it returns. */
int mcag_process_f32(mcag_proc p, const float *const *in, int nsamples, float *const *out, int out_capacity, int *nsamples_out);
int mcag_process_f64(mcag_proc p, const double *const *in, int nsamples, double *const *out, int out_capacity, int *nsamples_out);
int mcag_process_s16(mcag_proc p, const int16_t *const *in, int nsamples, int16_t *const *out, int out_capacity, int *nsamples_out);
/* Same with one contiguous host block per call: in [B*M][in_pitch], out [B*C][out_pitch] (pinned memory recommended). */
int mcag_process_packed_f32(mcag_proc p, const float *in, long long in_pitch, int nsamples, float *out, long long out_pitch, int *nsamples_out);
/* ... and with 16-bit PCM, the sample type of the reference's process(std::vector<int16_t*>&, ...) overload (test_mcarray.cpp:937): half the
 * host<->device bytes of the f32 call; samples are widened on the device, outputs rounded and saturated to int16. */
int mcag_process_packed_s16(mcag_proc p, const int16_t *in, long long in_pitch, int nsamples, int16_t *out, long long out_pitch, int *nsamples_out);
/* Device-resident variant: in / out are device pointers on the handle's device; asynchronous on the handle's stream. */
int mcag_process_device_f32(mcag_proc p, const float *d_in, long long in_pitch, int nsamples, float *d_out, long long out_pitch, int *nsamples_out);

int mcag_frames_done(mcag_proc p);                        /* frames per stream completed by the last process call */
long long mcag_frames_total(mcag_proc p);                 /* since creation / reset */
int mcag_fetch(mcag_proc p, int what, void *dst, long long bytes);   /* device -> host copy of a result array */
const void *mcag_device_ptr(mcag_proc p, int what);      /* zero-copy access for device-side consumers */
void *mcag_stream(mcag_proc p);                           /* cudaStream_t the handle launches on */
long long mcag_kernel_launches(mcag_proc p);              /* kernels launched by this handle since creation */

/* Per-kernel device time, measured with CUDA events recorded on the handle's own stream around each launch group.
 * ms[i] / count[i] accumulate over process calls while enabled; mcag_profile_read synchronises the stream. */
enum {
  MCAG_PROF_STFT = 0, MCAG_PROF_GATE, MCAG_PROF_GCC_TAU, MCAG_PROF_ENERGY, MCAG_PROF_SELECT_DOA, MCAG_PROF_DS_SELECT, MCAG_PROF_ISTFT,
  MCAG_PROF_CURVE_SCAN, MCAG_PROF_TDOA, MCAG_PROF_DS_FAN, MCAG_PROF_SRP, MCAG_PROF_MASK_STATS, MCAG_PROF_MASK_SCAN, MCAG_PROF_MASK_APPLY,
  MCAG_PROF_MASK_FUSED,
  MCAG_PROF_COUNT
};
int mcag_profile_enable(mcag_proc p, int on);
int mcag_profile_read(mcag_proc p, double *ms /* [MCAG_PROF_COUNT] */, long long *count /* [MCAG_PROF_COUNT] */, int reset);
const char *mcag_profile_name(int id);

/* Pinned host memory helpers for the end-to-end path */
void *mcag_host_alloc(long long bytes);
void mcag_host_free(void *ptr);

/* Device memory helpers so plain C / C++ hosts can drive the kernel-level entry points without linking the CUDA runtime
 * themselves (used by the frame-level classes mca::Beamformer / mca::SteeringBeamforming in include/mcarray/). */
void *mcag_dev_alloc(long long bytes);                   /* zero-initialised; NULL on failure */
void mcag_dev_free(void *d_ptr);
int mcag_dev_upload(void *d_dst, const void *h_src, long long bytes);     /* synchronous */
int mcag_dev_download(void *h_dst, const void *d_src, long long bytes);   /* synchronous (after all prior work on the default stream) */

/* ---- kernel-level entry points on DEVICE buffers (what the processors are made of; also used by the parity tests) ----
 * `stream` is a cudaStream_t (NULL = default stream).  Spectra rows have pitch N/2+2 complex bins. */
int mcag_k_twiddle_count(int N);                        /* float2 entries of the FFT tables for frame size N */
int mcag_k_twiddles(int N, void *d_tw /* float2 [mcag_k_twiddle_count(N)] */, void *stream);
int mcag_k_stft(const float *d_x, long long row_pitch, int rows, int M, int T, int N, int hop, const float *d_win, const void *d_tw,
                void *d_spec, float *d_chan_pow, void *stream);
int mcag_k_istft(const void *d_spec, int B, int T, int C_in, int C_out, int N, int hop, const float *d_win, const void *d_tw,
                 const float *d_tail_in, float *d_tail_out, float *d_out, long long out_pitch, void *stream);
int mcag_k_tdoa_lags(const void *d_spec, int B, int T, int M, int N, int max_lag, const void *d_tw, float *d_curves, int32_t *d_lags,
                     float *d_peaks, void *stream);
int mcag_k_phase_fx(const double *h_turns, long long n, uint64_t *d_fx, void *stream);   /* host turns -> device 0.64 fixed point */
/* fused STFT -> GCC-PHAT -> integer-lag argmax on sample rows (what MCAG_KIND_TDOA runs); d_spec / d_chan_pow / d_curves may be NULL */
int mcag_k_stft_tdoa(const float *d_x, long long row_pitch, int B, int T, int M, int N, int hop, int max_lag, const float *d_win, const void *d_tw,
                     void *d_spec, float *d_chan_pow, float *d_curves, int32_t *d_lags, void *stream);
int mcag_k_gcc_tau(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_pair_fx, int D, float *d_corr, void *stream);
/* the same on the tensor cores (tcgen05, 3xTF32): per pair a GEMM over the bins, D <= 64 delays; what the processors run for such grids */
int mcag_k_gcc_tau_tensor(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_pair_fx, int D, float *d_corr, void *stream);
int mcag_k_pair_sum(const float *d_corr, long long BT, int P, int D, float scale, float *d_esum, void *stream);
int mcag_k_energy_scan(const float *d_esum, int B, int T, int D, float a, const unsigned char *d_active, float *d_state, float *d_energy, void *stream);
int mcag_k_select_doa(const float *d_energy, long long BT, int D, int n_pairs, int S, int32_t *d_idx, float *d_prob, void *stream);
/* per-frame (max, first argmax) of an energy-map slice [rows][D] whose first column is global direction d_offset, packed as
 * ordered(E) << 31 | (0x7FFFFFFF - d): an int64 MAX all-reduce over the slices of a sharded grid gives the global arg-max cell */
int mcag_k_argmax_pack(const float *d_map, long long rows, int D, int d_offset, long long *d_packed, void *stream);
int mcag_k_ds_fan(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_steer_fx, int D, void *d_out, void *stream);
/* filter-and-sum: the same fan with loaded weights d_weights float2 [D][M][N/2+2] instead of generated delay phasors */
int mcag_k_fs_fan(const void *d_spec, int B, int T, int M, int N, const void *d_weights, int D, void *d_out, void *stream);
/* the fan on the tensor cores (tcgen05, 3xTF32; M in {16, 32, 48, 64}, other counts run mcag_k_ds_fan): four consecutive bins of a
 * (128-frame, 64-direction) tile stay resident in TMEM, so every (frame, direction) leaves as 32 contiguous bytes of its [B][T][D][K] row.
 * What MCAG_KIND_DSFAN runs for those microphone counts.  The pad bin of d_spec rows must be zero (it is, for spectra of this library). */
int mcag_k_ds_fan_tensor(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_steer_fx, int D, void *d_out, void *stream);
int mcag_k_srp_channel(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_mic_fx, int D, float *d_srp, void *stream);
int mcag_k_srp_tensor(const void *d_spec, int B, int T, int M, int N, const uint64_t *d_mic_fx, int D, float *d_srp, void *stream);

/* ---- host-side constructor math of the reference processors (no GPU work; float typing as the reference helpers) ----
 * microhponeArrayHelpers.cpp:38-72,110-120; SteeringBeamforming.cpp:34-94; Beamformer.cpp:59; FastBinauralMasking.cpp:342-366 */
int mcag_geom_frame_size(int fs, double frame_rate);                       /* 2^calculateOrderFromSampleRate */
int mcag_geom_grid_size(float doa_step);                                   /* round(pi/step)+1 */
double mcag_geom_cell_angle(int idx, float doa_step);                      /* doaIdx2angle */
int mcag_geom_pair_tau_reference(const double *mic_xyz, int M, int fs, float doa_step, double *tau /* [P][D] */);
int mcag_geom_steer_turns_reference(const double *mic_xyz, int M, int fs, int N, float doa_step, double *turns /* [D+1][M] */);
void mcag_geom_steer_turns(const double *mic_xyz, int M, int fs, int N, const double *doas, int D, double *turns /* [D][M] */);
void mcag_geom_mic_tau(const double *mic_xyz, int M, int fs, const double *dirs /* [D][3] */, int D, double *mic_tau /* [M][D] */);
void mcag_geom_pair_tau_from_mic_tau(const double *mic_tau, int M, int D, double *pair_tau /* [P][D] */);
void mcag_geom_mel_bank(int N, int n_bands, int fs, float lo, float hi, double mic_dist, double *H /* [nb][N/2+1] */,
                        double *fc_norm /* [nb] */, double *thresholds /* [nb] */);
/* MultibandBinarualLocalisation constructor (MultibandBinarualLocalisation.cpp:52-101): D = floor(pi/step)+1 delays tau [D] on the
 * 5 degree grid, n_bands linear bands between 100 Hz and c/(2 d) (maxFreqForSpatialAliasing).  Returns D; tau / H may be NULL. */
int mcag_geom_multiband(int fs, double mic_dist, int N, int n_bands, double *tau /* [D] */, double *H /* [nb][N/2+1] */);

#ifdef __cplusplus
}
#endif
#endif

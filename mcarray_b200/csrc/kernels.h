// Internal launcher prototypes (device pointers, stream-ordered).  The C ABI in include/mcarray_b200.h wraps these.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcag {

// stft.cu
int k_stft(const float *x, long long row_pitch, int rows, int M, int T, int N, int hop, const float *win, const float2 *tw, float2 *spec,
           float *chan_pow, float *chan_raw /* plain sum of |X|^2 per (frame, channel), may be NULL */, cudaStream_t st);
int k_istft(const float2 *spec, int B, int T, int C_in, int C_out, int N, int hop, const float *win, const float2 *tw, const float *tail_in,
            float *tail_out, float *out, long long out_pitch, int out_rows /* rows per stream in `out`, >= C_out */, cudaStream_t st);
int k_frame_power(const float2 *spec, long long rows, int N, float *pow, cudaStream_t st);

// gcc.cu
int k_tdoa_lags(const float2 *spec, int B, int T, int M, int N, int max_lag, const float2 *tw, float *curves, int32_t *lags, float *peaks,
                cudaStream_t st);
bool k_stft_tdoa_fits(int M, int N);   // false: k_stft_tdoa needs `spec` and runs stft + the channel-tiled lag kernel
int k_stft_tdoa(const float *x, long long row_pitch, int B, int T, int M, int N, int hop, int max_lag, const float *win, const float2 *tw,
                float2 *spec, float *chan_pow, float *curves, int32_t *lags, cudaStream_t st);
int k_gcc_tau(const float2 *spec, int B, int T, int M, int N, const uint64_t *pair_fx, int D, float *corr, cudaStream_t st);

// gcc_tc.cu (tcgen05): the tau-grid GCC-PHAT of one pair as a GEMM over the bins (D <= 64 delays)
bool k_gcc_tau_tc_supported(int D);
size_t k_gcc_tau_tc_table_bytes(int P, int N);
int k_gcc_tau_tc_build(const uint64_t *pair_fx, int P, int D, int N, float *table, cudaStream_t st);   // once per processor
int k_gcc_tau_tc(const float2 *spec, int B, int T, int M, int N, const float *table, int D, float *corr, cudaStream_t st);

// doa.cu
int k_pair_sum(const float *corr, long long BT, int P, int D, float scale, float *esum, cudaStream_t st);
int k_energy_scan(const float *esum, int B, int T, int D, float a, const unsigned char *active, float *state, float *energy, cudaStream_t st);
int k_argmax_pack(const float *x, long long rows, int D, int d_offset, long long *packed, cudaStream_t st);
int k_select_doa(const float *energy, long long BT, int D, int n_pairs, int S, int32_t *idx, float *prob, cudaStream_t st);
// FreqGCCBinauralLocalisation per-stream state carried across calls (BinauralLocalisation.h:200-213)
struct FgState { float alpha /* _corrMemoryFactor */, dalpha /* _doaMemoryFactor */; int silence /* _silenceFramesCounter */; float prob /* _prob[0] */;
                 double doa /* _currentDOA[0], rad */; };
int k_curve_scan_argmax(const float *corr, int B, int T, int D, float mem, float dmem, int windows_to_decay, float doa_step, int track,
                        const unsigned char *active, const unsigned char *est, float *state, FgState *fg, float *curves, int32_t *idx,
                        double *track_doa, float *track_prob, cudaStream_t st);

// beamform.cu
int k_steer_table(const uint64_t *fx, int DM, int N, float2 *tab, cudaStream_t st);
int k_ds_select(const float2 *spec, int B, int T, int M, int N, const float2 *steer_tab, const int32_t *cells, int S, int C_out, float2 *out,
                cudaStream_t st);
// out_pitch: complex bins per output row (0 = the spectrum pitch N/2 + 2); bins past N/2 are written as zeros
int k_ds_fan(const float2 *spec, int B, int T, int M, int N, const uint64_t *steer_fx, int D, float2 *out, cudaStream_t st, int out_pitch = 0);
int k_fs_fan(const float2 *spec, int B, int T, int M, int N, const float2 *weights /* [D][M][KP] */, int D, float2 *out, cudaStream_t st, int out_pitch = 0);

// srp.cu
int k_srp_channel(const float2 *spec, int B, int T, int M, int N, const uint64_t *mic_fx, int D, float *srp, cudaStream_t st);

// srp_tc.cu (tcgen05)
bool k_srp_tensor_supported(int M);
size_t k_srp_tensor_workspace_bytes(long long BT, int M, int N, int D);
int k_srp_tensor_ws(const float2 *spec, int B, int T, int M, int N, const uint64_t *mic_fx, int D, float *srp, void *workspace, size_t ws_bytes,
                    cudaStream_t st);
int k_srp_tensor(const float2 *spec, int B, int T, int M, int N, const uint64_t *mic_fx, int D, float *srp, cudaStream_t st);

// fan_tc.cu (tcgen05): the delay-and-sum fan, four bins of a (frame, direction) tile resident in TMEM
bool k_ds_fan_tensor_supported(int M);
int k_ds_fan_tensor(const float2 *spec, int B, int T, int M, int N, const uint64_t *steer_fx, int D, float2 *out, cudaStream_t st, int out_pitch = 0);
inline int fan_out_pitch(int N) { return (N / 2 + 2 + 3) & ~3; }   // beams rows of MCAG_KIND_DSFAN: 32-byte aligned (256-bit stores)

// mask.cu
int k_mask_stats(const float2 *spec, long long BT, int N, const float *H2, int nb, float *stats, cudaStream_t st);
int k_mask_scan(const float *stats, int B, int T, int N, int nb, int method, int alg, const float *thresholds, float *Q, float *noise,
                int first_call, float *gains, unsigned char *decisions, float *q_trace, cudaStream_t st);
int k_mask_apply(float2 *spec, long long BT, int N, const float *H, int nb, const float *gains, cudaStream_t st);

// mask_fused.cu: analysis + FastBinauralMasking + synthesis in one kernel, spectra never leave the SM (hop = N/2, N >= 512)
bool k_mask_fused_supported(int N, int hop, int nb);
// tab: compact filter-bank tables {binfo[nb][4], kinfo[K], h2c[n_h2c], hc[n_hc]} (see MfParams)
int k_mask_fused(const float *x, long long row_pitch, int B, int T, int N, int hop, const float *win, const float2 *tw, const int *tab, int n_h2c, int n_hc,
                 int nb, int method, int alg, const float *thresholds, float *Q, float *noise,
                 int first_call, const float *tail_in, float *tail_out, float *out, long long out_pitch, int out_rows, float *chan_pow,
                 unsigned char *decisions, float *q_trace, cudaStream_t st);

// multiband.cu
int k_mb_band(const float2 *spec, long long BT, int N, const float *H, int nb, const float2 *W, int D, float *band_raw, float *band_energy,
              float *floor_pow, cudaStream_t st);
bool k_mb_fused_supported(int D, int max_band_width, int kmin, int kmax, int nb);
int k_mb_fused(const float2 *spec, int B, int T, int N, const float *H, const int *band_lohi, int nb, int max_band_width, int kmin, int kmax,
               const float2 *W, int D, float mem, float *state, float *curves, float *band_energy, float *floor_pow, float *hist,
               int32_t *band_cells, int32_t *raw_cell, float *raw_prob, cudaStream_t st);
int k_mb_scan(const float *raw, int B, int T, int nb, int D, float mem, float *state, float *curves, cudaStream_t st);
int k_mb_summary(const float *curves, const float *band_energy, long long BT, int nb, int D, float *hist, int32_t *band_cells, int32_t *raw_cell,
                 float *raw_prob, cudaStream_t st);
int k_mb_gate(const float *floor_pow, const float *chan_pow, const int32_t *raw_cell, const float *raw_prob, int B, int T, int N, int use_floor,
              float margin_db, int needed, void *gate_state, int32_t *cell_state, float *power_out, unsigned char *active, int32_t *cells, float *prob,
              cudaStream_t st);

}  // namespace mcag

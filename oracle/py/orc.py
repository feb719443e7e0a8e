"""TEST INFRASTRUCTURE ONLY (oracle): ctypes front-end of oracle/_build/liboracle.so (prefix ``orc``,
the restatement) and oracle/_ref/libmcarray_ref.so (prefix ``ref``, the reference's own sources built
against the DSPONE/WIPP stand-in).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import
this module; nothing under mcarray_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIBS = {}

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def build(ref=True):
    """Compile the restatement (and, where /root/reference exists, the reference build)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    if ref and os.path.isdir("/root/reference/src/mcarray"):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)


def lib(prefix="orc"):
    if prefix not in _LIBS:
        path = os.path.join(_HERE, "_build", "liboracle.so") if prefix == "orc" else os.path.join(_HERE, "_ref", "libmcarray_ref.so")
        if not os.path.exists(path):
            if prefix == "orc":
                build(ref=False)
            else:
                raise FileNotFoundError(path)
        L = C.CDLL(path)
        for name in ("doa_idx_to_angle", "angle_to_doa_idx", "doa_to_delay_samples", "array_distance", "array_max_distance"):
            getattr(L, f"{prefix}_{name}").restype = C.c_double
        _LIBS[prefix] = L
    return _LIBS[prefix]


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libmcarray_ref.so"))


# ---------------------------------------------------------------------------------------------
# shared surface (both prefixes)
# ---------------------------------------------------------------------------------------------
def doa_idx_to_angle(idx, step, prefix="orc"):
    return getattr(lib(prefix), f"{prefix}_doa_idx_to_angle")(C.c_int(idx), C.c_float(step))


def angle_to_doa_idx(angle, step, prefix="orc"):
    return getattr(lib(prefix), f"{prefix}_angle_to_doa_idx")(C.c_float(angle), C.c_float(step))


def doa_to_delay_samples(doa, dist, fs, prefix="orc"):
    return getattr(lib(prefix), f"{prefix}_doa_to_delay_samples")(C.c_float(doa), C.c_float(dist), C.c_int(fs))


def array_distance(xyz, i, j, prefix="orc"):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    return getattr(lib(prefix), f"{prefix}_array_distance")(_dp(xyz), C.c_int(len(xyz)), C.c_int(i), C.c_int(j))


def array_max_distance(xyz, prefix="orc"):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    return getattr(lib(prefix), f"{prefix}_array_max_distance")(_dp(xyz), C.c_int(len(xyz)))


def frame_size(fs, frame_rate, prefix="orc"):
    return getattr(lib(prefix), f"{prefix}_frame_size")(C.c_int(fs), C.c_double(frame_rate))


def ssl_run(fs, mic_xyz, S, x, chunk=0, use_floor=False, analysis_only=False, want_corr=False, prefix="orc"):
    """SourceSeparationAndLocalisation over a whole signal x [M][n] (float64)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    xyz = np.ascontiguousarray(mic_xyz, dtype=np.float64)
    M, n = x.shape
    D, P = 37, M * (M - 1) // 2
    N = frame_size(fs, np.float32(0.025), prefix)
    maxf = max(1, n // (N // 2) + 2)
    cap = n + 2 * N
    out = np.zeros((M, cap))
    n_out, n_frames, n_fired = C.c_int(0), C.c_int(0), C.c_int(0)
    fired_frame = np.zeros(maxf, dtype=np.int32)
    doa = np.zeros((maxf, S)); prob = np.zeros((maxf, S)); power = np.zeros(maxf)
    energy = np.zeros((maxf, D))
    corr = np.zeros((maxf, P, D)) if want_corr else None
    r = getattr(lib(prefix), f"{prefix}_ssl_run")(
        C.c_int(fs), C.c_int(M), _dp(xyz), C.c_int(S), C.c_int(int(use_floor)), C.c_int(int(analysis_only)),
        _dp(x), C.c_int(n), C.c_int(chunk), _dp(out), C.c_int(cap), C.byref(n_out),
        C.c_int(maxf), C.byref(n_frames), C.byref(n_fired), _ip(fired_frame), _dp(doa), _dp(prob), _dp(power), _dp(energy), _dp(corr))
    assert r > 0, r
    f = n_fired.value
    return dict(N=r, out=out[:, :n_out.value].copy(), n_frames=n_frames.value, n_fired=f, fired_frame=fired_frame[:f].copy(),
                doa_deg=doa[:f].copy(), prob=prob[:f].copy(), power=power[:f].copy(), energy=energy[:f].copy(),
                corr_scaled=None if corr is None else corr[:f].copy())


def freqgcc_run(fs, mic_dist, x, chunk=0, use_floor=False, noise_preestimated=True, prefix="orc"):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[1]
    N = frame_size(fs, np.float32(0.075), prefix)
    maxf = max(1, n // (N // 2) + 2)
    n_frames, n_fired = C.c_int(0), C.c_int(0)
    fired_frame = np.zeros(maxf, dtype=np.int32); idx = np.zeros(maxf, dtype=np.int32)
    curves = np.zeros((maxf, 61)); power = np.zeros(maxf)
    r = getattr(lib(prefix), f"{prefix}_freqgcc_run")(
        C.c_int(fs), C.c_double(mic_dist), C.c_int(int(use_floor)), C.c_int(int(noise_preestimated)), _dp(x), C.c_int(n), C.c_int(chunk),
        C.c_int(maxf), C.byref(n_frames), C.byref(n_fired), _ip(fired_frame), _dp(curves), _ip(idx), _dp(power))
    f = n_fired.value
    return dict(N=r, n_frames=n_frames.value, n_fired=f, fired_frame=fired_frame[:f].copy(), curves=curves[:f].copy(),
                idx=idx[:f].copy(), power=power[:f].copy())


def freqgcc_track_run(fs, mic_dist, x, chunk=0, use_floor=True, noise_preestimated=False, N=0):
    """FreqGCC with the deterministic tracker (the `#else` branch, BinauralLocalisation.cpp:501-504): per-frame arrays for EVERY frame."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[1]
    Nn = N or frame_size(fs, np.float32(0.075))
    maxf = max(1, n // (Nn // 2) + 2)
    n_frames = C.c_int(0)
    active = np.zeros(maxf, dtype=np.int32); idx = np.zeros(maxf, dtype=np.int32)
    doa = np.zeros(maxf); prob = np.zeros(maxf); power = np.zeros(maxf); curves = np.zeros((maxf, 61))
    r = lib("orc").orc_freqgcc_track_run(C.c_int(fs), C.c_double(mic_dist), C.c_int(N), C.c_int(int(use_floor)), C.c_int(int(noise_preestimated)),
                                        _dp(x), C.c_int(n), C.c_int(chunk), C.c_int(maxf), C.byref(n_frames), _ip(active), _dp(doa), _dp(prob),
                                        _ip(idx), _dp(curves), _dp(power))
    f = n_frames.value
    return dict(N=r, n_frames=f, active=active[:f].copy(), doa_rad=doa[:f].copy(), prob=prob[:f].copy(), idx=idx[:f].copy(),
                curves=curves[:f].copy(), power=power[:f].copy())


def multiband_run(fs, mic_dist, x, nbins=15, chunk=0, use_floor=False, noise_preestimated=True, prefix="orc"):
    """MultibandBinarualLocalisation (MultibandBinarualLocalisation.cpp:52-259) over a whole stereo signal."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[1]
    fn = getattr(lib(prefix), f"{prefix}_multiband_run")
    D = C.c_int(0)
    N = fn(C.c_int(fs), C.c_double(mic_dist), C.c_int(nbins), 0, 1, None, 0, 0, 0, None, None, None, C.byref(D), None, None, None, None, None, None)
    D = D.value
    maxf = max(1, n // (N // 2) + 2)
    n_frames, n_fired = C.c_int(0), C.c_int(0)
    fired_frame = np.zeros(maxf, dtype=np.int32); cell = np.zeros(maxf, dtype=np.int32); band_cells = np.zeros((maxf, nbins), dtype=np.int32)
    prob = np.zeros(maxf); power = np.zeros(maxf); doa_deg = np.zeros(maxf); hist = np.zeros((maxf, D))
    Dd = C.c_int(0)
    r = fn(C.c_int(fs), C.c_double(mic_dist), C.c_int(nbins), C.c_int(int(use_floor)), C.c_int(int(noise_preestimated)), _dp(x), C.c_int(n),
           C.c_int(chunk), C.c_int(maxf), C.byref(n_frames), C.byref(n_fired), _ip(fired_frame), C.byref(Dd), _ip(cell), _dp(prob), _dp(power),
           _dp(doa_deg), _dp(hist), _ip(band_cells))
    f = n_fired.value
    return dict(N=r, D=D, n_frames=n_frames.value, n_fired=f, fired_frame=fired_frame[:f].copy(), cell=cell[:f].copy(), prob=prob[:f].copy(),
                power=power[:f].copy(), doa_deg=doa_deg[:f].copy(), hist=hist[:f].copy(), band_cells=band_cells[:f].copy())


def multiband_frames(spec, N, fs, mic_dist, nbins=15):
    """frame-level restatement on given spectra [T][2][K] (fresh state, gate off): cells, prob, histogram, per-band cells, H"""
    s = _ccs(spec); T = s.shape[0]
    H = np.zeros((nbins, N // 2 + 1))
    D = lib().orc_multiband_frames(None, 0, C.c_int(N), C.c_int(fs), C.c_double(mic_dist), C.c_int(nbins), None, None, None, None, _dp(H))
    cell = np.zeros(T, dtype=np.int32); prob = np.zeros(T); hist = np.zeros((T, D)); bc = np.zeros((T, nbins), dtype=np.int32)
    lib().orc_multiband_frames(_dp(s), C.c_int(T), C.c_int(N), C.c_int(fs), C.c_double(mic_dist), C.c_int(nbins), _ip(cell), _dp(prob), _dp(hist),
                               _ip(bc), None)
    return dict(cell=cell, prob=prob, hist=hist, band_cells=bc, H=H, D=D)


def freqgcc_probability(fs, mic_dist, curve, doas, prefix="orc"):
    curve = np.ascontiguousarray(curve, dtype=np.float64); doas = np.ascontiguousarray(doas, dtype=np.float64)
    probs = np.zeros(len(doas))
    getattr(lib(prefix), f"{prefix}_freqgcc_probability")(C.c_int(fs), C.c_double(mic_dist), _dp(curve), _dp(doas), _dp(probs), C.c_int(len(doas)))
    return probs


def mask_run(fs, mic_dist, lo, hi, method, alg, x, chunk=0, want_spectra=False, prefix="orc"):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[1]
    N = frame_size(fs, np.float32(0.050), prefix)
    maxf = max(1, n // (N // 2) + 2)
    cap = n + 2 * N
    out = np.zeros((2, cap)); n_out, n_frames = C.c_int(0), C.c_int(0)
    Q = np.zeros((maxf, 45))
    spec = np.zeros((maxf, 2, N + 2)) if want_spectra else None
    r = getattr(lib(prefix), f"{prefix}_mask_run")(
        C.c_int(fs), C.c_double(mic_dist), C.c_float(lo), C.c_float(hi), C.c_int(method), C.c_int(alg),
        _dp(x), C.c_int(n), C.c_int(chunk), _dp(out), C.c_int(cap), C.byref(n_out), C.c_int(maxf), C.byref(n_frames), _dp(Q), _dp(spec))
    assert r > 0, r
    T = n_frames.value
    return dict(N=r, out=out[:, :n_out.value].copy(), n_frames=T, Q=Q[:T].copy(), spectra=None if spec is None else spec[:T].copy())


def beamformer_frame(fs, mic_xyz, frames, doa, prefix="orc"):
    frames = np.ascontiguousarray(frames, dtype=np.float64); xyz = np.ascontiguousarray(mic_xyz, dtype=np.float64)
    M, ccs = frames.shape
    out = np.zeros(ccs)
    getattr(lib(prefix), f"{prefix}_beamformer_frame")(C.c_int(fs), C.c_int(M), _dp(xyz), C.c_int(ccs), _dp(frames), C.c_double(doa), _dp(out))
    return out


def steering_frames(fs, mic_xyz, frames, S, prefix="orc"):
    frames = np.ascontiguousarray(frames, dtype=np.float64); xyz = np.ascontiguousarray(mic_xyz, dtype=np.float64)
    T, M, ccs = frames.shape
    doa = np.zeros((T, S)); prob = np.zeros((T, S)); energy = np.zeros((T, 37))
    getattr(lib(prefix), f"{prefix}_steering_frames")(C.c_int(fs), C.c_int(M), _dp(xyz), C.c_int(ccs), C.c_int(S), _dp(frames), C.c_int(T),
                                                          _dp(doa), _dp(prob), _dp(energy))
    return dict(doa_rad=doa, prob=prob, energy=energy)


# ---------------------------------------------------------------------------------------------
# orc-only building blocks (explicit conventions, generalised geometry)
# ---------------------------------------------------------------------------------------------
def sqrt_hann(N):
    w = np.zeros(N)
    lib().orc_sqrt_hann(C.c_int(N), _dp(w))
    return w


def stft(x, N, hop, win=None):
    """x [M][n] -> complex spectra [T][M][K]."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    M, n = x.shape
    T = lib().orc_stft(_dp(x), C.c_int(M), C.c_int(n), C.c_int(N), C.c_int(hop), None, None, C.c_int(0))
    spec = np.zeros((T, M, N + 2))
    w = None if win is None else np.ascontiguousarray(win, dtype=np.float64)
    lib().orc_stft(_dp(x), C.c_int(M), C.c_int(n), C.c_int(N), C.c_int(hop), _dp(w), _dp(spec), C.c_int(T))
    return spec.view(np.complex128)


def _ccs(spec):
    return np.ascontiguousarray(spec, dtype=np.complex128).view(np.float64)


def istft(spec, N, hop, win=None, tail=None):
    s = _ccs(spec)
    T, Cc = s.shape[0], s.shape[1]
    out = np.zeros((Cc, T * hop))
    w = None if win is None else np.ascontiguousarray(win, dtype=np.float64)
    lib().orc_istft(_dp(s), C.c_int(T), C.c_int(Cc), C.c_int(N), C.c_int(hop), _dp(w), _dp(out), _dp(tail))
    return out


def fft_log_power(spec, N):
    s = _ccs(spec); T, M = s.shape[0], s.shape[1]
    p = np.zeros(T)
    lib().orc_fft_log_power(_dp(s), C.c_int(T), C.c_int(M), C.c_int(N), _dp(p))
    return p


def gcc_tau_frames(spec, N, pair_tau):
    s = _ccs(spec); T, M = s.shape[0], s.shape[1]
    tau = np.ascontiguousarray(pair_tau, dtype=np.float64); P, D = tau.shape
    assert P == M * (M - 1) // 2
    corr = np.zeros((T, P, D))
    lib().orc_gcc_tau_frames(_dp(s), C.c_int(T), C.c_int(M), C.c_int(N), _dp(tau), C.c_int(D), _dp(corr))
    return corr


def tdoa_lags(spec, N, max_lag, want_curves=True):
    s = _ccs(spec); T, M = s.shape[0], s.shape[1]
    P = M * (M - 1) // 2
    curves = np.zeros((T, P, 2 * max_lag + 1)) if want_curves else None
    lags = np.zeros((T, P), dtype=np.int32)
    lib().orc_tdoa_lags(_dp(s), C.c_int(T), C.c_int(M), C.c_int(N), C.c_int(max_lag), _dp(curves), _ip(lags))
    return curves, lags


def energy_scan(corr, a=float(np.float32(0.8)), b=float(1 - np.float32(0.8)), active=None, state=None):
    corr = np.ascontiguousarray(corr, dtype=np.float64); T, P, D = corr.shape
    st = np.zeros(D) if state is None else state
    e = np.zeros((T, D))
    act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
    lib().orc_energy_scan(_dp(corr), C.c_int(T), C.c_int(P), C.c_int(D), C.c_double(a), C.c_double(b),
                          None if act is None else act.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(st), _dp(e))
    return e, st


def select_doa(energy, n_pairs, S):
    energy = np.ascontiguousarray(energy, dtype=np.float64); T, D = energy.shape
    idx = np.zeros((T, S), dtype=np.int32); prob = np.zeros((T, S))
    lib().orc_select_doa(_dp(energy), C.c_int(T), C.c_int(D), C.c_int(n_pairs), C.c_int(S), _ip(idx), _dp(prob))
    return idx, prob


def ds_fan(spec, N, fs, mic_x, doas):
    s = _ccs(spec); T, M = s.shape[0], s.shape[1]
    mic_x = np.ascontiguousarray(mic_x, dtype=np.float64); doas = np.ascontiguousarray(doas, dtype=np.float64)
    out = np.zeros((T, len(doas), N + 2))
    lib().orc_ds_fan(_dp(s), C.c_int(T), C.c_int(M), C.c_int(N), C.c_int(fs), _dp(mic_x), _dp(doas), C.c_int(len(doas)), _dp(out))
    return out.view(np.complex128)


def fs_fan(spec, N, W):
    """filter-and-sum fan: spec [T][M][K] complex, W [D][M][K] complex -> [T][D][K] complex"""
    s = _ccs(spec); T, M = s.shape[0], s.shape[1]
    W = np.ascontiguousarray(W, dtype=np.complex128); D = W.shape[0]
    out = np.zeros((T, D, N + 2))
    lib().orc_fs_fan(_dp(s), C.c_int(T), C.c_int(M), C.c_int(N), _dp(W.view(np.float64)), C.c_int(D), _dp(out))
    return out.view(np.complex128)


def mic_tau(xyz, fs, dirs):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); dirs = np.ascontiguousarray(dirs, dtype=np.float64)
    out = np.zeros((len(xyz), len(dirs)))
    lib().orc_mic_tau(_dp(xyz), C.c_int(len(xyz)), C.c_int(fs), _dp(dirs), C.c_int(len(dirs)), _dp(out))
    return out


def pair_tau_from_mic_tau(mt):
    mt = np.ascontiguousarray(mt, dtype=np.float64); M, D = mt.shape
    out = np.zeros((M * (M - 1) // 2, D))
    lib().orc_pair_tau_from_mic_tau(_dp(mt), C.c_int(M), C.c_int(D), _dp(out))
    return out


def reference_pair_tau(xyz, fs, step):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); M = len(xyz)
    D = C.c_int(0)
    lib().orc_reference_pair_tau(_dp(xyz), C.c_int(M), C.c_int(fs), C.c_float(step), None, C.byref(D))
    out = np.zeros((M * (M - 1) // 2, D.value))
    lib().orc_reference_pair_tau(_dp(xyz), C.c_int(M), C.c_int(fs), C.c_float(step), _dp(out), C.byref(D))
    return out


def srp_channel(spec, N, mic_tau_, n_threads=8):
    s = _ccs(spec); T, M = s.shape[0], s.shape[1]
    mt = np.ascontiguousarray(mic_tau_, dtype=np.float64); D = mt.shape[1]
    out = np.zeros((T, D))
    lib().orc_srp_channel(_dp(s), C.c_int(T), C.c_int(M), C.c_int(N), _dp(mt), C.c_int(D), _dp(out), C.c_int(n_threads))
    return out


def mel_bank(N, n_bands, fs, lo, hi):
    H = np.zeros((n_bands, N // 2 + 1)); fc = np.zeros(n_bands)
    lib().orc_mel_bank(C.c_int(N), C.c_int(n_bands), C.c_int(fs), C.c_float(lo), C.c_float(hi), _dp(H), _dp(fc))
    return H, fc


def mask_frames(spec, N, fs, mic_dist, method, alg, H, fc, state=None):
    """spec [T][2][K] complex -> masked copy, decisions [T][nb], Q trace [T][nb]; state = dict(Q, noise, first_call)."""
    s = _ccs(spec).copy(); T = s.shape[0]
    H = np.ascontiguousarray(H, dtype=np.float64); fc = np.ascontiguousarray(fc, dtype=np.float64); nb = len(fc)
    if state is None:
        state = dict(Q=np.zeros(nb), noise=np.zeros(nb), first_call=0)
    fcall = C.c_int(state["first_call"])
    dec = np.zeros((T, nb), dtype=np.int32); Qt = np.zeros((T, nb))
    lib().orc_mask_frames(_dp(s), C.c_int(T), C.c_int(N), C.c_int(fs), C.c_double(mic_dist), C.c_int(method), C.c_int(alg), C.c_int(nb),
                          _dp(H), _dp(fc), _dp(state["Q"]), _dp(state["noise"]), C.byref(fcall), _ip(dec), _dp(Qt))
    state["first_call"] = fcall.value
    return s.view(np.complex128), dec, Qt, state


def freqgcc_frames(spec, N, fs, mic_dist):
    s = _ccs(spec); T = s.shape[0]
    D = lib().orc_freqgcc_grid(C.c_int(fs), C.c_double(mic_dist), None)
    curves = np.zeros((T, D)); idx = np.zeros(T, dtype=np.int32)
    lib().orc_freqgcc_frames(_dp(s), C.c_int(T), C.c_int(N), C.c_int(fs), C.c_double(mic_dist), _dp(curves), _ip(idx))
    return curves, idx


def freqgcc_grid(fs, mic_dist):
    D = lib().orc_freqgcc_grid(C.c_int(fs), C.c_double(mic_dist), None)
    tau = np.zeros(D)
    lib().orc_freqgcc_grid(C.c_int(fs), C.c_double(mic_dist), _dp(tau))
    return tau


def tdoa_pipeline(x, N, hop, max_lag, n_threads=1):
    """x [B][M][n] float64 -> lags [B][T][P] (CPU baseline leg)."""
    x = np.ascontiguousarray(x, dtype=np.float64); B, M, n = x.shape
    T = (n - N) // hop + 1; P = M * (M - 1) // 2
    lags = np.zeros((B, T, P), dtype=np.int32)
    lib().orc_tdoa_pipeline(_dp(x), C.c_int(B), C.c_int(M), C.c_int(n), C.c_int(N), C.c_int(hop), C.c_int(max_lag), _ip(lags), C.c_int(n_threads))
    return lags

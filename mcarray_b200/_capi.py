"""ctypes binding of libmcarray_b200.so (include/mcarray_b200.h).  No fallback: importing without the built
library, or creating a processor without a CUDA device, raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCAG_LIB_PATH") or os.path.join(_HERE, "libmcarray_b200.so")   # override: kernel-variant experiments (tools/variants.sh)

KIND_SSL, KIND_SL, KIND_FREQGCC, KIND_MASK, KIND_TDOA, KIND_DSFAN, KIND_SRP, KIND_MULTIBAND = range(8)
(OUT_SPECTRA, OUT_POWER_DB, OUT_CORR, OUT_ENERGY, OUT_CELL, OUT_PROB, OUT_LAGS, OUT_CURVES, OUT_ACTIVE, OUT_BEAMS,
 OUT_MASK_Q, OUT_MASK_DEC, OUT_BAND_CELL, OUT_TRACK_DOA) = range(14)
EMIT_CORR, EMIT_CURVES, EMIT_SPECTRA, EMIT_MASK_TRACE = 1, 2, 4, 8

c_dp = C.POINTER(C.c_double)


class Config(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("device", C.c_int), ("sample_rate", C.c_int), ("frame_size", C.c_int), ("hop", C.c_int),
        ("n_channels", C.c_int), ("n_streams", C.c_int), ("max_frames_per_call", C.c_int), ("window", c_dp), ("emit", C.c_int),
        ("n_dirs", C.c_int), ("pair_tau", c_dp), ("mic_tau", c_dp), ("steer_turns", c_dp), ("n_sources", C.c_int),
        ("energy_memory", C.c_float), ("corr_memory", C.c_float), ("use_power_floor", C.c_int), ("noise_margin_db", C.c_float),
        ("floor_seconds", C.c_float), ("floor_ccs_power", C.c_int), ("noise_preestimated", C.c_int), ("max_lag", C.c_int),
        ("mask_method", C.c_int), ("mask_alg", C.c_int), ("n_bands", C.c_int), ("band_coefs", c_dp), ("band_thresholds", c_dp),
        ("srp_form", C.c_int), ("doa_tracker", C.c_int), ("doa_memory", C.c_float), ("fs_weights", c_dp),
    ]


class Info(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "frame_size", "window_size", "hop", "analysis_length", "one_sided_length", "n_channels", "n_streams", "max_latency",
        "n_dirs", "n_pairs", "n_sources", "n_out_channels", "spectrum_pitch", "max_frames_per_call", "srp_form", "beams_pitch")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C mcarray_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.mcag_last_error.restype = C.c_char_p
        L.mcag_device_ptr.restype = C.c_void_p
        L.mcag_stream.restype = C.c_void_p
        L.mcag_host_alloc.restype = C.c_void_p
        L.mcag_host_alloc.argtypes = [C.c_longlong]
        L.mcag_host_free.argtypes = [C.c_void_p]
        L.mcag_frames_total.restype = C.c_longlong
        L.mcag_kernel_launches.restype = C.c_longlong
        L.mcag_geom_cell_angle.restype = C.c_double
        L.mcag_geom_frame_size.argtypes = [C.c_int, C.c_double]
        L.mcag_geom_grid_size.argtypes = [C.c_float]
        L.mcag_geom_cell_angle.argtypes = [C.c_int, C.c_float]
        _lib = L
    return _lib


class McagError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise McagError(f"mcarray_b200 error {rc}: {lib().mcag_last_error().decode()}")


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def vp(x):
    """device pointer of a torch tensor / raw int / None"""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


# ---- host geometry ---------------------------------------------------------------------------------------------------
def frame_size(fs, frame_rate):
    return lib().mcag_geom_frame_size(int(fs), float(np.float32(frame_rate)))


def grid_size(step):
    return lib().mcag_geom_grid_size(C.c_float(step))


def cell_angle(idx, step):
    return lib().mcag_geom_cell_angle(int(idx), C.c_float(step))


def pair_tau_reference(xyz, fs, step):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); M = len(xyz); D = grid_size(step)
    tau = np.zeros((M * (M - 1) // 2, D))
    lib().mcag_geom_pair_tau_reference(dp(xyz), M, int(fs), C.c_float(step), dp(tau))
    return tau


def steer_turns_reference(xyz, fs, N, step):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); M = len(xyz); D = grid_size(step)
    t = np.zeros((D + 1, M))
    lib().mcag_geom_steer_turns_reference(dp(xyz), M, int(fs), int(N), C.c_float(step), dp(t))
    return t


def steer_turns(xyz, fs, N, doas):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); doas = np.ascontiguousarray(doas, dtype=np.float64)
    t = np.zeros((len(doas), len(xyz)))
    lib().mcag_geom_steer_turns(dp(xyz), len(xyz), int(fs), int(N), dp(doas), len(doas), dp(t))
    return t


def mic_tau(xyz, fs, dirs):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); dirs = np.ascontiguousarray(dirs, dtype=np.float64)
    t = np.zeros((len(xyz), len(dirs)))
    lib().mcag_geom_mic_tau(dp(xyz), len(xyz), int(fs), dp(dirs), len(dirs), dp(t))
    return t


def pair_tau_from_mic_tau(mt):
    mt = np.ascontiguousarray(mt, dtype=np.float64); M, D = mt.shape
    t = np.zeros((M * (M - 1) // 2, D))
    lib().mcag_geom_pair_tau_from_mic_tau(dp(mt), M, D, dp(t))
    return t


def multiband_setup(fs, mic_dist, N, nb):
    D = lib().mcag_geom_multiband(int(fs), C.c_double(mic_dist), int(N), int(nb), None, None)
    tau = np.zeros(D); H = np.zeros((nb, N // 2 + 1))
    lib().mcag_geom_multiband(int(fs), C.c_double(mic_dist), int(N), int(nb), dp(tau), dp(H))
    return tau, H


def mel_bank(N, nb, fs, lo, hi, mic_dist):
    H = np.zeros((nb, N // 2 + 1)); fc = np.zeros(nb); thr = np.zeros(nb)
    lib().mcag_geom_mel_bank(int(N), int(nb), int(fs), C.c_float(lo), C.c_float(hi), C.c_double(mic_dist), dp(H), dp(fc), dp(thr))
    return H, fc, thr

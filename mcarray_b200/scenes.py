"""Deterministic synthetic far-field scenes: band-limited Gaussian sources, exact
fractional delays applied in the frequency domain over the whole signal, independent sensor noise.
Pure numpy, float64; shared by tests/ and bench.py so GPU, oracle and CPU baseline see identical input.
"""
import numpy as np

SPEED_OF_SOUND = 346.1  # reference: src/mcarray/microhponeArrayHelpers.cpp:38-43


def azimuth_dirs(theta):
    """Unit vectors for azimuths measured from broadside (+y) towards +x: u = (sin th, cos th, 0)."""
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    return np.stack([np.sin(theta), np.cos(theta), np.zeros_like(theta)], axis=1)


def az_el_dirs(az, el):
    az, el = np.broadcast_arrays(np.asarray(az, dtype=np.float64), np.asarray(el, dtype=np.float64))
    return np.stack([np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el)], axis=-1).reshape(-1, 3)


def linear_array(x):
    x = np.asarray(x, dtype=np.float64)
    return np.stack([x, np.zeros_like(x), np.zeros_like(x)], axis=1)


def circular_array(M, radius):
    a = 2 * np.pi * np.arange(M) / M
    return np.stack([radius * np.cos(a), radius * np.sin(a), np.zeros(M)], axis=1)


def planar_array(nx, ny, pitch):
    gx, gy = np.meshgrid((np.arange(nx) - (nx - 1) / 2) * pitch, (np.arange(ny) - (ny - 1) / 2) * pitch, indexing="ij")
    return np.stack([gx.ravel(), gy.ravel(), np.zeros(nx * ny)], axis=1)


def far_field_scene(mic_xyz, fs, n, src_dirs, seed, snr_db=20.0, amp=5000.0, band=None):
    """Returns x [M][n] float64.  A source in direction u reaches mic m with the time ADVANCE
    (p_m . u)/c, the convention Beamformer.cpp:59 implies (see oracle/CONVENTIONS.md C5)."""
    mic_xyz = np.asarray(mic_xyz, dtype=np.float64)
    src_dirs = np.asarray(src_dirs, dtype=np.float64).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    M = mic_xyz.shape[0]
    f = np.fft.rfftfreq(n, 1.0 / fs)
    lo, hi = (100.0, 0.45 * fs) if band is None else band
    bp = ((f >= lo) & (f <= hi)).astype(np.float64)
    x = np.zeros((M, n))
    for u in src_dirs:
        s = np.fft.rfft(rng.standard_normal(n)) * bp
        adv = mic_xyz @ u / SPEED_OF_SOUND  # seconds
        x += np.fft.irfft(s[None, :] * np.exp(2j * np.pi * f[None, :] * adv[:, None]), n)
    sig_rms = np.sqrt(np.mean(x ** 2))
    x += rng.standard_normal((M, n)) * sig_rms * 10 ** (-snr_db / 20.0)
    return x * (amp / np.max(np.abs(x)))


def stream_seed(stream_id):
    return 1234 + 7919 * int(stream_id)

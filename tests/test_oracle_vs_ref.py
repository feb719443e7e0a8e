"""Pins the oracle restatement (oracle/restated.hpp) to the REFERENCE'S OWN CODE: oracle/_ref is the
reference's .cpp files compiled from /root/reference against the DSPONE/WIPP stand-in (oracle/Makefile).
Everything mcarray owns must agree bit for bit.  Skipped only when no prebuilt oracle/_ref exists."""
import os

import numpy as np
import pytest

from mcarray_b200 import scenes


@pytest.fixture(autouse=True)
def _need_ref(ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref/libmcarray_ref.so not built (needs /root/reference)")


def _eq(a, b, keys):
    for k in keys:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("positions,fs,deg,S", [
    ([0, 0.07, 0.175, 0.21], 16000, 30, 2),          # Reem-C array, test_mcarray.cpp:397
    ([-2.25, -1.25, 1.25, 2.25], 16000, -40, 1),      # mcbeam's array, mcabeamf.cpp:182
    ([0, 0.07, 0.175, 0.21], 48000, 10, 3),
])
def test_ssl_bit_exact(orc, positions, fs, deg, S):
    xyz = scenes.linear_array(positions)
    x = scenes.far_field_scene(xyz, fs, fs, scenes.azimuth_dirs([np.deg2rad(deg)]), seed=11)
    a = orc.ssl_run(fs, xyz, S, x, chunk=1024, want_corr=True)
    b = orc.ssl_run(fs, xyz, S, x, chunk=1024, want_corr=True, prefix="ref")
    assert a["N"] == b["N"] and a["n_frames"] == b["n_frames"] and a["n_fired"] == b["n_fired"] > 0
    _eq(a, b, ["out", "doa_deg", "prob", "power", "energy", "corr_scaled", "fired_frame"])


def test_ssl_power_floor_and_analysis_only(orc):
    fs = 16000
    xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    x = scenes.far_field_scene(xyz, fs, 5 * fs, scenes.azimuth_dirs([np.deg2rad(20)]), seed=5)
    x[:, : 3 * fs + 4000] *= 1e-3  # quiet lead-in: floor estimated on it, then the source starts
    for analysis_only in (False, True):
        a = orc.ssl_run(fs, xyz, 1, x, chunk=4096, use_floor=True, analysis_only=analysis_only)
        b = orc.ssl_run(fs, xyz, 1, x, chunk=4096, use_floor=True, analysis_only=analysis_only, prefix="ref")
        assert 0 < a["n_fired"] < a["n_frames"]
        assert a["n_fired"] == b["n_fired"]
        _eq(a, b, ["out", "doa_deg", "prob", "power", "energy", "fired_frame"])


def test_freqgcc_bit_exact(orc):
    fs = 16000
    xyz = scenes.linear_array([0, 0.086])  # test_mcarray.cpp:284
    x = scenes.far_field_scene(xyz, fs, 3 * fs, scenes.azimuth_dirs([np.deg2rad(-21)]), seed=2)
    for floor in (False, True):
        a = orc.freqgcc_run(fs, 0.086, x, chunk=777, use_floor=floor)
        b = orc.freqgcc_run(fs, 0.086, x, chunk=777, use_floor=floor, prefix="ref")
        assert a["n_fired"] == b["n_fired"] > 0
        _eq(a, b, ["fired_frame", "curves", "idx", "power"])
    doas = np.linspace(-1.5, 1.5, 41)
    c = a["curves"][5]
    assert np.array_equal(orc.freqgcc_probability(fs, 0.086, c, doas), orc.freqgcc_probability(fs, 0.086, c, doas, prefix="ref"))


def test_freqgcc_power_floor_pauses_bit_exact(orc):
    """the silence branch of BinauralLocalisation.cpp:528-561 against the reference build, on the scene the golden fixture is made from"""
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(sys_path, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    x = mg.pause_scene()
    a = orc.freqgcc_run(16000, 0.086, x, chunk=999, use_floor=True)
    b = orc.freqgcc_run(16000, 0.086, x, chunk=999, use_floor=True, prefix="ref")
    voiced = b["power"] > 0          # the particle filter also delivers on silent frames inside the decay window (:540-544)
    assert a["n_fired"] == voiced.sum() == 30 and b["n_fired"] > a["n_fired"]
    for k in ("fired_frame", "curves", "idx", "power"):
        assert np.array_equal(a[k], b[k][voiced]), k


@pytest.mark.parametrize("method", [0, 1, 3, 4, 5])
@pytest.mark.parametrize("alg", [0, 1, 2])
def test_mask_bit_exact(orc, method, alg):
    n = 5 * 1024  # the reference's own spatial-masking test signal, test_mcarray.cpp:908-929
    i = np.arange(n)
    sig = np.round(5000 * np.cos(2 * np.pi * 0.1 * i)); inter = np.round(5000 * np.cos(2 * np.pi * 0.3 * i))
    L = sig + inter; R = sig.copy(); R[: n - 6] += inter[6:]
    x = np.stack([L, R])
    a = orc.mask_run(16000, 0.086, 500, 5000, method, alg, x, chunk=1000, want_spectra=True)
    b = orc.mask_run(16000, 0.086, 500, 5000, method, alg, x, chunk=1000, want_spectra=True, prefix="ref")
    assert a["n_frames"] == b["n_frames"] == 9
    _eq(a, b, ["out", "Q", "spectra"])


def test_helpers_and_frames_bit_exact(orc):
    for step_deg in (3, 5):
        st = np.float32(step_deg * np.pi / 180)
        for d in range(int(round(np.pi / st)) + 1):
            ang = orc.doa_idx_to_angle(d, st)
            assert ang == orc.doa_idx_to_angle(d, st, "ref")
            for dist in (0.035, 0.07, 0.086, 0.105, 0.21, 2.5, 4.5):
                for fs in (16000, 44100, 48000):
                    assert orc.doa_to_delay_samples(ang, dist, fs) == orc.doa_to_delay_samples(ang, dist, fs, "ref")
            for eps in (-0.02, 0.0, 0.011):
                assert orc.angle_to_doa_idx(ang + eps, st) == orc.angle_to_doa_idx(ang + eps, st, "ref")
    for fs, fr in ((16000, 0.025), (48000, 0.025), (16000, 0.075), (16000, 0.05), (44100, 0.025), (8000, 0.025)):
        assert orc.frame_size(fs, np.float32(fr)) == orc.frame_size(fs, np.float32(fr), "ref")
    rng = np.random.default_rng(0)
    xyz = scenes.linear_array([-2.25, -1.25, 1.25, 2.25])
    fr = rng.standard_normal((6, 4, 2050))
    for doa in (-1.2, 0.0, 0.3):
        assert np.array_equal(orc.beamformer_frame(48000, xyz, fr[0], doa), orc.beamformer_frame(48000, xyz, fr[0], doa, "ref"))
    a = orc.steering_frames(48000, xyz, fr, 3); b = orc.steering_frames(48000, xyz, fr, 3, "ref")
    _eq(a, b, ["doa_rad", "prob", "energy"])


def test_multiband_bit_exact(orc):
    """N2: MultibandBinarualLocalisation restatement against the reference's own MultibandBinarualLocalisation.cpp."""
    fs = 16000
    xyz = scenes.linear_array([0, 0.089])  # test_mcarray.cpp:325
    x = scenes.far_field_scene(xyz, fs, 2 * fs, scenes.azimuth_dirs([np.deg2rad(40)]), seed=3)
    x[:, fs:] *= 1e-4                       # near silence (the reference compares a LINEAR power with its dB floor, :221-225)
    for nbins, floor in ((15, False), (8, True)):
        a = orc.multiband_run(fs, 0.089, x, nbins=nbins, chunk=900, use_floor=floor)
        b = orc.multiband_run(fs, 0.089, x, nbins=nbins, chunk=900, use_floor=floor, prefix="ref")
        assert a["N"] == b["N"] == 512 and a["D"] == b["D"] == 37 and a["n_frames"] == b["n_frames"]
        assert a["n_fired"] == b["n_fired"] > 0
        _eq(a, b, ["fired_frame", "cell", "prob", "power", "doa_deg", "hist", "band_cells"])

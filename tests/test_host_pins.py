"""CPU pins of the host-side constructor math (mcarray_b200/csrc/hostgeom.cpp, exported through the C ABI as mcag_geom_*) and of the
generalised oracles, against the restatement and — where oracle/_ref exists — the reference's own code.  No GPU work: the geometry
entry points are plain host functions of libmcarray_b200.so."""
import numpy as np
import pytest

from mcarray_b200 import scenes

ARRAYS = {
    "mcbeam": [-2.25, -1.25, 1.25, 2.25],                  # mcabeamf.cpp:182
    "reemc": [0, 0.07, 0.175, 0.21],                       # test_mcarray.cpp:397
    "binaural": [0, 0.086],                                # test_mcarray.cpp:284
    "lin16": list((np.arange(16) - 7.5) * 0.035),          # BASELINE config 5
}


@pytest.fixture(scope="module")
def capi():
    from mcarray_b200 import capi as c
    c.lib()
    return c


@pytest.mark.parametrize("name", sorted(ARRAYS))
@pytest.mark.parametrize("fs", [16000, 44100, 48000])
@pytest.mark.parametrize("step_deg", [5, 3])
def test_pair_tau_reference_is_bit_identical(capi, orc, ref_available, name, fs, step_deg):
    """SteeringBeamforming::generateLookupTable (SteeringBeamforming.cpp:58-94) / BinauralLocalisation.cpp:363-366: float-typed helper chain"""
    xyz = scenes.linear_array(ARRAYS[name])
    step = np.float32(step_deg * np.pi / 180)
    got = capi.pair_tau_reference(xyz, fs, step)
    want = orc.reference_pair_tau(xyz, fs, step)
    assert got.shape == want.shape and np.array_equal(got, want)
    if ref_available:   # every entry through the reference's own doaToDelayFarFieldSamples(doaIdx2angle(d), |p_i - p_j|, fs)
        M, p = len(xyz), 0
        for i in range(M):
            for j in range(i + 1, M):
                dist = orc.array_distance(xyz, i, j, prefix="ref")
                for d in range(got.shape[1]):
                    ang = orc.doa_idx_to_angle(d, step, prefix="ref")
                    assert got[p, d] == orc.doa_to_delay_samples(ang, dist, fs, prefix="ref"), (p, d)
                p += 1


def test_cell_angle_and_frame_size_are_bit_identical(capi, orc, ref_available):
    prefixes = ["orc"] + (["ref"] if ref_available else [])
    for prefix in prefixes:
        for step_deg in (3, 5):
            step = np.float32(step_deg * np.pi / 180)
            for idx in range(0, capi.grid_size(step) + 2):
                assert capi.cell_angle(idx, step) == orc.doa_idx_to_angle(idx, step, prefix=prefix)
        for fs in (8000, 16000, 22050, 44100, 48000):
            for rate in (0.025, 0.050, 0.075):
                assert capi.frame_size(fs, rate) == orc.frame_size(fs, np.float32(rate), prefix=prefix), (fs, rate)
    assert capi.grid_size(np.float32(5 * np.pi / 180)) == 37 and capi.grid_size(np.float32(3 * np.pi / 180)) == 61   # SteeringBeamforming.cpp:40, BinauralLocalisation.cpp:329


@pytest.mark.parametrize("N,fs,lo,hi", [(512, 16000, 500, 5000), (1024, 16000, 500, 5000), (2048, 48000, 400, 4000), (256, 8000, 300, 3400)])
def test_mel_bank_is_bit_identical(capi, orc, N, fs, lo, hi):
    """FastBinauralMasking constructor tables (FastBinauralMasking.cpp:95-98,342-366): 45 band magnitudes, centre frequencies, and the
    spatial thresholds cos(2 pi f_b d sin(10 deg) / c)"""
    H, fc, thr = capi.mel_bank(N, 45, fs, lo, hi, 0.086)
    Ho, fco = orc.mel_bank(N, 45, fs, lo, hi)
    assert np.array_equal(H, Ho) and np.array_equal(fc, fco)
    # thresholds: decisions of the oracle's own mask state on a frame whose normalised correlation sits exactly at the threshold are
    # covered on the GPU (test_mask_decisions_vs_oracle); here the closed form the constructor evaluates
    want = np.cos(2 * np.pi * (fc * fs) * 0.086 * np.sin(10 * np.pi / 180) / 346.1)
    assert np.allclose(thr, want, rtol=0, atol=1e-12)


def test_multiband_setup_matches_the_helper_chain(capi, orc, ref_available):
    """MultibandBinarualLocalisation.cpp:62-63,99: D = floor(pi/step)+1 delays on the 5 degree grid"""
    step = np.float32(5 * np.pi / 180)
    for prefix in ["orc"] + (["ref"] if ref_available else []):
        for fs in (16000, 48000):
            tau, H = capi.multiband_setup(fs, 0.086, 512, 15)
            assert len(tau) == int(np.floor(np.pi / step) + 1) and H.shape == (15, 257)
            for d in range(len(tau)):
                assert tau[d] == orc.doa_to_delay_samples(orc.doa_idx_to_angle(d, step, prefix=prefix), np.float32(0.086), fs, prefix=prefix)


@pytest.mark.parametrize("name,fs,N", [("lin16", 16000, 512), ("reemc", 48000, 2048)])
def test_ds_fan_oracle_reduces_to_the_reference_beamformer(orc, ref_available, name, fs, N):
    """The generalised fan oracle (BASELINE config 3: every azimuth of a fan) must be Beamformer::processFrame (Beamformer.cpp:51-71) called
    once per azimuth — bit for bit against the reference build where it exists, and against the restated frame function always."""
    xyz = scenes.linear_array(ARRAYS[name])
    x = scenes.far_field_scene(xyz, fs, N + 2 * (N // 2), scenes.azimuth_dirs([0.4]), seed=17)
    S = orc.stft(x, N, N // 2)                                        # [T][M][K]
    doas = np.deg2rad(np.arange(-90, 91, 7.5))
    fan = orc.ds_fan(S, N, fs, xyz[:, 0], doas)                       # [T][D][K]
    ccs = np.ascontiguousarray(S).view(np.float64).reshape(S.shape[0], S.shape[1], N + 2)
    for prefix in ["orc"] + (["ref"] if ref_available else []):
        for t in range(S.shape[0]):
            for d, doa in enumerate(doas):
                one = orc.beamformer_frame(fs, xyz, ccs[t], doa, prefix=prefix)
                assert np.array_equal(np.ascontiguousarray(fan[t, d]).view(np.float64), one), (prefix, t, d)


def test_kernel_schedule_prototypes():
    """the numpy restatements of the two register-resident transform schedules the CUDA kernels implement (tools/proto): the half-warp
    real FFT of fft16.cuh for both sizes, and the decimated inverse transform of the warp-synchronous lag phase (tdoa_warp.cuh)"""
    import glob
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scripts = sorted(glob.glob(os.path.join(root, "tools", "proto", "*.py")))
    assert scripts
    for s in scripts:
        r = subprocess.run([sys.executable, s], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (s, r.stdout[-500:], r.stderr[-500:])

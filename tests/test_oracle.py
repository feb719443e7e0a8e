"""CPU tests of the oracle itself: golden fixtures generated from the reference build, the reference's
own known-answer values (testArrayDescription), analytic known answers and the numpy twin."""
import os

import numpy as np
import pytest

from mcarray_b200 import scenes

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_array_description_known_answers(orc):
    # /root/reference/test/test_mcarray.cpp:518-580, exact doubles
    xyz = np.array([[0, 2, 3], [0.035 * 2, 2, 3], [0.035 * 5, 2, 3], [0.035 * 6, 2, 3]], dtype=np.float64)
    want = {(0, 0): 0.0, (0, 1): 0.070, (0, 2): 0.175, (0, 3): 0.210, (1, 0): 0.070, (1, 1): 0.0, (1, 2): 0.105, (1, 3): 0.140,
            (2, 0): 0.175, (2, 1): 0.105, (2, 2): 0.0, (2, 3): 0.035, (3, 0): 0.210, (3, 1): 0.140, (3, 2): 0.035, (3, 3): 0.0}
    for (i, j), d in want.items():
        assert abs(orc.array_distance(xyz, i, j) - d) <= 4 * np.spacing(d)  # EXPECT_DOUBLE_EQ = 4 ULP
    assert abs(orc.array_max_distance(xyz) - 0.210) <= 4 * np.spacing(0.21)


def test_golden_ssl(orc):
    for name in ("ssl_reemc_16k.npz", "ssl_mcbeam_48k.npz"):
        g = np.load(os.path.join(G, name))
        r = orc.ssl_run(int(g["fs"]), g["xyz"], int(g["S"]), g["x"].astype(np.float64), chunk=int(g["chunk"]), want_corr=True)
        assert r["N"] == int(g["N"])
        assert np.array_equal(r["out"][: g["out"].shape[0]], g["out"])
        for k in ("doa_deg", "prob", "power", "energy", "fired_frame"):
            assert np.array_equal(r[k], g[k]), (name, k)
        if "corr_scaled" in g:
            assert np.array_equal(r["corr_scaled"], g["corr_scaled"])
    g = np.load(os.path.join(G, "ssl_reemc_16k.npz"))
    assert np.all(np.abs(g["doa_deg"][:, 0] - 30) < 1e-4)  # the scene's source sits on the 30 degree cell


def test_golden_freqgcc(orc):
    g = np.load(os.path.join(G, "freqgcc_16k.npz"))
    r = orc.freqgcc_run(int(g["fs"]), float(g["mic_dist"]), g["x"].astype(np.float64), chunk=int(g["chunk"]))
    for k in ("curves", "idx", "power"):
        assert np.array_equal(r[k], g[k]), k
    assert np.all(g["idx"] == 23)  # -21 degrees on the 3-degree grid


def test_golden_freqgcc_power_floor_pauses(orc):
    """usePowerFloor = true with a short and a long pause (fixture from the reference build): the restated _corrMemoryFactor state machine
    (BinauralLocalisation.cpp:523-561) keeps the 0.8 memory over the 10-frame pause and restarts the curve after the 52-frame one."""
    g = np.load(os.path.join(G, "freqgcc_floor_16k.npz"))
    r = orc.freqgcc_run(int(g["fs"]), float(g["mic_dist"]), g["x"].astype(np.float64), chunk=int(g["chunk"]), use_floor=True)
    assert r["n_frames"] == int(g["n_frames"])
    for k in ("fired_frame", "curves", "idx", "power"):
        assert np.array_equal(r[k], g[k]), k
    f = list(g["fired_frame"])
    assert g["idx"][f.index(17)] == 23 and g["idx"][f.index(20)] == 41      # short pause: the old curve still weighs 0.8
    assert g["idx"][f.index(79)] == 14                                        # long pause: the first voiced frame replaces the curve
    # the deterministic tracker (the `#else` branch, :501-504) rides on the same frames: DOA memory 0 -> 0.6 -> 0 after the long pause
    t = orc.freqgcc_track_run(int(g["fs"]), float(g["mic_dist"]), g["x"].astype(np.float64), chunk=int(g["chunk"]), use_floor=True, noise_preestimated=True)
    assert np.array_equal(np.nonzero(t["active"])[0], g["fired_frame"]) and np.array_equal(t["idx"][t["active"] > 0], g["idx"])
    deg = np.degrees(t["doa_rad"])
    assert abs(deg[0] + 21) < 1e-4 and abs(deg[79] + 48) < 1e-4 and -21 < deg[19] < -19 and 30 < deg[27] < 33


def test_golden_mask(orc):
    g = np.load(os.path.join(G, "mask_spatial_16k.npz"))
    x = g["x"].astype(np.float64)
    for name, method in (("full", 3), ("relative", 1), ("factor", 0), ("noisy", 4)):
        r = orc.mask_run(int(g["fs"]), float(g["mic_dist"]), float(g["lo"]), float(g["hi"]), method, 0, x, chunk=int(g["chunk"]), want_spectra=True)
        assert np.array_equal(r["out"], g[f"out_{name}"]) and np.array_equal(r["Q"], g[f"Q_{name}"]), name
    r = orc.mask_run(16000, 0.086, 500, 5000, 1, 0, x, chunk=1000, want_spectra=True)
    assert np.array_equal(r["spectra"], g["spectra_relative"])


def _band_log_power(y, f0, f1):
    """10*log10(mean square) of y band-passed to [f0, f1] (normalised frequency) — stands in for DSPONE's
    BandPassFIRFilter(256, f0, f1) + SignalPower::logPower used by test_mcarray.cpp:941-956."""
    Y = np.fft.rfft(y)
    f = np.fft.rfftfreq(len(y))
    Y[(f < f0) | (f > f1)] = 0
    return 10 * np.log10(np.mean(np.fft.irfft(Y, len(y)) ** 2) + 1e-30)


def test_mask_acceptance_spatial(orc):
    """Restated testSpatialMasking (test_mcarray.cpp:892-958): after FULL masking the 0.1 band and the 0.3
    band of the output must both sit within 70 +- 10 dB (EXPECT_LE(fabs(70-power), 10), four times)."""
    g = np.load(os.path.join(G, "mask_spatial_16k.npz"))
    x = g["x"].astype(np.float64)
    r = orc.mask_run(16000, 0.086, 500, 5000, 3, 0, x)  # FULL, BOTH
    y = np.round(r["out"][0])
    for band in ((0.05, 0.15), (0.25, 0.35)):
        assert abs(70 - _band_log_power(y, *band)) <= 10, band


def test_stft_matches_numpy(orc):
    rng = np.random.default_rng(3)
    for N in (256, 512, 1024, 2048):
        x = rng.standard_normal((3, 5 * N + 17))
        w = orc.sqrt_hann(N)
        assert np.allclose(w, np.sqrt(0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)))
        S = orc.stft(x, N, N // 2)
        T = (x.shape[1] - N) // (N // 2) + 1
        assert S.shape == (T, 3, N // 2 + 1)
        for t in (0, T - 1):
            want = np.fft.rfft(x[:, t * N // 2: t * N // 2 + N] * w, axis=1)
            assert np.allclose(S[t], want, rtol=1e-12, atol=1e-9)
        y = orc.istft(S, N, N // 2)
        assert np.allclose(y[:, N // 2: (T - 1) * N // 2], x[:, N // 2: (T - 1) * N // 2], atol=1e-9)  # COLA interior


def test_generalisation_reduces_to_reference(orc):
    """the generalised far-field table must reproduce the reference's pair delays on an
    ascending x-axis array (up to the reference's float rounding)."""
    xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    step = np.float32(5 * np.pi / 180)
    ref_tau = orc.reference_pair_tau(xyz, 48000, step)
    theta = np.array([orc.doa_idx_to_angle(d, step) for d in range(37)])
    gen = orc.pair_tau_from_mic_tau(orc.mic_tau(xyz, 48000, scenes.azimuth_dirs(theta)))
    assert ref_tau.shape == gen.shape == (6, 37)
    assert np.max(np.abs(ref_tau - gen)) < 2e-5  # float rounding of a <= 30-sample delay


def test_filter_and_sum_reduces_to_delay_and_sum(orc):
    """the filter-and-sum restatement with W[d][c][k] = exp(j k phi_c(d)) (Beamformer.cpp:59) is the delay-and-sum fan"""
    fs, N, M = 16000, 512, 6
    xs = (np.arange(M) - 2.5) * 0.05
    xyz = scenes.linear_array(xs)
    x = scenes.far_field_scene(xyz, fs, 3 * N, scenes.azimuth_dirs([0.4]), seed=9)
    S = orc.stft(x, N, N // 2)
    doas = np.deg2rad(np.arange(-90, 91, 30.0))
    k = np.arange(N // 2 + 1)
    phi = 2 * np.pi * fs / N / 346.1 * xs[None, :] * np.cos(doas[:, None] + np.pi / 2)      # [D][M]
    W = np.exp(1j * phi[:, :, None] * k[None, None, :])
    a, b = orc.fs_fan(S, N, W), orc.ds_fan(S, N, fs, xs, doas)
    assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b))


def test_channel_form_equals_pair_form(orc):
    """sum_{i<j} Re(G_ij e^{jw tau_ij}) == 1/2(|sum_m U_m e^{-jw tau_m}|^2 - M)."""
    fs, N = 48000, 512
    xyz = scenes.planar_array(3, 2, 0.04)
    dirs = scenes.az_el_dirs(np.linspace(-np.pi, np.pi, 11)[:, None], np.linspace(0.1, 1.3, 5)[None, :])
    x = scenes.far_field_scene(xyz, fs, 4 * N, dirs[17:18], seed=9)
    S = orc.stft(x, N, N // 2)
    mt = orc.mic_tau(xyz, fs, dirs)
    pair = orc.gcc_tau_frames(S, N, orc.pair_tau_from_mic_tau(mt)).sum(axis=1)
    chan = orc.srp_channel(S, N, mt, n_threads=2)
    assert np.allclose(pair, chan, rtol=1e-9, atol=1e-7)
    assert np.all(np.argmax(chan, axis=1) == 17)


def test_tdoa_lags_known_answer(orc):
    """An integer delay between two white channels must come back as that lag."""
    rng = np.random.default_rng(4)
    N, n = 1024, 8 * 1024
    s = rng.standard_normal(n + 64)
    for d in (-9, 0, 5, 28):
        x = np.stack([s[32: 32 + n], s[32 + d: 32 + d + n]])  # ch1[t] = ch0[t + d]: ch1 leads by d
        S = orc.stft(x, N, N // 2)
        curves, lags = orc.tdoa_lags(S, N, 28)
        assert curves.shape == (S.shape[0], 1, 57)
        assert np.all(lags == d), (d, lags.ravel())


def test_select_doa_edge_cases(orc):
    D = 37
    flat = np.zeros((1, D))
    idx, prob = orc.select_doa(flat, 6, 3)
    assert np.all(idx == 1) and np.all(prob == 0)  # no peaks: first max of an all-zero array, cell 0+1
    d = np.arange(D)  # one-cell spikes are removed by the median-3 (the reference's behaviour); use 5-cell bumps
    e = (np.maximum(0, 5 - np.abs(d - 10)) + 1.8 * np.maximum(0, 5 - np.abs(d - 25)))[None, :]
    idx, prob = orc.select_doa(e, 6, 3)
    assert idx[0, 0] == 25 and idx[0, 1] == 10 and prob[0, 0] > prob[0, 1] > 0 and prob[0, 2] == 0


def test_multiband_golden(orc):
    """the restatement reproduces the committed reference fixture (so the oracle is pinned on the GPU box too)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multiband_16k.npz"))
    r = orc.multiband_run(int(g["fs"]), float(g["mic_dist"]), g["x"].astype(np.float64), nbins=int(g["nbins"]), chunk=int(g["chunk"]))
    for k in ("cell", "prob", "power", "doa_deg", "hist", "band_cells", "fired_frame"):
        assert np.array_equal(r[k], g[k]), k
    # the source moves from +35 to -20 degrees: the published cells follow it (5 degree grid, cell = (deg + 90) / 5)
    assert np.median(g["cell"][5:25]) == 25 and np.median(g["cell"][-20:]) == 14


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside the GPU arm) on a tiny sample: one JSON line with the contract keys,
    no CUDA needed"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--frames", "8", "--streams", "8"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "channel_samples_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the config object is the GPU arm's: same keys, same batch (the driver compares the two arms' configs)
    sys.path.insert(0, root)
    import bench
    wl = bench.WORKLOADS["cfg2"]()
    assert line["config"] == bench.config_of(wl, 8, 8, 1)
    assert bench.config_of(wl, wl.B_default, wl.T_default, 1)["frames_per_stream_per_step"] == 750

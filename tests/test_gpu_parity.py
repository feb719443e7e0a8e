"""GPU parity tests (run on the B200 box with -m gpu): every CUDA path is driven through the C ABI and compared with the
float64 oracle on the same seeded inputs and with the golden fixtures generated from the reference build.

Tolerances (BASELINE.json north_star): integer outputs (TDOA lags, DOA cells, argmax cells, mask decisions) must be
bit-exact; float outputs must satisfy |gpu - ref| <= 1e-6 + 1e-4 * scale, where scale is the largest magnitude of the
reference in the same frame (an fp32 FFT cannot hold 1e-4 relative on individual near-zero bins)."""
import os

import numpy as np
import pytest

from mcarray_b200 import scenes

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RTOL, ATOL = 1e-4, 1e-6


def assert_close(a, ref, frame_axes, what=""):
    a = np.asarray(a); ref = np.asarray(ref)
    assert a.shape == ref.shape, (what, a.shape, ref.shape)
    scale = np.max(np.abs(ref), axis=frame_axes, keepdims=True)
    err = np.abs(a - ref)
    bad = err > ATOL + RTOL * scale
    ratio = err / (ATOL + RTOL * scale)
    # element-wise relative error (north_star's literal 1e-4 relative / 1e-6 absolute), for the record: an fp32 FFT cannot hold it on
    # near-zero bins, which is why the gate is the frame-scaled bound above
    rel = err / np.maximum(np.abs(ref), 1e-30)
    big = np.abs(ref) > 1e-3 * scale
    from conftest import record_margin
    record_margin(what=what, n=int(err.size), worst_ratio=float(np.max(ratio)) if err.size else 0.0, median_ratio=float(np.median(ratio)) if err.size else 0.0,
                  elementwise_fail_frac=float(np.mean(err > ATOL + RTOL * np.abs(ref))) if err.size else 0.0,
                  elementwise_rel_p99_above_1e3_of_scale=float(np.percentile(rel[big], 99)) if big.any() else 0.0)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} outside tolerance; worst ratio {np.max(ratio):.3g}"


def record_argmax_margin(what, ref_values, axis=-1):
    """how clear the reference's own arg-max decisions are: (best - runner-up) / |best| over the compared rows"""
    v = np.sort(np.asarray(ref_values, dtype=np.float64), axis=axis)
    best, second = np.take(v, -1, axis=axis), np.take(v, -2, axis=axis)
    m = (best - second) / np.maximum(np.abs(best), 1e-30)
    from conftest import record_margin
    record_margin(what=what + " (arg-max margin of the reference)", n=int(m.size), min_margin=float(np.min(m)), median_margin=float(np.median(m)))


@pytest.fixture(scope="module")
def mb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mcarray_b200
    return mcarray_b200


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", ["stft_kernel", "fused_stft_gcc"])
@pytest.mark.parametrize("N", [256, 512, 1024, 2048])
def test_stft_istft_parity(mb, orc, N, path):
    """K1 spectra of both analysis paths: the TMA-staged stft_kernel every processor runs, and the fused STFT->GCC kernel
    of the TDOA processor (spectra emitted on request)."""
    rng = np.random.default_rng(N)
    M, n = 3, 21 * N // 2 + 40
    x = (rng.standard_normal((M, n)) * 3000).astype(np.float32)
    if path == "stft_kernel":
        p = mb.DelayAndSumFan(16000, scenes.linear_array([0.0, 0.05, 0.1]), N, np.array([0.0, 0.3]), max_frames_per_call=64)
    else:
        p = mb.TdoaEstimator(16000, M, N, 8, max_frames_per_call=64, emit_spectra=True)
    p.process(x)
    T = p.frames_done
    S = orc.stft(x.astype(np.float64), N, N // 2)
    assert T == S.shape[0] == (n - N) // (N // 2) + 1
    assert_close(p.spectra()[0], S, (1, 2), f"stft N={N}")
    # SignalPower::FFTLogPower of each frame
    assert np.allclose(p.power_db()[0], orc.fft_log_power(S, N), rtol=0, atol=1e-3)


@pytest.mark.parametrize("N,hop,T,M,offset", [(512, 256, 13, 3, 0), (512, 128, 29, 2, 0), (512, 256, 7, 2, 1), (512, 64, 40, 1, 0),
                                              (1024, 512, 13, 3, 0), (1024, 256, 21, 2, 0), (1024, 512, 5, 2, 2), (1024, 1024, 9, 1, 0),
                                              (256, 128, 11, 2, 0), (2048, 1024, 6, 2, 0)])
def test_stft_kernel_shapes(mb, N, hop, T, M, offset):
    """mcag_k_stft on its own against numpy float64: the half-warp engine (N = 512 / 1024) and the Stockham engine (256 / 2048) with
    frame counts that leave a half-filled last work item (the idle half-warp of a warp), several hops, and sample rows that are not
    16-byte aligned (offset in floats: the element-wise staging path instead of the bulk copy); Parseval power beside it."""
    import ctypes as C
    import torch
    from mcarray_b200 import capi
    lib = capi.lib()
    rng = np.random.default_rng(N + hop + T)
    rows, n = 2 * M, (T - 1) * hop + N
    pitch = n + 8
    xh = (rng.standard_normal((rows, pitch)) * 2000).astype(np.float32)
    win = np.sqrt(0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    buf = torch.zeros(rows * pitch + 8, device="cuda")
    x = buf[offset:offset + rows * pitch]
    x.copy_(torch.from_numpy(xh.reshape(-1)))
    dwin = torch.from_numpy(win).cuda()
    tw = torch.empty(lib.mcag_k_twiddle_count(N), 2, device="cuda")
    capi.check(lib.mcag_k_twiddles(N, capi.vp(tw), None))
    KP = N // 2 + 2
    spec = torch.full((2, T, M, KP, 2), 7.0, device="cuda")
    pw = torch.zeros(2, T, M, device="cuda")
    torch.cuda.synchronize()
    capi.check(lib.mcag_k_stft(capi.vp(x), C.c_longlong(pitch), rows, M, T, N, hop, capi.vp(dwin), capi.vp(tw), capi.vp(spec), capi.vp(pw), None))
    torch.cuda.synchronize()
    got = spec.cpu().numpy()
    got = got[..., 0] + 1j * got[..., 1]
    frames = np.stack([xh[:, t * hop:t * hop + N].astype(np.float64) * win.astype(np.float64) for t in range(T)], axis=1)   # [rows][T][N]
    ref = np.fft.rfft(frames, axis=2).reshape(2, M, T, N // 2 + 1).transpose(0, 2, 1, 3)                                    # [B][T][M][K]
    assert np.all(got[..., N // 2 + 1] == 0)                                                                                # the pad bin
    assert_close(got[..., :N // 2 + 1].reshape(2 * T, M, -1), ref.reshape(2 * T, M, -1), (1, 2), f"stft kernel N={N} hop={hop}")
    k = np.arange(N // 2 + 1)
    wk = np.where((k == 0) | (k == N // 2), 1.0, 2.0)
    pref = (np.abs(ref) ** 2 * wk).sum(axis=3) / (N * N)
    np.testing.assert_allclose(pw.cpu().numpy(), pref, rtol=2e-5)


def test_tdoa_config2_parity(mb, orc):
    """BASELINE config 2: 8-mic circular array r = 0.10 m, 48 kHz, N = 1024, 28 pairs, lag window +-28."""
    fs, N, L = 48000, 1024, 28
    xyz = scenes.circular_array(8, 0.10)
    B = 3
    x = np.stack([scenes.far_field_scene(xyz, fs, 24 * 512 + N, scenes.azimuth_dirs([0.3 + 0.9 * b]), seed=scenes.stream_seed(b)) for b in range(B)])
    x32 = x.astype(np.float32)
    p = mb.TdoaEstimator(fs, 8, N, L, n_streams=B, max_frames_per_call=64, emit_curves=True)
    p.process(x32.reshape(B * 8, -1))
    lags, curves = p.lags(), p.curves()
    mism = 0
    for b in range(B):
        S = orc.stft(x32[b].astype(np.float64), N, N // 2)
        rc, rl = orc.tdoa_lags(S, N, L)
        assert_close(curves[b], rc, (2,), "gcc curves")
        mism += int(np.sum(lags[b] != rl))
    assert mism == 0, f"{mism} TDOA lags differ from the oracle"


@pytest.mark.parametrize("M,N", [(64, 1024), (30, 2048)])
def test_tdoa_large_array_channel_tiled(mb, orc, M, N):
    """arrays whose spectra do not fit one CTA's shared memory (64 microphones at N = 1024: 263 KB): STFT through HBM, then the lag
    kernel tile pair by tile pair; all M (M - 1) / 2 lags exact, curves within tolerance"""
    fs, L = 48000, 12
    xyz = scenes.circular_array(M, 0.25)
    x = scenes.far_field_scene(xyz, fs, 4 * (N // 2) + N, scenes.azimuth_dirs([1.1]), seed=M).astype(np.float32)
    p = mb.TdoaEstimator(fs, M, N, L, max_frames_per_call=8, emit_curves=True)
    p.process(x)
    S = orc.stft(x.astype(np.float64), N, N // 2)
    curves, lags = orc.tdoa_lags(S, N, L)
    assert p.frames_done == S.shape[0] and lags.shape[1] == M * (M - 1) // 2
    assert np.array_equal(p.lags()[0], lags)
    assert_close(p.curves()[0], curves, (2,), f"GCC curves, {M} microphones")


def test_tdoa_streaming_equals_one_shot(mb):
    fs, N, L = 16000, 512, 12
    xyz = scenes.circular_array(4, 0.05)
    x = scenes.far_field_scene(xyz, fs, 9000, scenes.azimuth_dirs([1.0]), seed=5).astype(np.float32)
    a = mb.TdoaEstimator(fs, 4, N, L, max_frames_per_call=64)
    a.process(x)
    want = a.lags()[0]
    b = mb.TdoaEstimator(fs, 4, N, L, max_frames_per_call=64)
    got = []
    for pos in range(0, x.shape[1], 777):
        b.process(x[:, pos:pos + 777])
        if b.frames_done:
            got.append(b.lags()[0])
    got = np.concatenate(got)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ssl_reemc_16k.npz", "ssl_mcbeam_48k.npz"])
def test_ssl_against_reference_golden(mb, name):
    """SourceSeparationAndLocalisation (mcbeam's processor) against fixtures produced by the reference's own code."""
    g = np.load(os.path.join(G, name))
    fs, S, xyz = int(g["fs"]), int(g["S"]), g["xyz"]
    x = g["x"].astype(np.float32)
    p = mb.SourceSeparationAndLocalisation(fs, xyz, S, usePowerFloor=False, max_frames_per_call=64, emit=1)
    fired = []
    p.setCallback(lambda doa, prob, power, n: fired.append((doa.copy(), prob.copy(), power)))
    outs, cells, energy, corr = [], [], [], []
    chunk = int(g["chunk"])
    for pos in range(0, x.shape[1], chunk):
        y = p.process(x[:, pos:pos + chunk])
        outs.append(y)
        if p.frames_done:
            cells.append(p.cells()[0]); energy.append(p.energy()[0]); corr.append(p.corr()[0])
    out = np.concatenate(outs, axis=1); cells = np.concatenate(cells); energy = np.concatenate(energy); corr = np.concatenate(corr)
    T = g["doa_deg"].shape[0]
    assert cells.shape == (T, S) and len(fired) == T
    # DOA cells bit-exact: the reference reports doaIdx2angle(cell) in degrees
    step = np.float32(5 * np.pi / 180)
    ref_cells = np.round((g["doa_deg"] * np.pi / 180 + np.pi / 2) / step).astype(np.int32)
    assert np.array_equal(cells, ref_cells), f"{np.sum(cells != ref_cells)} DOA cells differ"
    assert np.allclose(np.array([f[0] for f in fired]), g["doa_deg"], atol=1e-9)
    assert_close(energy, g["energy"], (1,), "energy map")
    assert_close(np.array([f[1] for f in fired]), g["prob"], (1,), "prob")
    assert np.allclose(np.array([f[2] for f in fired]), g["power"], atol=1e-3)
    if "corr_scaled" in g:
        assert_close(corr * float(1 - np.float32(0.8)), g["corr_scaled"], (1, 2), "pair correlations")
    nref = g["out"].shape[1]
    assert out.shape[1] == nref
    assert_close(out[: g["out"].shape[0]].T.reshape(-1, p.info.hop, g["out"].shape[0]), g["out"].T.reshape(-1, p.info.hop, g["out"].shape[0]), (1, 2), "separated audio")
    assert np.all(out[S:] == 0)


@pytest.mark.parametrize("form", ["pair", "channel"])
def test_ssl_16mic_against_reference_golden(mb, form):
    """BASELINE config 5 geometry (16-mic linear array, two sources) against the reference's own code: the pair form
    (gcc_tau_kernel, pair by pair like SteeringBeamforming::computeCorrelations) and the channel form on the tensor cores
    (srp_tc_kernel, what the processor picks by default for 16/32/48/64 channels) must both give the reference's DOA cells
    bit-exact and its energy map / prob / separated audio within tolerance."""
    g = np.load(os.path.join(G, "ssl_lin16_16k.npz"))
    fs, S, xyz = int(g["fs"]), int(g["S"]), g["xyz"]
    x = g["x"].astype(np.float32)
    p = mb.SourceSeparationAndLocalisation(fs, xyz, S, usePowerFloor=False, max_frames_per_call=64, srp_form={"pair": 1, "channel": 2}[form])
    assert p.info.srp_form == {"pair": 1, "channel": 2}[form]
    auto = mb.SourceSeparationAndLocalisation(fs, xyz, S, usePowerFloor=False, max_frames_per_call=64)
    assert auto.info.srp_form == 2                                         # the default for this shape is the tensor-core path
    auto.close()
    outs, cells, energy, prob = [], [], [], []
    chunk = int(g["chunk"])
    for pos in range(0, x.shape[1], chunk):
        outs.append(p.process(x[:, pos:pos + chunk]))
        if p.frames_done:
            cells.append(p.cells()[0]); energy.append(p.energy()[0]); prob.append(p.prob()[0])
    out = np.concatenate(outs, axis=1); cells = np.concatenate(cells); energy = np.concatenate(energy); prob = np.concatenate(prob)
    step = np.float32(5 * np.pi / 180)
    ref_cells = np.round((g["doa_deg"] * np.pi / 180 + np.pi / 2) / step).astype(np.int32)
    assert cells.shape == ref_cells.shape
    assert np.array_equal(cells, ref_cells), f"{np.sum(cells != ref_cells)} DOA cells differ ({form} form)"
    assert_close(energy, g["energy"], (1,), f"energy map ({form} form)")
    assert_close(prob, g["prob"], (1,), "prob")
    ref = g["out"]
    assert out.shape[1] == ref.shape[1]
    assert_close(out[:S].T.reshape(-1, p.info.hop, S), ref.T.reshape(-1, p.info.hop, S), (1, 2), "separated audio")
    assert np.all(out[S:] == 0)


def test_ssl_channel_form_needs_consistent_delays(mb):
    """the reference's scalar pair distances (SteeringBeamforming.cpp:67-73) only factor into per-microphone delays for linear
    arrays in monotonic order: a planar array keeps the pair form by default and refuses srp_form = channel."""
    from mcarray_b200 import capi
    xyz = scenes.planar_array(4, 4, 0.035)
    p = mb.SourceSeparationAndLocalisation(16000, xyz, 1, usePowerFloor=False, max_frames_per_call=8)
    assert p.info.srp_form == 1
    with pytest.raises(capi.McagError):
        mb.SourceSeparationAndLocalisation(16000, xyz, 1, usePowerFloor=False, max_frames_per_call=8, srp_form=2)


def test_ssl_power_floor_gate(mb, orc):
    fs = 16000
    xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    x = scenes.far_field_scene(xyz, fs, 5 * fs, scenes.azimuth_dirs([np.deg2rad(20)]), seed=5)
    x[:, : 3 * fs + 4000] *= 1e-3
    x = x.astype(np.float32)
    ref = orc.ssl_run(fs, xyz, 1, x.astype(np.float64), chunk=4096, use_floor=True)
    p = mb.SourceSeparationAndLocalisation(fs, xyz, 1, usePowerFloor=True, max_frames_per_call=64)
    act, cells, outs = [], [], []
    for pos in range(0, x.shape[1], 4096):
        outs.append(p.process(x[:, pos:pos + 4096]))
        if p.frames_done:
            act.append(p.active()[0]); cells.append(p.cells()[0])
    act = np.concatenate(act); cells = np.concatenate(cells); out = np.concatenate(outs, axis=1)
    assert np.array_equal(np.nonzero(act)[0], ref["fired_frame"])
    step = np.float32(5 * np.pi / 180)
    ref_cells = np.round((ref["doa_deg"] * np.pi / 180 + np.pi / 2) / step).astype(np.int32)
    assert np.array_equal(cells[act.astype(bool)], ref_cells)
    hop = p.info.hop
    assert_close(out[:1].T.reshape(-1, hop, 1), ref["out"][:1].T.reshape(-1, hop, 1), (1, 2), "gated separation")


def test_freqgcc_against_reference_golden(mb):
    g = np.load(os.path.join(G, "freqgcc_16k.npz"))
    x = g["x"].astype(np.float32)
    p = mb.FreqGCCBinauralLocalisation(int(g["fs"]), float(g["mic_dist"]), usePowerFloor=False, max_frames_per_call=64)
    curves, idx = [], []
    for pos in range(0, x.shape[1], int(g["chunk"])):
        p.process(x[:, pos:pos + int(g["chunk"])])
        if p.frames_done:
            curves.append(p.curves()[0]); idx.append(p.cells()[0])
    curves = np.concatenate(curves); idx = np.concatenate(idx)
    assert np.array_equal(idx, g["idx"])
    record_argmax_margin("FreqGCC cell", g["curves"])
    assert_close(curves, g["curves"], (1,), "smoothed GCC curve")


def test_freqgcc_power_floor_against_reference_golden(mb):
    """usePowerFloor = true with pauses, fixture from the reference build: the device _corrMemoryFactor state machine
    (BinauralLocalisation.cpp:523-561) must keep the 0.8 memory over the short pause and restart the curve after the >= 3 s one."""
    g = np.load(os.path.join(G, "freqgcc_floor_16k.npz"))
    x = g["x"].astype(np.float32)
    p = mb.FreqGCCBinauralLocalisation(int(g["fs"]), float(g["mic_dist"]), usePowerFloor=True, max_frames_per_call=8, noise_preestimated=True)
    fired = []
    p.setCallback(lambda doa, prob, power, n: fired.append(power))
    curves, idx, act = [], [], []
    for pos in range(0, x.shape[1], int(g["chunk"])):
        p.process(x[:, pos:pos + int(g["chunk"])])
        if p.frames_done:
            curves.append(p.curves()[0]); idx.append(p.cells()[0]); act.append(p.active()[0])
    curves = np.concatenate(curves); idx = np.concatenate(idx); act = np.concatenate(act).astype(bool)
    assert len(act) == int(g["n_frames"])
    assert np.array_equal(np.nonzero(act)[0], g["fired_frame"])
    assert np.array_equal(idx[act], g["idx"])
    assert_close(curves[act], g["curves"], (1,), "smoothed GCC curve, power floor on")
    assert len(fired) == len(g["fired_frame"]) and np.allclose(fired, g["power"], rtol=0, atol=1e-3)


def _floor_scene(fs=16000, hop=256):
    """quiet lead-in (the 3 s noise estimate + 10 more frames) | voiced 40 | quiet 200 (> windowsToDecay = 187) | voiced 40 | quiet 20 | voiced 30"""
    xyz = scenes.linear_array([0, 0.086])
    rng = np.random.default_rng(77)

    def voiced(nf, az, seed):
        return scenes.far_field_scene(xyz, fs, nf * hop, scenes.azimuth_dirs([np.deg2rad(az)]), seed=seed)

    def quiet(nf):
        return rng.standard_normal((2, nf * hop)) * 3.0
    return np.concatenate([quiet(105), voiced(40, 40, 11), quiet(200), voiced(40, -25, 12), quiet(20), voiced(30, 10, 13), quiet(5)], axis=1)


@pytest.mark.parametrize("chunk", [0, 3000])
def test_freqgcc_power_floor_state_machine_vs_oracle(mb, orc, chunk):
    """FreqGCC with usePowerFloor = true and the floor ESTIMATED on the device (noise_preestimated = False), N = 512: gate flags, arg-max
    cells and the deterministic tracker's DOA bit-exact against the oracle, curves and setProbability within tolerance; the scene has a
    pause longer than windowsToDecay (restart with alpha = 0) and a short one (alpha stays 0.8)."""
    x = _floor_scene()
    ref = orc.freqgcc_track_run(16000, 0.086, x, chunk=chunk, use_floor=True, noise_preestimated=False, N=512)
    T = ref["n_frames"]
    p = mb.FreqGCCBinauralLocalisation(16000, 0.086, usePowerFloor=True, max_frames_per_call=T + 1, frame_size=512, noise_preestimated=False,
                                       deterministic_tracker=True)
    step = chunk or x.shape[1]
    act, idx, curves, doa, prob = [], [], [], [], []
    for pos in range(0, x.shape[1], step):
        p.process(x[:, pos:pos + step].astype(np.float32))
        if p.frames_done:
            act.append(p.active()[0]); idx.append(p.cells()[0]); curves.append(p.curves()[0]); doa.append(p.tracked_doa()[0]); prob.append(p.prob()[0])
    act = np.concatenate(act).astype(bool); idx = np.concatenate(idx); curves = np.concatenate(curves); doa = np.concatenate(doa); prob = np.concatenate(prob)
    assert len(act) == T and 100 < act.sum() < 125
    assert np.array_equal(act, ref["active"].astype(bool))
    assert np.array_equal(idx[act], ref["idx"][act])
    assert_close(curves[act], ref["curves"][act], (1,), "smoothed GCC curve")
    assert_close(curves[~act], ref["curves"][~act], (1,), "held curve on gated frames")
    assert np.array_equal(doa, ref["doa_rad"]), np.max(np.abs(doa - ref["doa_rad"]))
    near_cut = np.abs(ref["prob"] - 0.01) < 1e-4                        # setProbability zeroes values below 0.01 (:628)
    assert np.allclose(prob[~near_cut], ref["prob"][~near_cut], rtol=1e-3, atol=1e-5)
    voiced = np.nonzero(act)[0]
    first_after_long_pause = voiced[np.nonzero(np.diff(voiced) > 187)[0][0] + 1]
    assert abs(np.degrees(doa[first_after_long_pause]) + 24) <= 1.5      # restarted: no memory of the +40 degree source


def test_multiband_against_reference_golden(mb):
    """N2 MultibandBinarualLocalisation against the fixture produced by the reference's own MultibandBinarualLocalisation.cpp:
    published cells and per-band arg-max cells bit-exact, histogram / prob within tolerance, callback deliveries in order."""
    g = np.load(os.path.join(G, "multiband_16k.npz"))
    x = g["x"].astype(np.float32)
    p = mb.MultibandBinarualLocalisation(int(g["fs"]), float(g["mic_dist"]), nbins=int(g["nbins"]), usePowerFloor=False, max_frames_per_call=64,
                                         noise_preestimated=True)
    assert p.info.n_dirs == int(g["D"]) and p.info.window_size == int(g["N"])
    fired = []
    p.setCallback(lambda doa, prob, power, n: fired.append((float(doa[0]), float(prob[0]), power)))
    cells, bcells, hist, prob = [], [], [], []
    for pos in range(0, x.shape[1], int(g["chunk"])):
        p.process(x[:, pos:pos + int(g["chunk"])])
        if p.frames_done:
            cells.append(p.cells()[0]); bcells.append(p.band_cells()[0]); hist.append(p.histogram()[0]); prob.append(p.prob()[0])
    cells = np.concatenate(cells); bcells = np.concatenate(bcells); hist = np.concatenate(hist); prob = np.concatenate(prob)
    assert np.array_equal(cells, g["cell"]), f"{np.sum(cells != g['cell'])} published cells differ"
    assert np.array_equal(bcells, g["band_cells"]), f"{np.sum(bcells != g['band_cells'])} band cells differ"
    assert_close(hist, g["hist"], (1,), "energy-weighted DOA histogram")
    assert np.allclose(prob, g["prob"], rtol=1e-4, atol=1e-6)
    assert len(fired) == len(g["cell"])
    assert np.allclose([f[0] for f in fired], g["doa_deg"], atol=1e-5)
    assert np.allclose([f[2] for f in fired], g["power"], rtol=1e-4)


def test_multiband_gate_and_streams_vs_oracle(mb, orc):
    """power gate (floor estimated on a quiet lead-in), several streams in one handle, chunked calls: fired frames, cells and
    band cells must match the oracle run of each stream"""
    fs, d, B = 16000, 0.089, 3
    xyz = scenes.linear_array([0, d])
    xs = []
    for b in range(B):
        x = scenes.far_field_scene(xyz, fs, 4 * fs + 1000, scenes.azimuth_dirs([np.deg2rad(-50 + 40 * b)]), seed=scenes.stream_seed(40 + b))
        x[:, : 3 * fs + 3000] *= 1e-3
        xs.append(np.round(x * 4))
    X = np.stack(xs).astype(np.float32)
    p = mb.MultibandBinarualLocalisation(fs, d, nbins=15, usePowerFloor=True, n_streams=B, max_frames_per_call=64)
    act, cells, bcells = [], [], []
    flat = X.reshape(B * 2, -1)
    for pos in range(0, flat.shape[1], 5000):
        p.process(flat[:, pos:pos + 5000])
        if p.frames_done:
            act.append(p.active()); cells.append(p.cells()); bcells.append(p.band_cells())
    act = np.concatenate(act, axis=1); cells = np.concatenate(cells, axis=1); bcells = np.concatenate(bcells, axis=1)
    for b in range(B):
        ref = orc.multiband_run(fs, d, X[b].astype(np.float64), nbins=15, chunk=5000, use_floor=True, noise_preestimated=False)
        on = act[b].astype(bool)
        assert 0 < ref["n_fired"] < ref["n_frames"]
        assert np.array_equal(np.nonzero(on)[0], ref["fired_frame"]), "gate decisions differ"
        assert np.array_equal(cells[b][on], ref["cell"])
        assert np.array_equal(bcells[b][on], ref["band_cells"])


def test_multiband_general_path_equals_fused_path(mb):
    """the fused band / scan / summary kernel (D <= 64, bands <= 16 bins) and the general three-kernel path (MCAG_MB_GENERAL=1) on
    the same streams, ragged chunks: cells, band cells and gate decisions identical, curves / histogram / prob within tolerance"""
    fs, d, B = 16000, 0.089, 5
    xyz = scenes.linear_array([0, d])
    X = np.stack([scenes.far_field_scene(xyz, fs, 2 * fs + 777, scenes.azimuth_dirs([np.deg2rad(-60 + 30 * b)]), seed=scenes.stream_seed(70 + b))
                  for b in range(B)]).astype(np.float32)
    flat = X.reshape(B * 2, -1)

    def run(general):
        if general:
            os.environ["MCAG_MB_GENERAL"] = "1"
        try:
            p = mb.MultibandBinarualLocalisation(fs, d, nbins=15, usePowerFloor=False, n_streams=B, max_frames_per_call=80)
        finally:
            os.environ.pop("MCAG_MB_GENERAL", None)
        out = {k: [] for k in ("cells", "band_cells", "curves", "hist", "prob")}
        for pos in range(0, flat.shape[1], 7001):
            p.process(flat[:, pos:pos + 7001])
            if p.frames_done:
                out["cells"].append(p.cells()); out["band_cells"].append(p.band_cells()); out["hist"].append(p.histogram())
                out["prob"].append(p.prob()); out["curves"].append(p.band_curves())
        p.close()
        return {k: np.concatenate(v, axis=1) for k, v in out.items()}

    a, b = run(False), run(True)
    assert a["cells"].shape[1] > 100
    assert np.array_equal(a["cells"], b["cells"]) and np.array_equal(a["band_cells"], b["band_cells"])
    assert_close(a["curves"], b["curves"], (3,), "band curves")
    assert_close(a["hist"], b["hist"], (2,), "energy histogram")
    assert np.allclose(a["prob"], b["prob"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name,method", [("full", "FULL"), ("relative", "RELATIVE"), ("factor", "FACTOR"), ("noisy", "NOISY")])
def test_mask_against_reference_golden(mb, name, method):
    g = np.load(os.path.join(G, "mask_spatial_16k.npz"))
    x = g["x"].astype(np.float32)
    p = mb.FastBinauralMasking(int(g["fs"]), float(g["mic_dist"]), float(g["lo"]), float(g["hi"]), method, "BOTH", max_frames_per_call=64, emit_trace=True)
    outs, Q = [], []
    for pos in range(0, x.shape[1], int(g["chunk"])):
        outs.append(p.process(x[:, pos:pos + int(g["chunk"])]))
        if p.frames_done:
            Q.append(p.Q()[0])
    out = np.concatenate(outs, axis=1); Q = np.concatenate(Q)
    assert_close(Q, g[f"Q_{name}"], (1,), "short-time band power")
    ref = g[f"out_{name}"]
    assert out.shape == ref.shape
    hop = p.info.hop
    assert_close(out.T.reshape(-1, hop, 2), ref.T.reshape(-1, hop, 2), (1, 2), f"masked audio ({name})")


def test_mask_decisions_vs_oracle(mb, orc):
    """every boolean mask decision on a broadband two-source scene must match the oracle"""
    fs, d = 16000, 0.086
    xyz = scenes.linear_array([0, d])
    x = scenes.far_field_scene(xyz, fs, 20 * 1024, scenes.azimuth_dirs([0.0, np.deg2rad(60)]), seed=77).astype(np.float32)
    # the fused kernel (default: spectra stay on chip, decisions kept on request) and the staged kernels (spectra fetchable)
    p = mb.FastBinauralMasking(fs, d, 500, 5000, "RELATIVE", "BOTH", max_frames_per_call=64, emit_trace=True)
    q = mb.FastBinauralMasking(fs, d, 500, 5000, "RELATIVE", "BOTH", max_frames_per_call=64, emit_spectra=True)
    yp, yq = p.process(x), q.process(x)
    N = p.info.window_size
    S = orc.stft(x.astype(np.float64), N, N // 2)
    H, fc = orc.mel_bank(N, 45, fs, 500, 5000)
    ref_spec, ref_dec, _, _ = orc.mask_frames(S, N, fs, d, 1, 0, H, fc)
    for name, proc in (("fused", p), ("staged", q)):
        dec = proc.decisions()[0]
        assert np.array_equal(dec.astype(np.int32), ref_dec), f"{name}: {np.sum(dec != ref_dec)} decisions differ"
    assert_close(q.masked_spectra()[0], ref_spec, (1, 2), "masked spectra")
    with pytest.raises(Exception):
        p.masked_spectra()                                             # the fused path never materialises them
    hop = p.info.hop
    assert_close(yp.T.reshape(-1, hop, 2), yq.T.reshape(-1, hop, 2), (1, 2), "fused vs staged audio")


@pytest.mark.parametrize("N,method,alg", [(512, "NOISY", "BOTH"), (1024, "FACTOR", "TEMPORAL"), (2048, "FULL", "SPATIAL"), (512, "RELATIVE", "BOTH")])
def test_mask_fused_kernel_equals_staged_kernels(mb, N, method, alg):
    """mask_fused_kernel (analysis + mask + synthesis in one kernel) against the stft / stats / scan / apply / istft chain on several
    streams fed in ragged chunks: same decisions, Q trace and frame powers bit for bit where both run the same FFT schedule (within
    rounding at N = 512 / 1024, where they do not), audio within tolerance."""
    fs, d, B = 16000, 0.086, 5
    xyz = scenes.linear_array([0, d])
    x = np.concatenate([scenes.far_field_scene(xyz, fs, 9 * N + 77, scenes.azimuth_dirs([0.0, np.deg2rad(20 + 15 * b)]), seed=300 + b) for b in range(B)]).astype(np.float32)
    f = mb.FastBinauralMasking(fs, d, 500, 5000, method, alg, n_streams=B, max_frames_per_call=32, frame_size=N, emit_trace=True)
    s = mb.FastBinauralMasking(fs, d, 500, 5000, method, alg, n_streams=B, max_frames_per_call=32, frame_size=N, emit_spectra=True)
    pos = 0
    for n in (N // 2 + 3, 3 * N, 5, 2 * N + 1, 10 ** 9):
        a, b_ = f.process(x[:, pos:pos + n]), s.process(x[:, pos:pos + n])
        pos += n
        assert a.shape == b_.shape and f.frames_done == s.frames_done
        if f.frames_done:
            assert np.array_equal(f.decisions(), s.decisions())
            if N in (512, 1024):   # the staged STFT runs the half-warp engine (fft16.cuh), the fused kernel the Stockham engine: rounding differs
                np.testing.assert_allclose(f.Q(), s.Q(), rtol=2e-5, atol=0)
                np.testing.assert_allclose(f.power_db(), s.power_db(), rtol=0, atol=1e-4)
            else:
                assert np.array_equal(f.Q(), s.Q())
                assert np.array_equal(f.power_db(), s.power_db())
            hop = f.info.hop
            assert_close(a.T.reshape(-1, hop, 2 * B), b_.T.reshape(-1, hop, 2 * B), (1, 2), f"fused vs staged audio N={N} {method}")
        if pos >= x.shape[1]:
            break


# ---------------------------------------------------------------------------------------------------------------------
def test_ds_fan_config3_parity(mb, orc):
    """BASELINE config 3 (reduced frame count): 32-mic linear array, 0.04 m pitch, 181 azimuths, N = 2048."""
    fs, N, M = 48000, 2048, 32
    xyz = scenes.linear_array((np.arange(M) - (M - 1) / 2) * 0.04)
    doas = np.deg2rad(np.arange(-90, 91, 1.0))
    x = scenes.far_field_scene(xyz, fs, 5 * 1024 + N, scenes.azimuth_dirs([np.deg2rad(25)]), seed=31).astype(np.float32)
    p = mb.DelayAndSumFan(fs, xyz, N, doas, max_frames_per_call=16)
    p.process(x)
    beams = p.beams()[0]
    S = orc.stft(x.astype(np.float64), N, N // 2)
    ref = orc.ds_fan(S, N, fs, xyz[:, 0], doas)
    assert_close(beams, ref, (1, 2), "beamformed spectra")
    pw = np.sum(np.abs(beams) ** 2, axis=2).mean(axis=0)
    assert abs(int(np.argmax(pw)) - (25 + 90)) <= 1     # the fan peaks at the source azimuth


def test_filter_and_sum_fan_parity(mb, orc):
    """filter-and-sum (loaded per-bin complex weights) against the oracle on random weights, and with delay phasors as weights against
    the delay-and-sum fan of the same processor family"""
    fs, N, M, D = 16000, 512, 6, 9
    xs = (np.arange(M) - 2.5) * 0.05
    xyz = scenes.linear_array(xs)
    x = scenes.far_field_scene(xyz, fs, 6 * (N // 2) + N, scenes.azimuth_dirs([0.4]), seed=9).astype(np.float32)
    S = orc.stft(x.astype(np.float64), N, N // 2)
    rng = np.random.default_rng(3)
    W = rng.standard_normal((D, M, N // 2 + 1)) + 1j * rng.standard_normal((D, M, N // 2 + 1))
    p = mb.FilterAndSumFan(fs, M, N, W, max_frames_per_call=16)
    p.process(x)
    assert_close(p.beams()[0], orc.fs_fan(S, N, W.astype(np.complex64)), (2,), "filter-and-sum beams")
    doas = np.deg2rad(np.arange(-80, 81, 20.0))
    k = np.arange(N // 2 + 1)
    phi = 2 * np.pi * fs / N / 346.1 * xs[None, :] * np.cos(doas[:, None] + np.pi / 2)
    q = mb.FilterAndSumFan(fs, M, N, np.exp(1j * phi[:, :, None] * k[None, None, :]), max_frames_per_call=16)
    r = mb.DelayAndSumFan(fs, xyz, N, doas, max_frames_per_call=16)
    q.process(x); r.process(x)
    assert_close(q.beams()[0], r.beams()[0], (2,), "filter-and-sum with delay phasors vs delay-and-sum")


def test_srp_config4_parity(mb, orc):
    """BASELINE config 4 (reduced): 64-mic 8x8 planar array, 3600-direction az x el grid, N = 1024; energy map and argmax."""
    fs, N = 48000, 1024
    xyz = scenes.planar_array(8, 8, 0.04)
    az = np.linspace(-np.pi, np.pi, 120, endpoint=False); el = np.linspace(0.05, 1.45, 30)
    dirs = scenes.az_el_dirs(az[:, None], el[None, :])
    src = 57 * 30 + 11
    x = scenes.far_field_scene(xyz, fs, 3 * 512 + N, dirs[src:src + 1], seed=41).astype(np.float32)
    p = mb.SrpPhat(fs, xyz, N, dirs, numOfSources=1, max_frames_per_call=8)
    p.process(x)
    e = p.energy()[0]
    S = orc.stft(x.astype(np.float64), N, N // 2)
    raw = orc.srp_channel(S, N, orc.mic_tau(xyz, fs, dirs), n_threads=8)
    ref, _ = orc.energy_scan(raw[:, None, :])
    assert_close(e, ref, (1,), "SRP energy map")
    assert np.array_equal(np.argmax(e, axis=1), np.argmax(ref, axis=1))
    record_argmax_margin("SRP-PHAT cell, 3600-direction grid", ref)
    assert np.all(np.argmax(e, axis=1) == src)
    idx, _ = orc.select_doa(ref, 64 * 63 // 2, 1)
    assert np.array_equal(p.cells()[0], idx)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16])
def test_host_call_overlapped_stream_groups(mb, dtype):
    """A host-buffer process call large enough to be cut into stream groups (copy / compute overlap inside the C ABI) must
    give exactly what one-stream handles give, for every sample type of the reference's process() overloads, with state
    (FIFO, overlap-add tail, energy smoothing) carried across two calls."""
    fs, B = 16000, 12
    xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    n1, n2 = 52000, 3000
    x = np.stack([scenes.far_field_scene(xyz, fs, n1 + n2, scenes.azimuth_dirs([np.deg2rad(-60 + 10 * b)]), seed=scenes.stream_seed(b)) for b in range(B)])
    x = np.round(x).astype(dtype)                                         # integers are exact in all three types
    assert B * 4 * n1 * x.itemsize >= 8 << 20 or dtype == np.int16
    big = mb.SourceSeparationAndLocalisation(fs, xyz, 1, usePowerFloor=False, n_streams=B, max_frames_per_call=256)
    flat = x.reshape(B * 4, -1)
    y1 = big.process(flat[:, :n1]); c1 = big.cells()
    y2 = big.process(flat[:, n1:]); c2 = big.cells()
    for b in range(B):
        one = mb.SourceSeparationAndLocalisation(fs, xyz, 1, usePowerFloor=False, max_frames_per_call=256)
        z1 = one.process(x[b][:, :n1]); d1 = one.cells()
        z2 = one.process(x[b][:, n1:]); d2 = one.cells()
        assert np.array_equal(c1[b], d1[0]) and np.array_equal(c2[b], d2[0])
        assert np.array_equal(y1[4 * b:4 * b + 4], z1) and np.array_equal(y2[4 * b:4 * b + 4], z2)
        one.close()


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,BT,D,N", [(64, 200, 300, 512), (32, 130, 257, 256), (16, 5, 64, 1024), (48, 128, 128, 512),
                                      (16, 300, 37, 512), (16, 128, 3, 256), (16, 129, 33, 2048), (16, 20000, 37, 512), (16, 131, 65, 512)])
def test_srp_tensor_kernel_vs_cuda_core_kernel(mb, M, BT, D, N):
    """K3 on tcgen05 (3xTF32, TMA-fed frames operand, generated steering operand) against the CUDA-core channel-form kernel on
    the same random spectra: partial frame / direction tiles, every supported microphone count, zero bins included."""
    import ctypes as C
    import torch
    from mcarray_b200 import capi
    lib = capi.lib()
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + BT)
    KP = N // 2 + 2
    spec = torch.randn(BT, M, KP, 2, device="cuda", generator=g) * 100
    spec[:, :, N // 2 + 1:] = 0
    spec[:, :, 0, 1] = 0; spec[:, :, N // 2, 1] = 0
    spec[3 % BT, 1, 7] = 0                                   # a silent bin: the nz term must follow it
    rng = np.random.default_rng(D)
    turns = np.ascontiguousarray(rng.uniform(-40, 40, size=(D, M)) / N)       # -tau/N turns per bin, |tau| up to 40 samples
    fx = torch.empty(D * M, dtype=torch.int64, device="cuda")
    capi.check(lib.mcag_k_phase_fx(capi.dp(turns), C.c_longlong(D * M), capi.vp(fx), None))
    out_c = torch.empty(BT, D, device="cuda"); out_t = torch.empty(BT, D, device="cuda")
    torch.cuda.synchronize()
    capi.check(lib.mcag_k_srp_channel(capi.vp(spec), 1, BT, M, N, capi.vp(fx), D, capi.vp(out_c), None))
    capi.check(lib.mcag_k_srp_tensor(capi.vp(spec), 1, BT, M, N, capi.vp(fx), D, capi.vp(out_t), None))
    torch.cuda.synchronize()
    a, b = out_t.cpu().numpy(), out_c.cpu().numpy()
    assert np.isfinite(a).all()
    scale = np.max(np.abs(b), axis=1, keepdims=True)
    assert np.max(np.abs(a - b) / (ATOL + RTOL * scale)) <= 1.0, np.max(np.abs(a - b) / scale)


@pytest.mark.parametrize("M,BT,D,N", [(2, 300, 61, 512), (4, 130, 37, 2048), (3, 5, 64, 256), (6, 129, 1, 1024)])
def test_gcc_tau_tensor_kernel_vs_cuda_core_kernel(mb, M, BT, D, N):
    """tau-grid GCC-PHAT of every pair as a tcgen05 GEMM over the bins (gcc_tau_tc_kernel) against the CUDA-core register-tile kernel on
    the same random spectra: partial frame tiles, every frame size, silent bins, 1..64 delays"""
    import ctypes as C
    import torch
    from mcarray_b200 import capi
    lib = capi.lib()
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + BT)
    KP, P = N // 2 + 2, M * (M - 1) // 2
    spec = torch.randn(BT, M, KP, 2, device="cuda", generator=g) * 100
    spec[:, :, N // 2 + 1:] = 0
    spec[:, :, 0, 1] = 0; spec[:, :, N // 2, 1] = 0
    spec[3 % BT, 1, 7] = 0                                   # a silent bin: PHAT of 0 is 0
    rng = np.random.default_rng(D)
    turns = np.ascontiguousarray(rng.uniform(-30, 30, size=(P, D)) / N)
    fx = torch.empty(P * D, dtype=torch.int64, device="cuda")
    capi.check(lib.mcag_k_phase_fx(capi.dp(turns), C.c_longlong(P * D), capi.vp(fx), None))
    out_c = torch.zeros(BT, P, D, device="cuda"); out_t = torch.zeros(BT, P, D, device="cuda")
    torch.cuda.synchronize()
    capi.check(lib.mcag_k_gcc_tau(capi.vp(spec), 1, BT, M, N, capi.vp(fx), D, capi.vp(out_c), None))
    capi.check(lib.mcag_k_gcc_tau_tensor(capi.vp(spec), 1, BT, M, N, capi.vp(fx), D, capi.vp(out_t), None))
    torch.cuda.synchronize()
    a, b = out_t.cpu().numpy(), out_c.cpu().numpy()
    assert np.isfinite(a).all()
    scale = np.max(np.abs(b), axis=(1, 2), keepdims=True)      # the frame's largest correlation (a single delay can sit at zero)
    assert np.max(np.abs(a - b) / (ATOL + RTOL * scale)) <= 1.0, np.max(np.abs(a - b) / scale)


@pytest.mark.parametrize("M,BT,D,N", [(32, 200, 181, 512), (16, 130, 70, 256), (64, 5, 300, 1024), (48, 128, 128, 2048)])
def test_ds_fan_tensor_kernel_vs_cuda_core_kernel(mb, M, BT, D, N):
    """K4 fan on tcgen05 (srp_tc_kernel<NKC, FAN>: 3xTF32 contraction of the raw spectra, beams written per bin) against the CUDA-core
    register-tile kernel on the same random spectra: partial frame / direction tiles, every supported microphone count, pad bin zero."""
    import ctypes as C
    import torch
    from mcarray_b200 import capi
    lib = capi.lib()
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + BT)
    KP = N // 2 + 2
    spec = torch.randn(BT, M, KP, 2, device="cuda", generator=g) * 1000
    spec[:, :, N // 2 + 1:] = 0
    spec[:, :, 0, 1] = 0; spec[:, :, N // 2, 1] = 0
    rng = np.random.default_rng(D)
    turns = np.ascontiguousarray(rng.uniform(-40, 40, size=(D, M)) / N)
    fx = torch.empty(D * M, dtype=torch.int64, device="cuda")
    capi.check(lib.mcag_k_phase_fx(capi.dp(turns), C.c_longlong(D * M), capi.vp(fx), None))
    out_c = torch.full((BT, D, KP, 2), 7.0, device="cuda"); out_t = torch.full((BT, D, KP, 2), 7.0, device="cuda")
    torch.cuda.synchronize()
    capi.check(lib.mcag_k_ds_fan(capi.vp(spec), 1, BT, M, N, capi.vp(fx), D, capi.vp(out_c), None))
    capi.check(lib.mcag_k_ds_fan_tensor(capi.vp(spec), 1, BT, M, N, capi.vp(fx), D, capi.vp(out_t), None))
    torch.cuda.synchronize()
    a, b = out_t.cpu().numpy(), out_c.cpu().numpy()
    assert np.isfinite(a).all()
    assert np.all(a[:, :, N // 2 + 1:] == 0) and np.all(b[:, :, N // 2 + 1:] == 0)         # pad bin
    scale = np.max(np.abs(b), axis=(1, 2, 3), keepdims=True)
    assert np.max(np.abs(a - b) / (ATOL + RTOL * scale)) <= 1.0, np.max(np.abs(a - b) / scale)

"""GPU edge cases (run on the B200 box with -m gpu): empty and ragged inputs, silence, capacity and configuration errors,
maximum shapes, integer saturation — the situations a drop-in for the reference's process() must survive.  Every compute
call goes through the C ABI; expectations come from the float64 oracle or from size-independent properties."""
import numpy as np
import pytest

from mcarray_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mcarray_b200
    return mcarray_b200


def _chunks(n, sizes):
    pos, i = 0, 0
    while pos < n:
        m = min(sizes[i % len(sizes)], n - pos)
        yield pos, m
        pos += m
        i += 1


def test_empty_and_short_calls_complete_no_frames(mb):
    """process() with zero samples or fewer than a frame returns no audio, fires nothing and keeps the samples for later
    (dsp::ShortTimeProcess buffers leftovers: mcabeamf.cpp:101-112 feeds arbitrary chunk sizes)"""
    fs = 16000
    xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    x = scenes.far_field_scene(xyz, fs, 3000, scenes.azimuth_dirs([0.4]), seed=3).astype(np.float32)
    p = mb.SourceSeparationAndLocalisation(fs, xyz, 1, usePowerFloor=False, max_frames_per_call=16)
    fired = []
    p.setCallback(lambda *a: fired.append(a))
    y = p.process(x[:, :0])
    assert y.shape == (4, 0) and p.frames_done == 0
    y = p.process(x[:, :300])                                   # N = 512: not a frame yet
    assert y.shape == (4, 0) and p.frames_done == 0 and not fired
    y = p.process(x[:, 300:512])                                # exactly one frame now
    assert y.shape == (4, 256) and p.frames_done == 1 and len(fired) == 1
    one = mb.SourceSeparationAndLocalisation(fs, xyz, 1, usePowerFloor=False, max_frames_per_call=16)
    z = one.process(x[:, :512])
    assert np.array_equal(y, z)


@pytest.mark.parametrize("kind", ["ssl", "mask", "freqgcc", "multiband", "tdoa"])
def test_ragged_chunks_equal_one_shot(mb, kind):
    """any chunking (primes, a chunk of 1 sample, chunks longer than several frames) gives bit-identical results to one call"""
    fs = 16000
    n = 7000
    if kind in ("ssl", "tdoa"):
        xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    else:
        xyz = scenes.linear_array([0, 0.086])
    x = np.round(scenes.far_field_scene(xyz, fs, n, scenes.azimuth_dirs([0.6, -0.3]), seed=17)).astype(np.float32)

    def make():
        if kind == "ssl":
            return mb.SourceSeparationAndLocalisation(fs, xyz, 2, usePowerFloor=True, max_frames_per_call=32)
        if kind == "mask":
            return mb.FastBinauralMasking(fs, 0.086, 500, 5000, "NOISY", "BOTH", max_frames_per_call=32, frame_size=512, emit_trace=True)
        if kind == "freqgcc":
            return mb.FreqGCCBinauralLocalisation(fs, 0.086, usePowerFloor=True, max_frames_per_call=32, frame_size=512, noise_preestimated=False)
        if kind == "multiband":
            return mb.MultibandBinarualLocalisation(fs, 0.086, usePowerFloor=True, max_frames_per_call=32)
        return mb.TdoaEstimator(fs, 4, 512, 12, max_frames_per_call=32, emit_curves=True)

    def results(p):
        if kind == "ssl":
            return [p.cells()[0], p.prob()[0], p.energy()[0], p.active()[0]]
        if kind == "mask":
            return [p.Q()[0], p.decisions()[0]]
        if kind == "freqgcc":
            return [p.cells()[0], p.curves()[0], p.active()[0]]
        if kind == "multiband":
            return [p.cells()[0], p.band_cells()[0], p.histogram()[0], p.prob()[0], p.active()[0]]
        return [p.lags()[0], p.curves()[0]]

    a = make()
    ya = a.process(x)
    ra = results(a)
    b = make()
    ys, rs = [], []
    for pos, m in _chunks(n, [1, 997, 13, 2311, 255, 257, 512]):
        ys.append(b.process(x[:, pos:pos + m]))
        if b.frames_done:
            rs.append(results(b))
    yb = np.concatenate(ys, axis=1)
    assert np.array_equal(ya, yb)
    for i, want in enumerate(ra):
        got = np.concatenate([r[i] for r in rs])
        assert np.array_equal(got, want), (kind, i)


def test_silence_and_dc_match_the_oracle(mb, orc):
    """all-zero frames (PHAT of 0 is 0, energy map flat) and a constant offset (only the DC bin is non-zero): cells, energy and audio
    must be what the float64 restatement gives, without NaNs"""
    fs = 16000
    xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    x = np.zeros((4, 4096))
    x[:, 2048:] = 1000.0
    x[:, 3000:] += scenes.far_field_scene(xyz, fs, 1096, scenes.azimuth_dirs([0.3]), seed=9)
    p = mb.SourceSeparationAndLocalisation(fs, xyz, 1, usePowerFloor=False, max_frames_per_call=32)
    y = p.process(x.astype(np.float32))
    ref = orc.ssl_run(fs, xyz, 1, x.astype(np.float32).astype(np.float64))
    cells = np.round((ref["doa_deg"] * np.pi / 180 + np.pi / 2) / np.float32(5 * np.pi / 180)).astype(np.int32)
    assert np.isfinite(y).all() and np.isfinite(p.energy()).all() and np.isfinite(p.prob()).all()
    assert np.array_equal(p.cells()[0][:7], cells[:7])            # the all-zero frames: flat map, reference picks cell 1 by its tie rule
    assert np.max(np.abs(y[0] - ref["out"][0])) <= 1e-6 + 1e-4 * np.max(np.abs(ref["out"][0]))
    # two-channel kinds on pure silence
    z = np.zeros((2, 4096), dtype=np.float32)
    m = mb.FastBinauralMasking(fs, 0.086, 500, 5000, "RELATIVE", "BOTH", max_frames_per_call=32, frame_size=512)
    assert np.array_equal(m.process(z), np.zeros((2, m.frames_done * 256), dtype=np.float32))
    t = mb.TdoaEstimator(fs, 2, 512, 10, max_frames_per_call=32)
    t.process(z)
    assert np.all(t.lags() == -10)                                  # flat curve: wipp::maxidx returns the first entry


def test_capacity_and_configuration_errors(mb):
    """errors are reported as status codes + mcag_last_error (McagError here, MCArrayException in C++), never a crash or a silent fallback"""
    from mcarray_b200 import capi
    fs = 16000
    xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    p = mb.TdoaEstimator(fs, 4, 512, 12, max_frames_per_call=4)
    x = np.zeros((4, 512 + 256 * 4), dtype=np.float32)              # 5 frames > max_frames_per_call
    with pytest.raises(capi.McagError, match="max_frames_per_call"):
        p.process(x)
    p.process(x[:, : 512 + 256 * 3])                                # the handle is still usable afterwards
    assert p.frames_done == 4
    with pytest.raises(capi.McagError):                             # FastBinauralMasking.cpp:88-91
        mb.Processor(kind=capi.KIND_MASK, sample_rate=fs, frame_size=512, hop=256, n_channels=3)
    with pytest.raises(capi.McagError, match="frame_size"):
        mb.TdoaEstimator(fs, 4, 500, 12)
    with pytest.raises(capi.McagError, match="hop"):
        mb.TdoaEstimator(fs, 4, 512, 12, hop=100)
    # 64 x 2048-point spectra do not fit one CTA's shared memory: the processor runs the channel-tiled lag kernel instead of failing
    big = mb.TdoaEstimator(48000, 64, 2048, 30, max_frames_per_call=2)
    big.process(np.zeros((64, 2048), dtype=np.float32))
    assert big.frames_done == 1 and np.all(big.lags() == -30)      # silence: flat curves, the first maximum is the lowest lag
    with pytest.raises(capi.McagError, match="device"):
        mb.Processor(kind=capi.KIND_TDOA, device=99, sample_rate=fs, frame_size=512, hop=256, n_channels=2, max_lag=4)
    out = np.zeros(4, dtype=np.int32)
    with pytest.raises(capi.McagError):                             # fetching a result this kind does not produce
        capi.check(capi.lib().mcag_fetch(p.handle, capi.OUT_BEAMS, out.ctypes.data, 16))


def test_largest_shapes(mb, orc):
    """the largest frame size with the widest supported tensor-core array, and the longest lag window of the pruned inverse transform"""
    fs, N, M = 48000, 2048, 64
    xyz = scenes.planar_array(8, 8, 0.04)
    dirs = scenes.az_el_dirs(np.linspace(-3, 3, 40)[:, None], np.array([0.3, 1.0])[None, :])
    x = scenes.far_field_scene(xyz, fs, N + 1024, dirs[33:34], seed=5).astype(np.float32)
    p = mb.SrpPhat(fs, xyz, N, dirs, max_frames_per_call=4)
    p.process(x)
    S = orc.stft(x.astype(np.float64), N, N // 2)
    raw = orc.srp_channel(S, N, orc.mic_tau(xyz, fs, dirs), n_threads=8)
    ref, _ = orc.energy_scan(raw[:, None, :])
    assert np.array_equal(np.argmax(p.energy()[0], axis=1), np.argmax(ref, axis=1))
    assert np.all(np.argmax(p.energy()[0], axis=1) == 33)
    # lag windows up to N/2 - 1: both the output-pruned last pass (window inside the lowest / highest NC/R outputs) and the full transform
    xyz4 = scenes.circular_array(4, 0.2)
    x4 = scenes.far_field_scene(xyz4, fs, 1024 + 512 * 5, scenes.azimuth_dirs([1.1]), seed=6).astype(np.float32)
    for L in (1, 63, 126, 127, 300, 511):
        t = mb.TdoaEstimator(fs, 4, 1024, L, max_frames_per_call=8, emit_curves=True)
        t.process(x4)
        S4 = orc.stft(x4.astype(np.float64), 1024, 512)
        rc, rl = orc.tdoa_lags(S4, 1024, L)
        assert np.array_equal(t.lags()[0], rl), L
        assert np.max(np.abs(t.curves()[0] - rc)) <= 1e-6 + 1e-4 * np.max(np.abs(rc)), L


def test_int16_output_saturates(mb):
    """the int16 overload of process() (test_mcarray.cpp:937) rounds to nearest and clips instead of wrapping"""
    fs = 16000
    x = np.zeros((2, 2048), dtype=np.int16)
    x[:, 600:1400] = 32767
    p = mb.FastBinauralMasking(fs, 0.086, 500, 5000, "NOTHING", "BOTH", max_frames_per_call=16, frame_size=512)
    y = p.process(x)
    assert y.dtype == np.int16 and y.shape == (2, p.frames_done * 256)
    # NOTHING passes the signal through (FastBinauralMasking.cpp:130-134): analysis + synthesis reconstruct it after the first hop
    assert np.max(np.abs(y[:, 256:1536].astype(np.int32) - x[:, 256:1536].astype(np.int32))) <= 1
    assert y.max() == 32767

"""Known-answer tests at the FULL sizes of the BASELINE.json configurations bench.py runs (too large for the float64 oracle in a test,
so they use properties that do not depend on size): integer delays must come back as exactly those lags, a source on a grid direction
must be the arg-max cell of every frame, batched streams must equal the same streams run alone, and overlap-add must reconstruct."""
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


def test_cfg2_full_size_integer_delays_come_back_exactly(mb):
    """config 2 at bench size: 64 streams x 8 mics x 750 frames of 1024 samples, 28 pairs each (1.3 M lags).  Every channel is the
    same white source delayed by an integer number of samples, so the GCC-PHAT peak of pair (i, j) is a clean delta at d_i - d_j."""
    fs, N, hop, M, L, B, T = 48000, 1024, 512, 8, 28, 64, 750
    n = N + (T - 1) * hop
    rng = np.random.default_rng(2024)
    x = np.empty((B * M, n), dtype=np.float32)
    delays = rng.integers(-14, 15, size=(B, M))                        # |d_j - d_i| <= 28 = the lag window
    for b in range(B):
        s = (rng.standard_normal(n + 64) * 3000).astype(np.float32)
        for m in range(M):
            x[b * M + m] = s[32 - delays[b, m]: 32 - delays[b, m] + n]  # x_m[t] = s[t - d_m]
    p = mb.TdoaEstimator(fs, M, N, L, n_streams=B, max_frames_per_call=T)
    p.process(x)
    assert p.frames_done == T
    lags = p.lags()                                                    # [B][T][P]
    pi, pj = np.array([(i, j) for i in range(M) for j in range(i + 1, M)]).T
    # convention C5: G = X_i conj(X_j) = |S|^2 e^{-j w (d_i - d_j)} and corr[tau] = sum G e^{+j w tau} peaks at tau = d_i - d_j
    want = (delays[:, pi] - delays[:, pj])[:, None, :]
    assert lags.shape == (B, T, 28)
    assert np.array_equal(lags, np.broadcast_to(want, lags.shape)), f"{np.sum(lags != want)} of {lags.size} lags differ"


def test_cfg5_full_size_batch_equals_single_streams_and_finds_the_source(mb):
    """config 5 at per-GPU bench size: 128 streams x 16 mics x 125 frames through SourceSeparationAndLocalisation (channel form on the
    tensor cores).  The source of every stream sits on a grid direction: after the 0.8 smoothing has settled every frame's cell is that
    direction; streams taken out of the batch and run alone give bit-identical cells and audio."""
    fs, M, B, T = 16000, 16, 128, 125
    xyz = scenes.linear_array((np.arange(M) - (M - 1) / 2) * 0.035)
    n = 512 + (T - 1) * 256
    cells_want = 2 + (np.arange(B) * 7) % 33                           # grid cells 2..34 (-80 .. 80 degrees)
    base = {}
    x = np.empty((B * M, n), dtype=np.float32)
    for b in range(B):
        c = int(cells_want[b])
        if c not in base:
            base[c] = scenes.far_field_scene(xyz, fs, n + 4000, scenes.azimuth_dirs([np.deg2rad(-90 + 5 * c)]), seed=scenes.stream_seed(c)).astype(np.float32)
        off = 31 * (b // 33)
        x[b * M:(b + 1) * M] = base[c][:, off:off + n]
    p = mb.SourceSeparationAndLocalisation(fs, xyz, 1, usePowerFloor=False, n_streams=B, max_frames_per_call=T)
    assert p.info.srp_form == 2
    y = p.process(x)
    cells = p.cells()[:, :, 0]
    assert cells.shape == (B, T) and y.shape == (B * M, T * 256)
    assert np.array_equal(cells[:, 10:], np.broadcast_to(cells_want[:, None], (B, T - 10))), "a stream lost its source direction"
    for b in (0, 57, 127):
        one = mb.SourceSeparationAndLocalisation(fs, xyz, 1, usePowerFloor=False, max_frames_per_call=T)
        z = one.process(x[b * M:(b + 1) * M])
        assert np.array_equal(one.cells()[0, :, 0], cells[b])
        assert np.array_equal(z, y[b * M:(b + 1) * M])
        one.close()
    assert np.all(y.reshape(B, M, -1)[:, 1:] == 0)                     # channels >= numOfSources are zeroed


def test_cfg4_full_size_argmax_is_the_source_cell(mb):
    """config 4 at bench size: 4 arrays x 64 mics x 256 frames, 3600-direction az x el grid on tcgen05.  Each array hears one source
    placed on a grid direction: the arg-max of the smoothed SRP-PHAT map is that cell in every frame."""
    fs, N, B, T = 48000, 1024, 4, 256
    xyz = scenes.planar_array(8, 8, 0.04)
    az = np.linspace(-np.pi, np.pi, 120, endpoint=False); el = np.linspace(0.05, 1.45, 30)
    dirs = scenes.az_el_dirs(az[:, None], el[None, :])
    n = N + (T - 1) * 512
    src = [(b * 997 + 57 * 30 + 11) % len(dirs) for b in range(B)]
    x = np.concatenate([scenes.far_field_scene(xyz, fs, n, dirs[s:s + 1], seed=scenes.stream_seed(b)) for b, s in enumerate(src)]).astype(np.float32)
    p = mb.SrpPhat(fs, xyz, N, dirs, numOfSources=1, n_streams=B, max_frames_per_call=T)
    p.process(x)
    e = p.energy()
    assert e.shape == (B, T, 3600) and np.isfinite(e).all()
    assert np.array_equal(np.argmax(e, axis=2), np.broadcast_to(np.array(src)[:, None], (B, T)))


def test_cfg1_full_size_masking_passthrough_reconstructs(mb):
    """config 1 masking chain at bench size (2048 stereo streams x 125 frames) with method NOTHING: analysis -> synthesis with the
    sqrt-Hann pair at 50 % overlap is the identity after the first hop (w^2 overlap-adds to 1), for every stream."""
    fs, B, T = 16000, 2048, 125
    n = 512 + (T - 1) * 256
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((B * 2, n)) * 2000).astype(np.float32)
    p = mb.FastBinauralMasking(fs, 0.086, 500, 5000, "NOTHING", "BOTH", n_streams=B, max_frames_per_call=T, frame_size=512)
    y = p.process(x)
    assert y.shape == (B * 2, T * 256)
    err = np.max(np.abs(y[:, 256:] - x[:, 256:T * 256]))
    assert err <= 1e-6 + 1e-4 * np.max(np.abs(x)), err

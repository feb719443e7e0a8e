"""Generates tests/golden/*.npz from the REFERENCE BUILD (oracle/_ref/libmcarray_ref.so = the reference's
own sources compiled in this container against the DSPONE/WIPP stand-in; `make -C oracle ref`).
Run here (needs /root/reference):   python tests/golden/make_golden.py
The fixtures are committed; on the GPU box the tests read them, never /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
import orc  # noqa: E402
from mcarray_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pause_scene(fs=16000, hop=1024):
    """voiced(-21 deg) 8 frames | quiet 10 | voiced(+33) 10 | quiet 52 | voiced(-48) 10 | quiet 3; int16 PCM (the pauses are sparse +-1 LSB
    dither, about -13 dB against the forced 0 dB floor)"""
    xyz = scenes.linear_array([0, 0.086])

    def voiced(nf, az, seed):
        return scenes.far_field_scene(xyz, fs, nf * hop, scenes.azimuth_dirs([np.deg2rad(az)]), seed=seed)

    def quiet(nf, seed):
        return np.random.default_rng(seed).choice([-1.0, 0.0, 1.0], size=(2, nf * hop), p=[0.05, 0.9, 0.05])
    x = np.concatenate([voiced(8, -21, 1), quiet(10, 2), voiced(10, 33, 3), quiet(52, 4), voiced(10, -48, 5), quiet(3, 6)], axis=1)
    return np.round(x)


def main():
    orc.build(ref=True)
    assert orc.have_ref(), "reference build unavailable"
    # 1. SourceSeparationAndLocalisation, Reem-C array (test_mcarray.cpp:397), 2 sources asked, 1 present
    fs = 16000
    xyz = scenes.linear_array([0, 0.07, 0.175, 0.21])
    x = np.round(scenes.far_field_scene(xyz, fs, 6144, scenes.azimuth_dirs([np.deg2rad(30)]), seed=101))
    r = orc.ssl_run(fs, xyz, 2, x, chunk=1000, want_corr=True, prefix="ref")
    np.savez_compressed(os.path.join(HERE, "ssl_reemc_16k.npz"), fs=fs, xyz=xyz, S=2, x=x.astype(np.int16), chunk=1000,
                        out=r["out"], doa_deg=r["doa_deg"], prob=r["prob"], power=r["power"], energy=r["energy"],
                        corr_scaled=r["corr_scaled"], fired_frame=r["fired_frame"], N=r["N"])
    # 2. mcbeam's own configuration (mcabeamf.cpp:182-194), 48 kHz -> N = 2048
    fs = 48000
    xyz = scenes.linear_array([-2.25, -1.25, 1.25, 2.25])
    x = np.round(scenes.far_field_scene(xyz, fs, 5 * 2048, scenes.azimuth_dirs([np.deg2rad(-20)]), seed=102))
    r = orc.ssl_run(fs, xyz, 1, x, chunk=1024, prefix="ref")
    np.savez_compressed(os.path.join(HERE, "ssl_mcbeam_48k.npz"), fs=fs, xyz=xyz, S=1, x=x.astype(np.int16), chunk=1024,
                        out=r["out"][:1], doa_deg=r["doa_deg"], prob=r["prob"], power=r["power"], energy=r["energy"],
                        fired_frame=r["fired_frame"], N=r["N"])
    # 2b. BASELINE config 5 geometry: 16-mic linear array, 0.035 m pitch, 16 kHz (the shape whose pair sum the GPU build runs in channel
    #     form on the tensor cores); two sources asked, two present
    fs = 16000
    xyz = scenes.linear_array((np.arange(16) - 7.5) * 0.035)
    x = np.round(scenes.far_field_scene(xyz, fs, 512 + 47 * 256 + 100, scenes.azimuth_dirs([np.deg2rad(-35), np.deg2rad(50)]), seed=104))
    r = orc.ssl_run(fs, xyz, 2, x, chunk=3000, prefix="ref")
    np.savez_compressed(os.path.join(HERE, "ssl_lin16_16k.npz"), fs=fs, xyz=xyz, S=2, x=x.astype(np.int16), chunk=3000,
                        out=r["out"][:2].astype(np.float32), doa_deg=r["doa_deg"], prob=r["prob"], power=r["power"], energy=r["energy"],
                        fired_frame=r["fired_frame"], N=r["N"])
    # 3. FreqGCCBinauralLocalisation, 0.086 m (test_mcarray.cpp:284)
    fs = 16000
    xyz = scenes.linear_array([0, 0.086])
    x = np.round(scenes.far_field_scene(xyz, fs, 10 * 1024, scenes.azimuth_dirs([np.deg2rad(-21)]), seed=103))
    r = orc.freqgcc_run(fs, 0.086, x, chunk=2048, prefix="ref")
    np.savez_compressed(os.path.join(HERE, "freqgcc_16k.npz"), fs=fs, mic_dist=0.086, x=x.astype(np.int16), chunk=2048,
                        curves=r["curves"], idx=r["idx"], power=r["power"], N=r["N"])
    # 3a. the same processor with usePowerFloor = true and pauses: exercises the _corrMemoryFactor / _silenceFramesCounter state machine
    #     (BinauralLocalisation.cpp:523-561).  N = 2048, hop = 1024 -> windowsToDecay = 46 frames: the 10-frame pause keeps the 0.8
    #     memory, the 52-frame pause resets it (the next voiced frame replaces the curve).  _noiseEstimated is forced (floor 0 dB), the
    #     only way the reference build runs (oracle/capi.h).  The reference's particle filter also fires the callback on silent frames
    #     inside the decay window (:540-544); those deliveries carry power <= floor and are dropped here.
    x = pause_scene()
    r = orc.freqgcc_run(fs, 0.086, x, chunk=2500, use_floor=True, prefix="ref")
    voiced = r["power"] > 0
    np.savez_compressed(os.path.join(HERE, "freqgcc_floor_16k.npz"), fs=fs, mic_dist=0.086, x=x.astype(np.int16), chunk=2500,
                        curves=r["curves"][voiced], idx=r["idx"][voiced], power=r["power"][voiced], fired_frame=r["fired_frame"][voiced],
                        n_frames=r["n_frames"], N=r["N"])
    # 3b. MultibandBinarualLocalisation, 0.086 m, 15 linear bands; source moves from +35 to -20 degrees half way
    fs = 16000
    xyz = scenes.linear_array([0, 0.086])
    n = 30 * 256 + 512
    xa = scenes.far_field_scene(xyz, fs, n, scenes.azimuth_dirs([np.deg2rad(35)]), seed=105)
    xb = scenes.far_field_scene(xyz, fs, n, scenes.azimuth_dirs([np.deg2rad(-20)]), seed=106)
    x = np.round(np.concatenate([xa, xb], axis=1))
    r = orc.multiband_run(fs, 0.086, x, nbins=15, chunk=3000, prefix="ref")
    np.savez_compressed(os.path.join(HERE, "multiband_16k.npz"), fs=fs, mic_dist=0.086, nbins=15, x=x.astype(np.int16), chunk=3000,
                        cell=r["cell"], prob=r["prob"], power=r["power"], doa_deg=r["doa_deg"], hist=r["hist"], band_cells=r["band_cells"],
                        fired_frame=r["fired_frame"], N=r["N"], D=r["D"])
    # 4. FastBinauralMasking on the reference's spatial-masking test signal (test_mcarray.cpp:908-929)
    n = 5 * 1024
    i = np.arange(n)
    sig = np.round(5000 * np.cos(2 * np.pi * 0.1 * i)); inter = np.round(5000 * np.cos(2 * np.pi * 0.3 * i))
    L = sig + inter; R = sig.copy(); R[: n - 6] += inter[6:]
    x = np.stack([L, R])
    fx = {}
    for name, method in (("full", 3), ("relative", 1), ("factor", 0), ("noisy", 4)):
        r = orc.mask_run(16000, 0.086, 500, 5000, method, 0, x, chunk=1000, want_spectra=True, prefix="ref")
        fx[f"out_{name}"] = r["out"]; fx[f"Q_{name}"] = r["Q"]
        if name == "relative":
            fx["spectra_relative"] = r["spectra"]
    np.savez_compressed(os.path.join(HERE, "mask_spatial_16k.npz"), fs=16000, mic_dist=0.086, lo=500.0, hi=5000.0,
                        x=x.astype(np.int16), chunk=1000, N=r["N"], **fx)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — channel-samples/s and xRT of the mcarray hot path on B200 (driver contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg5|...] [--also cfg5,cfg4|none] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic multichannel audio:
  cfg2 (default, the configuration the BASELINE.json metric is quoted on that fits one GPU):
       B array streams x 8-mic circular array, 48 kHz, N = 1024, hop = 512: STFT -> GCC-PHAT on all 28 pairs ->
       integer-lag TDOA, T frames per stream per step.
  cfg5: B streams x 16-mic linear array, 16 kHz, N = 512: SourceSeparationAndLocalisation (STFT -> GCC-PHAT tau grid ->
       SRP energy -> selectDOA -> delay-and-sum -> overlap-add), the literal mcbeam processor batched over streams.
Streams are independent, so N > 1 shards them across ranks with no collective (weak scaling: B streams PER GPU).

  value  whole-job channel-samples/s with the input already resident in HBM (mcag_process_device_f32)
  e2e    the same metric through the host-buffer C-ABI call from pinned memory + result fetch, host<->device copies inside the timed
         region; the default leg carries 16-bit PCM (mcag_process_packed_s16, the reference's process(int16_t*) overload), `e2e_f32`
         the fp32 block; `h2d_ceiling_gbs` is the bare pinned copy rate measured in the same run
  roofline      dominant kernel (largest share of the per-kernel CUDA-event times recorded on the handle's stream); `bound` names the
                roofline that binds it (hbm / fp32 / tensor) and the other fractions sit beside it
  sustained     the same step repeated back to back for ~1.5 s (clocks and power under a load longer than the timed region)
  workloads     short runs of the other BASELINE configs in the same process: cfg5, cfg4 (tensor-pipe fraction), cfg3, cfg1m, cfg1l, and at
                N > 1 cfg4s (grid sharded over the GPUs: ms/step, all-reduce time, speed-up over one GPU of the same node)
  cpu_baseline  the CPU path on a bounded sample, all host cores: oracle/_ref (the reference's own sources) where the workload is a
                reference class, else the float64 restatement

--impl reference times the CPU path on the SAME config (same streams and frames per step as the GPU arm; cfg3 / cfg4 bounded).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle", "py")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC, UNIT = "channel_samples_per_sec", "channel-samples/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------------
class Workload:
    result_what = None   # MCAG_OUT_* id of the per-step result the e2e leg reads back (None: the audio the call itself returns)

    def fetch_bytes(self, p, B, T):
        return self.result_bytes(p, B, T)

    def fetch_result(self, p):
        """the step's result into a pinned host buffer kept across steps (a fresh pageable array per step costs page faults and a staged copy:
        0.4 ms for cfg2's 5.4 MB of lags, 150 ms for cfg3's 760 MB of beams)"""
        if self.result_what is None:
            return None
        import torch
        from mcarray_b200 import capi
        nbytes = self.fetch_bytes(p, p.info.n_streams, p.frames_done)
        if getattr(self, "_pin", None) is None or self._pin.numel() < nbytes:
            self._pin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        capi.check(capi.lib().mcag_fetch(p.handle, self.result_what, C.c_void_p(self._pin.data_ptr()), C.c_longlong(nbytes)))
        return self._pin

    name = ""

    def scene(self, stream_id, n):
        raise NotImplementedError

    def host_input(self, rank, B, n, unique=8):
        """[B*M][n] float32.  `unique` distinct seeded scenes per rank are generated in float64 and rotated in time for the
        remaining streams (every stream still has its own samples at its own addresses)."""
        from mcarray_b200 import scenes  # noqa: F401
        base = [self.scene(rank * B + u, n).astype(np.float32) for u in range(min(unique, B))]
        x = np.empty((B * self.M, n), dtype=np.float32)
        for b in range(B):
            src = base[b % len(base)]
            x[b * self.M:(b + 1) * self.M] = np.roll(src, 1009 * (b // len(base)), axis=1)
        return x


class Cfg2(Workload):
    """8-mic circular array GCC-PHAT TDOA on all 28 pairs, 48 kHz, 1024-sample frames (BASELINE.json configs[1])."""
    name = "cfg2: 8-mic circular array r=0.10 m, GCC-PHAT TDOA on all 28 pairs (lags +-28), 48 kHz, N=1024, hop=512"
    fs, N, hop, M, max_lag = 48000, 1024, 512, 8, 28
    B_default, T_default = 64, 750
    bound = {"stft_gcc": "fp32"}     # 67 flop per compulsory byte, ridge 11: FP32 (issue / shared-memory) bound, not HBM
    flops_note = ("algorithmic FFT count of SURVEY.md 8d: (M + P) real FFTs at 2.5 N log2 N, PHAT whitening 8 flop x M x K, cross-spectrum 6 flop x P x K; "
                  "the pruned / decimated inverse the kernel actually runs is NOT discounted")

    def kernel_flops_per_frame(self):
        M, N, K, P = self.M, self.N, self.N // 2 + 1, self.M * (self.M - 1) // 2
        f = (M + P) * fft_flops(N) + 8 * M * K + 6 * P * K
        return {"stft_gcc": f, "tdoa": f}

    def scene(self, stream_id, n):
        from mcarray_b200 import scenes
        xyz = scenes.circular_array(self.M, 0.10)
        az = -np.pi + 2 * np.pi * ((stream_id * 0.6180339887) % 1.0)
        return scenes.far_field_scene(xyz, self.fs, n, scenes.azimuth_dirs([az]), seed=scenes.stream_seed(stream_id))

    def make(self, mb, B, T):
        return mb.TdoaEstimator(self.fs, self.M, self.N, self.max_lag, n_streams=B, max_frames_per_call=T)

    def result_bytes(self, p, B, T):
        return B * T * p.info.n_pairs * 4

    result_what = 6   # MCAG_OUT_LAGS

    # algorithmic (compulsory) HBM bytes per frame per stream for each kernel of the chain, fp32
    def kernel_bytes_per_frame(self):
        M, hop, K, P = self.M, self.hop, self.N // 2 + 1, self.M * (self.M - 1) // 2
        return {"stft": 4 * M * hop + 8 * M * K, "tdoa": 8 * M * K + 4 * P, "stft_gcc": 4 * M * hop + 4 * P}

    def pipeline_bytes_per_frame(self):
        return 4 * self.M * self.hop + 4 * (self.M * (self.M - 1) // 2)   # production mode, 16 496 B

    def cpu_run(self, orc, x64, n_threads):
        return orc.tdoa_pipeline(x64, self.N, self.hop, self.max_lag, n_threads=n_threads)


class Cfg5(Workload):
    """1024 independent 16-mic array streams of GCC-PHAT + DS beamforming + overlap-add (BASELINE.json configs[4]): 128 per GPU."""
    name = "cfg5: 16-mic linear array 0.035 m pitch, SourceSeparationAndLocalisation (GCC-PHAT 37-cell grid + DS + OLA), 16 kHz, N=512, hop=256"
    fs, N, hop, M = 16000, 512, 256, 16
    audio_channels = 1
    B_default, T_default = 128, 125
    dominant_hint = "gcc_tau"

    def xyz(self):
        from mcarray_b200 import scenes
        return scenes.linear_array((np.arange(self.M) - (self.M - 1) / 2) * 0.035)

    def scene(self, stream_id, n):
        from mcarray_b200 import scenes
        az = np.deg2rad(-80 + 5 * (stream_id * 7 % 33))
        return scenes.far_field_scene(self.xyz(), self.fs, n, scenes.azimuth_dirs([az]), seed=scenes.stream_seed(stream_id))

    def make(self, mb, B, T):
        return mb.SourceSeparationAndLocalisation(self.fs, self.xyz(), 1, usePowerFloor=False, n_streams=B, max_frames_per_call=T)

    def result_bytes(self, p, B, T):
        return B * T * 4 + B * T * self.hop * 4

    def fetch_bytes(self, p, B, T):
        return B * T * 4            # the cells; the audio comes back through the call itself

    result_what = 4   # MCAG_OUT_CELL

    def kernel_bytes_per_frame(self):
        M, hop, K, P, D = self.M, self.hop, self.N // 2 + 1, self.M * (self.M - 1) // 2, 37
        return {"stft": 4 * M * hop + 8 * M * K, "gcc_tau": 8 * M * K + 4 * P * D, "energy": 4 * P * D + 4 * D, "select_doa": 4 * D + 8,
                "ds_select": 8 * M * K + 8 * K, "istft": 8 * K + 4 * hop,
                # channel-form SRP on tcgen05 (srp_tc_small_kernel): reads the M spectra once, writes the D-cell energy map
                "srp": 8 * M * K + 4 * D}

    def pipeline_bytes_per_frame(self):
        return 4 * self.M * self.hop + 4 * (self.M * (self.M - 1) // 2) + 4 * self.hop   # 17 888 B

    def cpu_run_ref(self, orc, x64, n_threads):
        return self.cpu_run(orc, x64, n_threads, prefix="ref")     # mca::SourceSeparationAndLocalisation itself (oracle/_ref)

    def cpu_run(self, orc, x64, n_threads, prefix="orc"):
        # one stream per thread, exactly the reference object per stream
        res = [None] * len(x64)

        def work(i0, i1):
            for i in range(i0, i1):
                res[i] = orc.ssl_run(self.fs, self.xyz(), 1, x64[i], prefix=prefix)["doa_deg"]
        th = [threading.Thread(target=work, args=(len(x64) * i // n_threads, len(x64) * (i + 1) // n_threads)) for i in range(n_threads)]
        [t.start() for t in th]
        [t.join() for t in th]
        return res


class Cfg4(Workload):
    """64-mic planar array SRP-PHAT over a 3600-direction azimuth x elevation grid (BASELINE.json configs[3]); tensor-core contraction."""
    name = "cfg4: 64-mic 8x8 planar array 0.04 m pitch, SRP-PHAT over 120 az x 30 el = 3600 directions, 48 kHz, N=1024, hop=512"
    fs, N, hop, M, D = 48000, 1024, 512, 64, 3600
    B_default, T_default = 4, 256
    bound = {"srp": "tensor"}
    cpu_frames, cpu_streams = (4, 16), 1                            # ~0.95 GFLOP (x2 in float64 complex) per frame on the CPU

    def xyz(self):
        from mcarray_b200 import scenes
        return scenes.planar_array(8, 8, 0.04)

    def dirs(self):
        from mcarray_b200 import scenes
        az = np.linspace(-np.pi, np.pi, 120, endpoint=False); el = np.linspace(0.05, 1.45, 30)
        return scenes.az_el_dirs(az[:, None], el[None, :])

    def scene(self, stream_id, n):
        from mcarray_b200 import scenes
        d = self.dirs()
        src = (stream_id * 997 + 57 * 30 + 11) % len(d)
        return scenes.far_field_scene(self.xyz(), self.fs, n, d[src:src + 1], seed=scenes.stream_seed(stream_id))

    def make(self, mb, B, T):
        return mb.SrpPhat(self.fs, self.xyz(), self.N, self.dirs(), numOfSources=1, n_streams=B, max_frames_per_call=T)

    def result_bytes(self, p, B, T):
        return B * T * 4

    result_what = 4   # MCAG_OUT_CELL

    def kernel_bytes_per_frame(self):
        M, hop, K, D = self.M, self.hop, self.N // 2 + 1, self.D
        return {"stft": 4 * M * hop + 8 * M * K, "srp": 8 * M * K + 4 * D, "energy": 8 * D, "select_doa": 4 * D + 8}

    def kernel_flops_per_frame(self):
        return {"srp": 8 * self.D * self.M * (self.N // 2 + 1)}     # complex MAC = 8 real flops

    def pipeline_bytes_per_frame(self):
        return 4 * self.M * self.hop + 4 * self.D + 8               # 145 480 B

    def cpu_run(self, orc, x64, n_threads):
        mt = orc.mic_tau(self.xyz(), self.fs, self.dirs())
        out = []
        for x in x64:                                               # threads are inside orc.srp_channel (over directions)
            S = orc.stft(x, self.N, self.hop)
            out.append(np.argmax(orc.srp_channel(S, self.N, mt, n_threads=n_threads), axis=1))
        return out


class Cfg4Sharded(Cfg4):
    """cfg4 with the direction grid split over the GPUs: same input on every rank, D/G directions each, one NCCL MAX all-reduce of the
    packed per-frame (peak, cell) keys.  Strong scaling: the total work is fixed as N grows."""
    name = Cfg4.name + "; direction grid sharded over the GPUs, one NCCL max-allreduce per step"
    sharded_grid = True

    def make(self, mb, B, T):
        return mb.sharding.ShardedSrpPhat(self.fs, self.xyz(), self.N, self.dirs(), n_streams=B, max_frames_per_call=T)

    def host_input(self, rank, B, n, unique=8):
        return super().host_input(0, B, n, unique)                  # every rank sees the same array signals

    result_what = None


def _threaded(fn, items, n_threads):
    res = [None] * len(items)

    def work(i0, i1):
        for i in range(i0, i1):
            res[i] = fn(items[i])
    th = [threading.Thread(target=work, args=(len(items) * i // n_threads, len(items) * (i + 1) // n_threads)) for i in range(n_threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    return res


class Cfg1Mask(Workload):
    """BASELINE.json configs[0], masking half: FastBinauralMasking (45 mel bands, RELATIVE / BOTH) on 16 kHz stereo, 512-sample frames."""
    name = "cfg1m: 2-channel FastBinauralMasking (45 mel bands 500-5000 Hz, RELATIVE, BOTH), 0.086 m, 16 kHz, N=512, hop=256"
    fs, N, hop, M, d = 16000, 512, 256, 2, 0.086
    audio_channels = 2
    B_default, T_default = 2048, 125
    cpu_frames = (128, 512)

    def scene(self, stream_id, n):
        from mcarray_b200 import scenes
        az = [0.0, np.deg2rad(30 + 10 * (stream_id % 5))]
        return scenes.far_field_scene(scenes.linear_array([0, self.d]), self.fs, n, scenes.azimuth_dirs(az), seed=scenes.stream_seed(stream_id))

    def make(self, mb, B, T):
        return mb.FastBinauralMasking(self.fs, self.d, 500, 5000, "RELATIVE", "BOTH", n_streams=B, max_frames_per_call=T, frame_size=self.N)

    def result_bytes(self, p, B, T):
        return B * 2 * T * self.hop * 4

    result_what = None

    def kernel_bytes_per_frame(self):
        K, hop, nb = self.N // 2 + 1, self.hop, 45
        return {"stft": 4 * 2 * hop + 8 * 2 * K, "mask_stats": 8 * 2 * K + 4 * 6 * nb, "mask_scan": 4 * 6 * nb + 4 * 2 * nb, "mask_apply": 2 * 8 * 2 * K + 4 * 2 * nb,
                "istft": 8 * 2 * K + 4 * 2 * hop, "mask_fused": 4 * 2 * hop + 4 * 2 * hop + 8}   # fused: samples in, samples out, frame powers

    def pipeline_bytes_per_frame(self):
        return 4 * 2 * self.hop + 4 * 2 * self.hop                    # samples in, samples out

    def cpu_run(self, orc, x64, n_threads):
        return _threaded(lambda x: orc.mask_run(self.fs, self.d, 500, 5000, 1, 0, x)["out"], list(x64), n_threads)
    # (no cpu_run_ref: the reference's FastBinauralMasking picks N = 1024 at 16 kHz itself; this workload runs BASELINE's N = 512)


class Cfg1Loc(Workload):
    """BASELINE.json configs[0], localisation half: FreqGCCBinauralLocalisation (61-cell GCC-PHAT curve, 0.8 smoothing, arg-max)."""
    name = "cfg1l: 2-channel FreqGCCBinauralLocalisation (61 delays, 3 degree grid), 0.086 m, 16 kHz, N=512, hop=256"
    fs, N, hop, M, d = 16000, 512, 256, 2, 0.086
    B_default, T_default = 2048, 125
    cpu_frames = (128, 512)

    def scene(self, stream_id, n):
        from mcarray_b200 import scenes
        az = [np.deg2rad(-60 + 7 * (stream_id % 17))]
        return scenes.far_field_scene(scenes.linear_array([0, self.d]), self.fs, n, scenes.azimuth_dirs(az), seed=scenes.stream_seed(stream_id))

    def make(self, mb, B, T):
        return mb.FreqGCCBinauralLocalisation(self.fs, self.d, usePowerFloor=False, n_streams=B, max_frames_per_call=T, frame_size=self.N)

    def result_bytes(self, p, B, T):
        return B * T * 4

    result_what = 4   # MCAG_OUT_CELL

    def kernel_bytes_per_frame(self):
        K, hop, D = self.N // 2 + 1, self.hop, 61
        return {"stft": 4 * 2 * hop + 8 * 2 * K, "gcc_tau": 8 * 2 * K + 4 * D, "curve_scan": 2 * 4 * D + 4}

    def pipeline_bytes_per_frame(self):
        return 4 * 2 * self.hop + 4                                   # samples in, arg-max cell out

    def cpu_run(self, orc, x64, n_threads):
        return _threaded(lambda x: orc.freqgcc_run(self.fs, self.d, x)["idx"], list(x64), n_threads)


class Cfg1Multiband(Cfg1Loc):
    """MultibandBinarualLocalisation (15 linear sub-bands, 37-cell GCC-PHAT curves with 0.4 memory, energy histogram)."""
    name = "cfg1b: 2-channel MultibandBinarualLocalisation (15 linear bands, 37 delays, 5 degree grid), 0.086 m, 16 kHz, N=512, hop=256"

    def make(self, mb, B, T):
        return mb.MultibandBinarualLocalisation(self.fs, self.d, nbins=15, usePowerFloor=False, n_streams=B, max_frames_per_call=T, frame_size=self.N,
                                                noise_preestimated=True)

    def kernel_bytes_per_frame(self):
        K, hop, D, nb = self.N // 2 + 1, self.hop, 37, 15
        return {"stft": 4 * 2 * hop + 8 * 2 * K, "gcc_tau": 8 * 2 * K + 4 * nb * D + 4 * nb, "curve_scan": 2 * 4 * nb * D, "select_doa": 4 * nb * D + 4 * D + 4 * nb + 8}

    def cpu_run(self, orc, x64, n_threads):
        return _threaded(lambda x: orc.multiband_run(self.fs, self.d, x)["cell"], list(x64), n_threads)


class Cfg3(Workload):
    """32-mic linear array delay-and-sum beamformer steered to 181 azimuths, 2048-sample frames (BASELINE.json configs[2])."""
    name = "cfg3: 32-mic linear array 0.04 m pitch, delay-and-sum to 181 azimuths (spectra out), 48 kHz, N=2048, hop=1024"
    fs, N, hop, M, D = 48000, 2048, 1024, 32, 181
    B_default, T_default = 8, 64
    cpu_frames, cpu_streams = (8, 16), 16

    def xyz(self):
        from mcarray_b200 import scenes
        return scenes.linear_array((np.arange(self.M) - (self.M - 1) / 2) * 0.04)

    def doas(self):
        return np.deg2rad(np.arange(-90, 91, 1.0))

    def scene(self, stream_id, n):
        from mcarray_b200 import scenes
        az = np.deg2rad(-75 + 11 * (stream_id % 14))
        return scenes.far_field_scene(self.xyz(), self.fs, n, scenes.azimuth_dirs([az]), seed=scenes.stream_seed(stream_id))

    def make(self, mb, B, T):
        return mb.DelayAndSumFan(self.fs, self.xyz(), self.N, self.doas(), n_streams=B, max_frames_per_call=T)

    def result_bytes(self, p, B, T):
        return B * T * self.D * p.info.beams_pitch * 8

    result_what = 9   # MCAG_OUT_BEAMS, 760 MB per step

    def kernel_bytes_per_frame(self):
        M, hop, K, D = self.M, self.hop, self.N // 2 + 1, self.D
        return {"stft": 4 * M * hop + 8 * M * K, "ds_fan": 8 * M * K + 8 * D * K}

    def pipeline_bytes_per_frame(self):
        return 4 * self.M * self.hop + 8 * self.D * (self.N // 2 + 1)  # 1.62 MB / frame

    def cpu_run(self, orc, x64, n_threads):
        xs = self.xyz()[:, 0]
        return _threaded(lambda x: orc.ds_fan(orc.stft(x, self.N, self.hop), self.N, self.fs, xs, self.doas()).shape, list(x64), n_threads)


WORKLOADS = {"cfg1m": Cfg1Mask, "cfg1l": Cfg1Loc, "cfg1b": Cfg1Multiband, "cfg2": Cfg2, "cfg3": Cfg3, "cfg4": Cfg4, "cfg4s": Cfg4Sharded, "cfg5": Cfg5}


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every ~25 ms from a thread (the same counters
    the nvidia-smi recipe in B200_PROFILING.md prints; nvidia-smi's 100 ms loop is too coarse for a tens-of-ms region)."""
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}

    def __init__(self, gpu_index, period=0.010):
        # NVML queries take driver locks that kernel / NCCL launches also need: poll gently (10 ms), not in a tight loop
        self.gpu, self.rows, self._stop, self.thread, self.err, self.period = gpu_index, [], threading.Event(), None, None, period
        self._ready, self._armed = threading.Event(), threading.Event()

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)              # first query pays the lazy initialisation
            self._ready.set()
            self._armed.wait()                                          # samples are taken only inside the timed region
            while not self._stop.is_set():
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, nv.nvmlDeviceGetPowerUsage(h) / 1e3,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
                time.sleep(self.period)
            nv.nvmlShutdown()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self._ready.set()

    def prepare(self):
        """start the thread and wait until NVML is initialised (outside the timed region)"""
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        self._ready.wait(timeout=10)

    def start(self):
        if self.thread is None:
            self.prepare()
        self._armed.set()

    def stop(self):
        self._stop.set()
        self._armed.set()
        self.thread.join(timeout=5)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"no NVML samples ({self.err})"]}
        sm = [r[0] for r in self.rows]
        reasons = sorted({name for r in self.rows for name, bit in self.BITS.items() if r[3] & bit})
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": float(max(r[2] for r in self.rows)), "how": "NVML polled every ~10 ms during the timed region"}


def profile_read(p, reset=True):
    from mcarray_b200 import capi
    n = 32                                                         # >= MCAG_PROF_COUNT; names past the count are empty
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    capi.check(capi.lib().mcag_profile_read(p.handle, ms, cnt, C.c_int(int(reset))))
    capi.lib().mcag_profile_name.restype = C.c_char_p
    return {capi.lib().mcag_profile_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i] > 0}


def config_of(wl, B, T, world):
    """The `config` object of the JSON line: identical in the GPU arm and in the --impl reference arm (same workload, same batch)."""
    sharded = getattr(wl, "sharded_grid", False)
    n = wl.N + (T - 1) * wl.hop
    return {"workload": wl.name, "streams_per_gpu": B, "frames_per_stream_per_step": T, "samples_per_channel_per_step": n,
            "input_bytes_per_gpu": B * wl.M * n * 4, "l2_policy": "inputs larger than L2 (input + intermediates per step >> 126 MB)",
            "parallelism": (f"direction grid sharded over {world} GPU(s), one NCCL max-allreduce per step" if sharded else
                            f"independent array streams sharded over {world} GPU(s), no collective")}


FP32_LANES_PER_SM, SM_COUNT = 128, 148


def fp32_peak_tflops(sm_mhz):
    """CUDA-core FP32 peak of the B200: 148 SMs x 128 lanes x 2 flop (FMA) x SM clock (74.4 TFLOP/s at 1965 MHz)."""
    return SM_COUNT * FP32_LANES_PER_SM * 2 * (sm_mhz or 1965.0) * 1e6 / 1e12


def fft_flops(N):
    """real N-point FFT by the packed N/2-point complex transform: 2.5 N log2 N (SURVEY.md 8d uses the same count)"""
    return 2.5 * N * np.log2(N)


# ----------------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    """Pin this process (and the pinned host buffers it allocates afterwards) to the CPUs of the NUMA node the GPU hangs off."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": None, "note": "single NUMA node"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "note": repr(e)[:80]}


def cpu_arm(wl, B, T, cores, steps, warmup, prefer_ref=True):
    """The reference's CPU path on the host cores: oracle/_ref (the reference's own sources, kind "reference") when the workload is a
    reference class and the build is present, else the float64 restatement (kind "port").  Returns (seconds per step, kind, sample)."""
    import orc
    n = wl.N + (T - 1) * wl.hop
    use_ref = prefer_ref and orc.have_ref() and hasattr(wl, "cpu_run_ref")
    x = np.stack([wl.scene(10_000 + b % 4, n) for b in range(min(B, 4))])
    x64 = np.ascontiguousarray(np.concatenate([x] * ((B + len(x) - 1) // len(x)))[:B])
    run = (lambda xx: wl.cpu_run_ref(orc, xx, cores)) if use_ref else (lambda xx: wl.cpu_run(orc, xx, cores))
    for _ in range(max(warmup, 0)):
        run(x64[:cores])
    t0 = time.perf_counter()
    for _ in range(steps):
        run(x64)
    dt = (time.perf_counter() - t0) / steps
    units = B * wl.M * T * wl.hop
    kind = "reference" if use_ref else "port"
    what = "oracle/_ref (the reference's own .cpp files against the DSPONE/WIPP stand-in)" if use_ref else "float64 restatement (oracle/restated.hpp)"
    sample = f"{B} streams x {wl.M} ch x {T} frames ({units / 1e6:.1f} M channel-samples) per step, {what}, {cores} threads, one stream per thread"
    return dt, kind, sample, units


def run_reference_arm(wl, args):
    cores = os.cpu_count() or 1
    B = args.streams or wl.B_default
    T = args.frames or wl.T_default
    Bc, Tc = getattr(wl, "cpu_streams", B), T
    if getattr(wl, "cpu_frames", None) and not args.full_reference:
        Tc = min(T, wl.cpu_frames[0])          # heavy workloads (cfg3 / cfg4): bounded sample, stated in cpu_baseline.sample
    dt, kind, sample, units = cpu_arm(wl, Bc, Tc, cores, args.steps, min(args.warmup, 1))
    v = units / dt
    bounded = (Bc, Tc) != (B, T)
    return {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, B, T, args.gpus),
            "xrt_aggregate": (units / wl.M / wl.fs) / dt,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": sample + ("; bounded sample of the GPU arm's batch" if bounded else "; the GPU arm's full per-GPU batch")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ----------------------------------------------------------------------------------------------------------------------
class Env:
    pass


def measure(key, env, args, steps, warmup, top):
    """One workload on this rank's GPU (all ranks call it in lock step).  Returns the record on rank 0, None elsewhere."""
    torch, dist, mb, capi = env.torch, env.dist, env.mb, env.capi
    rank, world, local = env.rank, env.world, env.local
    wl = WORKLOADS[key]()
    B = (args.streams if top else 0) or wl.B_default
    T = (args.frames if top else 0) or wl.T_default
    n = wl.N + (T - 1) * wl.hop
    rows = B * wl.M
    units_rank = rows * T * wl.hop                                   # channel-samples consumed per step per rank
    dev = f"cuda:{local}"

    # ---- inputs: pinned host block (e2e arm) + a device-resident copy (value arm) -----------------------------------
    x = wl.host_input(rank, B, n)
    pin = torch.empty((rows, n), dtype=torch.float32).pin_memory()
    pin.numpy()[:] = x
    del x
    d_in = pin.to(dev, non_blocking=False)
    input_bytes = rows * n * 4

    mb.set_default_device(local)
    sharded = getattr(wl, "sharded_grid", False)
    sp = wl.make(mb, B, T) if sharded else None                   # grid-sharded processor: wraps a local handle + the all-reduce
    p = sp.local if sharded else wl.make(mb, B, T)
    stream = torch.cuda.ExternalStream(capi.lib().mcag_stream(p.handle), device=torch.device("cuda", local))
    d_out = None
    if p.info.n_out_channels:
        d_out = torch.empty((B * p.info.n_out_channels, T * wl.hop), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        p.flush_input()
        if sharded:
            sp.process_device(d_in, n, n)                           # local slice + packed arg-max + NCCL MAX all-reduce
        else:
            p.process_device(d_in, n, n, d_out, T * wl.hop if d_out is not None else 0)

    def reduce_ranks(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(ms):
        return reduce_ranks(ms, dist.ReduceOp.MAX) if world > 1 else ms

    # ---- value: input resident in HBM ---------------------------------------------------------------------------------
    for _ in range(warmup):
        step_device()
    p.synchronize()
    assert p.frames_done == T, (p.frames_done, T)
    capi.check(capi.lib().mcag_profile_enable(p.handle, 1)); profile_read(p)
    if sharded:
        sp.time_allreduce = True; sp.allreduce_ms()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.prepare()
    launches0 = p.kernel_launches
    barrier()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step_device()
    ev1.record(stream)
    p.synchronize()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    launches = p.kernel_launches - launches0
    prof = profile_read(p)
    capi.check(capi.lib().mcag_profile_enable(p.handle, 0))
    allreduce_ms = None
    if sharded:
        allreduce_ms = sp.allreduce_ms(); sp.time_allreduce = False
    ms_step = ms_total / steps
    units_job = units_rank if sharded else units_rank * world      # sharded grid: every rank works on the same samples
    value = units_job / (ms_step * 1e-3)

    # ---- sustained: the same step back to back for ~args.sustain seconds (the timed region above is tens of ms: clocks and power there
    #      are burst figures) ------------------------------------------------------------------------------------------------------
    sustained = None
    if top and args.sustain > 0:
        reps = max(steps, int(args.sustain * 1e3 / ms_step))
        s2 = ClockSampler(local, period=0.05) if rank == 0 else None
        if s2:
            s2.prepare()
        barrier()
        if s2:
            s2.start()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(reps):
            step_device()
        a1.record(stream)
        p.synchronize()
        barrier()
        ms_sus = max_over_ranks(a0.elapsed_time(a1)) / reps
        c2 = s2.stop() if s2 else None
        sustained = {"steps": reps, "seconds": ms_sus * reps * 1e-3, "ms_per_step": ms_sus, "value": units_job / (ms_sus * 1e-3), "clocks": c2}

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ------------------------------------------
    e2e = e2e_f32 = None
    res_bytes = int(wl.result_bytes(p, B, T))
    orows = B * p.info.n_out_channels
    if not args.no_e2e:
        # bare pinned-host -> device copy ceiling of this rank while every rank copies (what bounds the e2e leg from above)
        scratch = torch.empty_like(d_in)
        for _ in range(2):
            scratch.copy_(pin, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            scratch.copy_(pin, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = input_bytes * 3 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        h2d_min = reduce_ranks(h2d_gbs, dist.ReduceOp.MIN) if world > 1 else h2d_gbs
        del scratch

        def timed(step_fn):
            for _ in range(warmup):
                step_fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step_fn()
            p.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            barrier()
            return max_over_ranks(ms) / steps

        # (a) 16-bit PCM, the sample type of the reference's process(std::vector<int16_t*>&, ...) overload (test_mcarray.cpp:937,1023)
        if not sharded:
            pin16 = torch.empty((rows, n), dtype=torch.int16).pin_memory()
            pin16.copy_(pin.round().clamp_(-32768, 32767).to(torch.int16))
            out16 = torch.empty((orows, T * wl.hop), dtype=torch.int16).pin_memory() if orows else None
            nout16 = C.c_int(0)

            def step_s16():
                p.flush_input()
                capi.check(capi.lib().mcag_process_packed_s16(p.handle, C.c_void_p(pin16.data_ptr()), C.c_longlong(n), C.c_int(n),
                                                              C.c_void_p(out16.data_ptr()) if orows else None,
                                                              C.c_longlong(T * wl.hop if orows else 0), C.byref(nout16)))
                return wl.fetch_result(p)
            ms16 = timed(step_s16)
            audio = B * getattr(wl, "audio_channels", 0) * T * wl.hop           # synthesised samples that cross PCIe (the zero channels do not)
            d2h = res_bytes - audio * 4 + audio * 2
            e2e = {"value": units_job / (ms16 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": rows * n * 2, "d2h_bytes_per_step": int(d2h), "ms_per_step": ms16,
                   "sample_type": "int16 PCM (mcag_process_packed_s16 = the reference's process(int16_t*) overload), pinned host block, "
                                  "timed by the host clock around synchronous calls, max over ranks",
                   "h2d_ceiling_gbs": h2d_gbs, "h2d_ceiling_gbs_min_over_ranks": h2d_min,
                   "h2d_achieved_gbs": rows * n * 2 / (ms16 * 1e-3) / 1e9, "numa": env.numa}
            del pin16, out16
        # (b) fp32 host buffers
        if sharded or top or args.e2e_f32:
            out_host = torch.empty((orows, T * wl.hop), dtype=torch.float32).pin_memory() if orows else None
            nout = C.c_int(0)

            def step_f32():
                p.flush_input()
                if sharded:
                    capi.check(capi.lib().mcag_process_packed_f32(p.handle, C.c_void_p(pin.data_ptr()), C.c_longlong(n), C.c_int(n), None, C.c_longlong(0), C.byref(nout)))
                    return [t.cpu() for t in sp._reduce()]
                capi.check(capi.lib().mcag_process_packed_f32(p.handle, C.c_void_p(pin.data_ptr()), C.c_longlong(n), C.c_int(n),
                                                              C.c_void_p(out_host.data_ptr()) if orows else None,
                                                              C.c_longlong(T * wl.hop if orows else 0), C.byref(nout)))
                return wl.fetch_result(p)
            ms32 = timed(step_f32)
            e2e_f32 = {"value": units_job / (ms32 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": input_bytes, "d2h_bytes_per_step": res_bytes, "ms_per_step": ms32,
                       "sample_type": "fp32 (mcag_process_packed_f32)", "h2d_ceiling_gbs": h2d_gbs, "h2d_achieved_gbs": input_bytes / (ms32 * 1e-3) / 1e9}
            if e2e is None:
                e2e = e2e_f32; e2e_f32 = None
            del out_host

    p.close()
    del d_in, pin, d_out
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------------
    peaks = env.peaks
    kb = wl.kernel_bytes_per_frame()
    shares = {k: v[0] for k, v in prof.items()}
    tot = sum(shares.values()) or 1.0
    dom = max(shares, key=shares.get)
    dom_ms, dom_n = prof[dom]
    per_launch_ms = dom_ms / dom_n
    bytes_per_launch = kb.get(dom, wl.pipeline_bytes_per_frame()) * B * T
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    flops = getattr(wl, "kernel_flops_per_frame", lambda: {})().get(dom)
    if sharded and flops:
        flops = flops * (sp.d1 - sp.d0) / wl.D                     # this rank contracts only its slice of the direction grid
    bound = getattr(wl, "bound", {}).get(dom, "hbm") if isinstance(getattr(wl, "bound", None), dict) else "hbm"
    hbm = {"algorithmic_bytes_per_launch": bytes_per_launch, "achieved_gbs": achieved, "peak_gbs": peaks["hbm_gbs"], "frac": achieved / peaks["hbm_gbs"]}
    common = {"kernel": dom, "ms_per_launch": per_launch_ms, "kernel_share_of_step": dom_ms / tot, "traffic": None,
              "kernels_ms_per_step": {k: v[0] / steps for k, v in prof.items()}}
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32 = None
    if flops and bound != "tensor":
        tf = flops * B * T / (per_launch_ms * 1e-3) / 1e12
        pk = fp32_peak_tflops(sm_mhz)
        fp32 = {"achieved_tflops": tf, "peak": pk, "frac": tf / pk, "algorithmic_flops_per_launch": flops * B * T,
                "peak_source": f"148 SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)",
                "flop_count": getattr(wl, "flops_note", "")}
    if bound == "tensor" and flops:
        # 3xTF32: every algorithmic flop is issued three times on the TF32 pipe; TF32 dense peak = half the measured bf16 GEMM rate
        issued = 3.0 * flops * B * T / (per_launch_ms * 1e-3) / 1e12
        peak_tf32 = peaks["bf16_tflops"] / 2.0
        roofline = {"bound": "tensor", "achieved": issued, "peak": peak_tf32, "unit": "TFLOP/s", "frac": issued / peak_tf32,
                    "peak_source": peaks["source"] + ": cuBLAS bf16 burst / 2 = dense TF32 rate", "algorithmic_flops_per_launch": flops * B * T,
                    "algorithmic_tflops": flops * B * T / (per_launch_ms * 1e-3) / 1e12, "issue_factor": "3xTF32 (hi*hi + lo*hi + hi*lo)", "hbm": hbm}
    elif bound == "fp32" and fp32:
        # above the CUDA-core ridge (flop per compulsory byte >> FP32 peak / HBM peak = 11): the roofline that bounds it is the FP32 one
        roofline = {"bound": "fp32", "achieved": fp32["achieved_tflops"], "peak": fp32["peak"], "unit": "TFLOP/s", "frac": fp32["frac"],
                    "peak_source": fp32["peak_source"], "arithmetic_intensity_flop_per_byte": flops * B * T / bytes_per_launch,
                    "ridge_flop_per_byte": fp32["peak"] * 1e3 / peaks["hbm_gbs"], "hbm": hbm, "fp32": fp32}
    else:
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "peak_source": peaks["source"], "algorithmic_bytes_per_launch": bytes_per_launch}
        if fp32:
            roofline["fp32"] = fp32
    roofline.update(common)
    pb = wl.pipeline_bytes_per_frame()
    roofline["pipeline"] = {"algorithmic_bytes_per_frame": pb, "achieved_gbs": pb * B * T / (ms_step * 1e-3) / 1e9,
                            "hbm_frac": pb * B * T / (ms_step * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):   # measured DRAM bytes of one launch of the dominant kernel + what ncu says binds it (committed captures)
        with open(tr) as f:
            tj = json.load(f)
        roofline["traffic"] = tj.get(key, {}).get(dom)
        lim = tj.get("_limiter", {}).get(key, {}).get(dom)
        if lim:
            roofline["limiter"] = lim
        if roofline["traffic"] is not None and (B, T) != (wl.B_default, wl.T_default):
            roofline["traffic_note"] = "captured at the default streams / frames of this workload"

    rec = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_of(wl, B, T, world),
           "xrt_aggregate": (B * (1 if sharded else world) * T * wl.hop / wl.fs) / (ms_step * 1e-3), "xrt_per_stream": (T * wl.hop / wl.fs) / (ms_step * 1e-3),
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if sustained:
        rec["sustained"] = sustained
    if allreduce_ms is not None:
        rec["allreduce_ms_per_step"] = allreduce_ms
    if e2e:
        rec["e2e"] = e2e
    if e2e_f32:
        rec["e2e_f32"] = e2e_f32
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="array streams per GPU (default: per workload)")
    ap.add_argument("--frames", type=int, default=0, help="frames per stream per step (default: per workload)")
    ap.add_argument("--also", default=None, help="comma list of further workloads measured in short runs and reported under `workloads` "
                    "(default: cfg5,cfg4,cfg3,cfg1m,cfg1l + cfg4s when N > 1, for the default cfg2 run; 'none' to skip)")
    ap.add_argument("--also-steps", type=int, default=10)
    ap.add_argument("--sustain", type=float, default=1.5, help="seconds of back-to-back steps for the `sustained` record (0 = skip)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames per stream of the cpu_baseline sample (default 512)")
    ap.add_argument("--full-reference", action="store_true", help="--impl reference: run the full batch even for cfg3 / cfg4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-f32", action="store_true", help="also time fp32 host buffers for the `workloads` sub-records")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(run_reference_arm(WORKLOADS[args.workload](), args)), flush=True)
        return 0

    env = Env()
    env.rank, env.world, env.local = rank, world, local
    env.numa = bind_to_gpu_numa(local)                               # before torch allocates its pinned buffers
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (mcarray_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mcarray_b200 as mb
    from mcarray_b200 import capi
    env.torch, env.dist, env.mb, env.capi, env.peaks = torch, dist, mb, capi, measured_peaks()

    line = measure(args.workload, env, args, args.steps, args.warmup, True)
    also = args.also
    if also is None:
        also = ("cfg5,cfg4,cfg3,cfg1m,cfg1l" + (",cfg4s" if world > 1 else "")) if args.workload == "cfg2" else "none"
    subs = {}
    for key in [k for k in also.split(",") if k and k != "none"]:
        rec = measure(key, env, args, args.also_steps, 3, False)
        if rec is not None:
            for drop in ("metric", "unit", "higher_is_better", "vs_baseline", "data", "dtype"):
                rec.pop(drop, None)
            subs[key] = rec
    if world > 1 and "cfg4s" in subs or (world > 1 and "cfg4s" in also):
        # same-node single-GPU time of the unsharded grid: rank 0 alone, the other ranks wait at the barrier
        solo = None
        if rank == 0:
            e1 = Env(); e1.__dict__.update(env.__dict__); e1.world = 1
            solo = measure("cfg4", e1, argparse.Namespace(**{**vars(args), "no_e2e": True, "sustain": 0}), args.also_steps, 3, False)
        dist.barrier()
        if rank == 0 and "cfg4s" in subs and solo:
            subs["cfg4s"]["single_gpu_same_node"] = {"ms_per_step": solo["ms_per_step"], "value": solo["value"]}
            subs["cfg4s"]["speedup_over_single_gpu"] = solo["ms_per_step"] / subs["cfg4s"]["ms_per_step"]

    if rank == 0:
        if subs:
            line["workloads"] = subs
        if world == 1 and not args.no_cpu_baseline:
            wl = WORKLOADS[args.workload]()
            cores = os.cpu_count() or 1
            Bc = getattr(wl, "cpu_streams", max(cores, 1) * 4)
            Tc = args.cpu_frames or getattr(wl, "cpu_frames", (128, 512))[1]
            dt, kind, sample, units = cpu_arm(wl, Bc, Tc, cores, 1, 1)
            line["cpu_baseline"] = {"value": units / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": dt}
            if "cfg5" in subs:   # the literal mcbeam processor: the reference's own code can run it
                w5 = WORKLOADS["cfg5"]()
                dt, kind, sample, units = cpu_arm(w5, cores, 32, cores, 1, 0)
                subs["cfg5"]["cpu_baseline"] = {"value": units / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": dt}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

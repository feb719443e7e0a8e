#!/usr/bin/env python
"""bench.py — channel-samples/s and xRT of the mcarray hot path on B200 (driver contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic multichannel audio:
  cfg2 (default, the configuration the BASELINE.json metric is quoted on that fits one GPU):
       B array streams x 8-mic circular array, 48 kHz, N = 1024, hop = 512: STFT -> GCC-PHAT on all 28 pairs ->
       integer-lag TDOA, T frames per stream per step.
  cfg5: B streams x 16-mic linear array, 16 kHz, N = 512: SourceSeparationAndLocalisation (STFT -> GCC-PHAT tau grid ->
       SRP energy -> selectDOA -> delay-and-sum -> overlap-add), the literal mcbeam processor batched over streams.
Streams are independent, so N > 1 shards them across ranks with no collective (weak scaling: B streams PER GPU).

  value  whole-job channel-samples/s with the input already resident in HBM (mcag_process_device_f32)
  e2e    the same metric through the host-buffer C-ABI call (mcag_process_packed_f32 from pinned memory + result fetch),
         host<->device copies inside the timed region
  roofline      dominant kernel (largest share of the per-kernel CUDA-event times recorded on the handle's stream)
  cpu_baseline  the float64 oracle (oracle/, checker + CPU baseline only) on a bounded sample, all host cores

--impl reference times the reference's CPU path (the oracle restatement; the reference's own DSPONE/WIPP/FFTW build is
not available) on the same config.
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
    dominant_hint = "tdoa"

    def scene(self, stream_id, n):
        from mcarray_b200 import scenes
        xyz = scenes.circular_array(self.M, 0.10)
        az = -np.pi + 2 * np.pi * ((stream_id * 0.6180339887) % 1.0)
        return scenes.far_field_scene(xyz, self.fs, n, scenes.azimuth_dirs([az]), seed=scenes.stream_seed(stream_id))

    def make(self, mb, B, T):
        return mb.TdoaEstimator(self.fs, self.M, self.N, self.max_lag, n_streams=B, max_frames_per_call=T)

    def result_bytes(self, p, B, T):
        return B * T * p.info.n_pairs * 4

    def fetch_result(self, p):
        return p.lags()

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

    def fetch_result(self, p):
        return p.cells()

    def kernel_bytes_per_frame(self):
        M, hop, K, P, D = self.M, self.hop, self.N // 2 + 1, self.M * (self.M - 1) // 2, 37
        return {"stft": 4 * M * hop + 8 * M * K, "gcc_tau": 8 * M * K + 4 * P * D, "energy": 4 * P * D + 4 * D, "select_doa": 4 * D + 8,
                "ds_select": 8 * M * K + 8 * K, "istft": 8 * K + 4 * hop,
                # channel-form SRP on tcgen05 (srp_tc_small_kernel): reads the M spectra once, writes the D-cell energy map
                "srp": 8 * M * K + 4 * D}

    def pipeline_bytes_per_frame(self):
        return 4 * self.M * self.hop + 4 * (self.M * (self.M - 1) // 2) + 4 * self.hop   # 17 888 B

    def cpu_run(self, orc, x64, n_threads):
        # one stream per thread, exactly the reference object per stream
        res = [None] * len(x64)

        def work(i0, i1):
            for i in range(i0, i1):
                res[i] = orc.ssl_run(self.fs, self.xyz(), 1, x64[i])["doa_deg"]
        th = [threading.Thread(target=work, args=(len(x64) * i // n_threads, len(x64) * (i + 1) // n_threads)) for i in range(n_threads)]
        [t.start() for t in th]
        [t.join() for t in th]
        return res


class Cfg4(Workload):
    """64-mic planar array SRP-PHAT over a 3600-direction azimuth x elevation grid (BASELINE.json configs[3]); tensor-core contraction."""
    name = "cfg4: 64-mic 8x8 planar array 0.04 m pitch, SRP-PHAT over 120 az x 30 el = 3600 directions, 48 kHz, N=1024, hop=512"
    fs, N, hop, M, D = 48000, 1024, 512, 64, 3600
    B_default, T_default = 4, 256
    bound = "tensor"
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

    def fetch_result(self, p):
        return p.cells()

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

    def fetch_result(self, p):
        return None


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

    def fetch_result(self, p):
        return None

    def kernel_bytes_per_frame(self):
        K, hop, nb = self.N // 2 + 1, self.hop, 45
        return {"stft": 4 * 2 * hop + 8 * 2 * K, "mask_stats": 8 * 2 * K + 4 * 6 * nb, "mask_scan": 4 * 6 * nb + 4 * 2 * nb, "mask_apply": 2 * 8 * 2 * K + 4 * 2 * nb,
                "istft": 8 * 2 * K + 4 * 2 * hop}

    def pipeline_bytes_per_frame(self):
        return 4 * 2 * self.hop + 4 * 2 * self.hop                    # samples in, samples out

    def cpu_run(self, orc, x64, n_threads):
        return _threaded(lambda x: orc.mask_run(self.fs, self.d, 500, 5000, 1, 0, x)["out"], list(x64), n_threads)


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

    def fetch_result(self, p):
        return p.cells()

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
        return B * T * self.D * p.info.spectrum_pitch * 8

    def fetch_result(self, p):
        # MCAG_OUT_BEAMS (760 MB per step) into a pinned buffer kept across steps: a fresh pageable array per step cost 150 ms
        import torch
        from mcarray_b200 import capi
        nbytes = self.result_bytes(p, p.info.n_streams, p.frames_done)
        if getattr(self, "_pin", None) is None or self._pin.numel() < nbytes:
            self._pin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        capi.check(capi.lib().mcag_fetch(p.handle, 9, C.c_void_p(self._pin.data_ptr()), C.c_longlong(nbytes)))
        return self._pin

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
    n = 14
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    capi.check(capi.lib().mcag_profile_read(p.handle, ms, cnt, C.c_int(int(reset))))
    capi.lib().mcag_profile_name.restype = C.c_char_p
    return {capi.lib().mcag_profile_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i] > 0}


# ----------------------------------------------------------------------------------------------------------------------
def run_cpu(wl, args, rank, world, as_reference_arm):
    """The reference's CPU path (float64 oracle restatement) on a bounded sample of the workload, all host cores."""
    import orc
    cores = os.cpu_count() or 1
    if not args.cpu_frames:
        args.cpu_frames = getattr(wl, "cpu_frames", (128, 512))[0 if as_reference_arm else 1]
    n = wl.N + (args.cpu_frames - 1) * wl.hop
    Bc = getattr(wl, "cpu_streams", max(cores, 1) * args.cpu_streams_per_core)
    x = np.stack([wl.scene(10_000 + b % 4, n) for b in range(min(Bc, 4))])
    x64 = np.ascontiguousarray(np.concatenate([x] * ((Bc + len(x) - 1) // len(x)))[:Bc])
    units = Bc * wl.M * args.cpu_frames * wl.hop
    sample = f"{Bc} streams x {wl.M} ch x {args.cpu_frames} frames ({units / 1e6:.1f} M channel-samples) per step, float64, {cores} threads, one stream per thread"
    if not as_reference_arm:
        wl.cpu_run(orc, x64[:cores], cores)                     # warm-up (page-in, thread start)
        t0 = time.perf_counter(); wl.cpu_run(orc, x64, cores); dt = time.perf_counter() - t0
        return {"value": units / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "seconds": dt}
    for _ in range(args.warmup):
        wl.cpu_run(orc, x64[:cores], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wl.cpu_run(orc, x64, cores)
    dt = time.perf_counter() - t0
    v = units * args.steps / dt
    return {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "streams_per_step": Bc, "frames_per_stream": args.cpu_frames, "note": "bounded sample of the GPU arm's workload"},
            "xrt_aggregate": (units / wl.M / wl.fs) * args.steps / dt,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="array streams per GPU (default: per workload)")
    ap.add_argument("--frames", type=int, default=0, help="frames per stream per step (default: per workload)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames per stream of the CPU sample (default 512 for cpu_baseline, 128 per step for --impl reference)")
    ap.add_argument("--cpu-streams-per-core", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-s16", action="store_true", default=None,
                    help="also time the end-to-end path with int16 PCM host buffers (extra key e2e_s16); default: on at N = 1 for cfg2")
    ap.add_argument("--no-e2e-s16", dest="e2e_s16", action="store_false")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]()

    if args.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(run_cpu(wl, args, rank, world, True)), flush=True)
        return 0

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

    B = args.streams or wl.B_default
    T = args.frames or wl.T_default
    n = wl.N + (T - 1) * wl.hop
    rows = B * wl.M
    units_rank = rows * T * wl.hop                                   # channel-samples consumed per step per rank

    # ---- inputs: pinned host block (e2e arm) + a device-resident copy (value arm) -----------------------------------
    x = wl.host_input(rank, B, n)
    pin = torch.empty((rows, n), dtype=torch.float32).pin_memory()
    pin.numpy()[:] = x
    del x
    d_in = pin.to(f"cuda:{local}", non_blocking=False)
    input_bytes = rows * n * 4

    mb.set_default_device(local)
    sharded = getattr(wl, "sharded_grid", False)
    sp = wl.make(mb, B, T) if sharded else None                   # grid-sharded processor: wraps a local handle + the all-reduce
    p = sp.local if sharded else wl.make(mb, B, T)
    stream = torch.cuda.ExternalStream(capi.lib().mcag_stream(p.handle), device=torch.device("cuda", local))
    d_out = None
    if p.info.n_out_channels:
        d_out = torch.empty((B * p.info.n_out_channels, T * wl.hop), dtype=torch.float32, device=f"cuda:{local}")

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

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: input resident in HBM ---------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    p.synchronize()
    assert p.frames_done == T, (p.frames_done, T)
    capi.check(capi.lib().mcag_profile_enable(p.handle, 1)); profile_read(p)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.prepare()
    launches0 = p.kernel_launches
    barrier()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    p.synchronize()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    launches = p.kernel_launches - launches0
    prof = profile_read(p)
    capi.check(capi.lib().mcag_profile_enable(p.handle, 0))
    ms_step = ms_total / args.steps
    units_job = units_rank if sharded else units_rank * world      # sharded grid: every rank works on the same samples
    value = units_job / (ms_step * 1e-3)

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ------------------------------------------
    e2e = None
    if not args.no_e2e:
        out_host = None
        if p.info.n_out_channels:
            out_host = torch.empty((B * p.info.n_out_channels, T * wl.hop), dtype=torch.float32).pin_memory()
        res_bytes = wl.result_bytes(p, B, T)
        nout = C.c_int(0)

        def step_e2e():
            p.flush_input()
            if sharded:
                p.flush_input()
                capi.check(capi.lib().mcag_process_packed_f32(p.handle, C.c_void_p(pin.data_ptr()), C.c_longlong(n), C.c_int(n), None, C.c_longlong(0), C.byref(nout)))
                return [t.cpu() for t in sp._reduce()]
            capi.check(capi.lib().mcag_process_packed_f32(p.handle, C.c_void_p(pin.data_ptr()), C.c_longlong(n), C.c_int(n),
                                                          C.c_void_p(out_host.data_ptr()) if out_host is not None else None,
                                                          C.c_longlong(T * wl.hop if out_host is not None else 0), C.byref(nout)))
            return wl.fetch_result(p)
        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        e1.record(stream)
        p.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms)) / args.steps
        e2e = {"value": units_job / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": input_bytes, "d2h_bytes_per_step": int(res_bytes),
               "ms_per_step": ms_e2e}

    # ---- e2e with 16-bit PCM host buffers (the int16 overload of process(), test_mcarray.cpp:937): half the PCIe bytes ----------------
    e2e_s16 = None
    if args.e2e_s16 is None:
        args.e2e_s16 = world == 1 and args.workload == "cfg2"
    if not args.no_e2e and not sharded and args.e2e_s16:
        pin16 = torch.empty((rows, n), dtype=torch.int16).pin_memory()
        pin16.copy_(pin.round().clamp_(-32768, 32767).to(torch.int16))
        in_ptrs = (C.POINTER(C.c_int16) * rows)(*[C.cast(pin16.data_ptr() + 2 * r * n, C.POINTER(C.c_int16)) for r in range(rows)])
        out16, out_ptrs, orows = None, None, B * p.info.n_out_channels
        if orows:
            out16 = torch.empty((orows, T * wl.hop), dtype=torch.int16).pin_memory()
            out_ptrs = (C.POINTER(C.c_int16) * orows)(*[C.cast(out16.data_ptr() + 2 * r * T * wl.hop, C.POINTER(C.c_int16)) for r in range(orows)])
        nout16 = C.c_int(0)

        def step_s16():
            p.flush_input()
            capi.check(capi.lib().mcag_process_s16(p.handle, in_ptrs, C.c_int(n), out_ptrs, C.c_int(T * wl.hop if orows else 0), C.byref(nout16)))
            return wl.fetch_result(p)
        for _ in range(args.warmup):
            step_s16()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_s16()
        p.synchronize()
        ms16 = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        barrier()
        e2e_s16 = {"value": units_job / (ms16 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": rows * n * 2, "d2h_bytes_per_step": int(wl.result_bytes(p, B, T)),
                   "ms_per_step": ms16, "note": "same samples rounded to int16 PCM through mcag_process_s16 (planar pinned host rows); timed by host clock around synchronous calls"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------------
    peaks = measured_peaks()
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
    if getattr(wl, "bound", "hbm") == "tensor" and flops:
        # 3xTF32: every algorithmic flop is issued three times on the TF32 pipe; TF32 dense peak = half the measured bf16 GEMM rate
        issued = 3.0 * flops * B * T / (per_launch_ms * 1e-3) / 1e12
        peak_tf32 = peaks["bf16_tflops"] / 2.0
        roofline_t = {"bound": "tensor", "kernel": dom, "achieved": issued, "peak": peak_tf32, "unit": "TFLOP/s", "frac": issued / peak_tf32, "traffic": None,
                      "peak_source": peaks["source"] + ": cuBLAS bf16 burst / 2 = dense TF32 rate", "algorithmic_flops_per_launch": flops * B * T,
                      "algorithmic_tflops": flops * B * T / (per_launch_ms * 1e-3) / 1e12, "issue_factor": "3xTF32 (hi*hi + lo*hi + hi*lo)",
                      "ms_per_launch": per_launch_ms, "kernel_share_of_step": dom_ms / tot,
                      "hbm": {"algorithmic_bytes_per_launch": bytes_per_launch, "achieved_gbs": achieved, "frac": achieved / peaks["hbm_gbs"]},
                      "kernels_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()}}
    else:
        roofline_t = None
    roofline = roofline_t or {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": None, "peak_source": peaks["source"], "algorithmic_bytes_per_launch": bytes_per_launch, "ms_per_launch": per_launch_ms,
                "kernel_share_of_step": dom_ms / tot,
                "pipeline": {"algorithmic_bytes_per_frame": wl.pipeline_bytes_per_frame(),
                             "achieved": wl.pipeline_bytes_per_frame() * B * T / (ms_step * 1e-3) / 1e9,
                             "frac": wl.pipeline_bytes_per_frame() * B * T / (ms_step * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                "kernels_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):   # measured DRAM bytes of one launch of the dominant kernel + what ncu says binds it (committed captures)
        with open(tr) as f:
            tj = json.load(f)
        roofline["traffic"] = tj.get(args.workload, {}).get(dom)
        lim = tj.get("_limiter", {}).get(args.workload, {}).get(dom)
        if lim:
            roofline["limiter"] = lim
        if roofline["traffic"] is not None and (B, T) != (wl.B_default, wl.T_default):
            roofline["traffic_note"] = "captured at the default streams / frames of this workload"

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "streams_per_gpu": B, "frames_per_stream_per_step": T, "samples_per_channel_per_step": n,
                       "input_bytes_per_gpu": input_bytes, "l2_policy": "inputs larger than L2 (input + intermediates per step >> 126 MB)",
                       "parallelism": (f"direction grid sharded over {world} GPU(s), one NCCL max-allreduce per step" if sharded else
                                       f"independent array streams sharded over {world} GPU(s), no collective")},
            "xrt_aggregate": (B * (1 if sharded else world) * T * wl.hop / wl.fs) / (ms_step * 1e-3), "xrt_per_stream": (T * wl.hop / wl.fs) / (ms_step * 1e-3),
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if e2e:
        line["e2e"] = e2e
    if e2e_s16:
        line["e2e_s16"] = e2e_s16
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = run_cpu(wl, args, rank, world, False)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Python mirror of the reference's processor interface over the C ABI (include/mcarray_b200.h).

The shipping host layer is C++ (include/mcarray/*.h); this module exists so the parity tests and bench.py can drive
the same C entry points with numpy / torch buffers.  Class names, constructor arguments and the
process()/getMaxLatency()/setCallback() vocabulary follow the reference classes:
  SourceSeparationAndLocalisation  include/mcarray/SourceSeparationAndLocalisation.h:42-56
  SourceLocalisation               include/mcarray/SourceLocalisation.h:38-52
  FreqGCCBinauralLocalisation      include/mcarray/BinauralLocalisation.h:188-192
  FastBinauralMasking              include/mcarray/FastBinauralMasking.h:71-104
"""
import ctypes as C

import numpy as np

from . import _capi as capi

MaskingMethod = dict(FACTOR=0, RELATIVE=1, FULL=3, NOISY=4, NOTHING=5)   # ArrayModules.h:81
MaskingAlg = dict(BOTH=0, SPATIAL=1, TEMPORAL=2)                          # ArrayModules.h:89

_OUT_DTYPE = {capi.OUT_SPECTRA: np.complex64, capi.OUT_POWER_DB: np.float32, capi.OUT_CORR: np.float32, capi.OUT_ENERGY: np.float32,
              capi.OUT_CELL: np.int32, capi.OUT_PROB: np.float32, capi.OUT_LAGS: np.int32, capi.OUT_CURVES: np.float32,
              capi.OUT_ACTIVE: np.uint8, capi.OUT_BEAMS: np.complex64, capi.OUT_MASK_Q: np.float32, capi.OUT_MASK_DEC: np.uint8,
              capi.OUT_BAND_CELL: np.int32, capi.OUT_TRACK_DOA: np.float64}


_default_device = 0


def set_default_device(ordinal):
    """CUDA device ordinal new processors are created on (one process per GPU: call it with LOCAL_RANK)."""
    global _default_device
    _default_device = int(ordinal)


class Processor:
    """One C-ABI handle.  Keeps the numpy tables the config points at alive."""

    def __init__(self, **kw):
        kw.setdefault("device", _default_device)
        self._keep = []
        cfg = capi.Config()
        capi.lib().mcag_config_init(C.byref(cfg))
        for k, v in kw.items():
            if isinstance(v, np.ndarray):
                v = np.ascontiguousarray(v, dtype=np.float64)
                self._keep.append(v)
                setattr(cfg, k, capi.dp(v))
            elif v is not None:
                setattr(cfg, k, v)
        self.cfg = cfg
        self.handle = C.c_void_p()
        capi.check(capi.lib().mcag_create(C.byref(cfg), C.byref(self.handle)))
        self.info = capi.Info()
        capi.check(capi.lib().mcag_get_info(self.handle, C.byref(self.info)))

    def close(self):
        if self.handle:
            capi.lib().mcag_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- DSPONE-style getters -------------------------------------------------------------------------
    def getFrameSize(self): return self.info.frame_size
    def getWindowSize(self): return self.info.window_size
    def getAnalysisLength(self): return self.info.analysis_length
    def getOneSidedFFTLength(self): return self.info.one_sided_length
    def getNumberOfChannels(self): return self.info.n_channels
    def getMaxLatency(self): return self.info.max_latency

    def reset(self):
        capi.check(capi.lib().mcag_reset(self.handle))

    def flush_input(self):
        capi.check(capi.lib().mcag_flush_input(self.handle))

    @property
    def frames_done(self):
        return capi.lib().mcag_frames_done(self.handle)

    @property
    def kernel_launches(self):
        return capi.lib().mcag_kernel_launches(self.handle)

    def max_samples_per_call(self):
        return self.info.max_frames_per_call * self.info.hop

    # -- process(): host buffers, any chunking, like dsp::ShortTimeProcess::process ----------------------
    def process(self, x):
        """x: [B*M][n] (or [M][n] when B = 1) float32 / float64 / int16 host array.  Returns the synthesised audio
        [B*C][n_out] in the input dtype (empty for analysis-only processors)."""
        x = np.ascontiguousarray(x)
        rows, n = x.shape
        assert rows == self.info.n_streams * self.info.n_channels
        fn, ct = {np.dtype(np.float32): ("mcag_process_f32", C.c_float), np.dtype(np.float64): ("mcag_process_f64", C.c_double),
                  np.dtype(np.int16): ("mcag_process_s16", C.c_int16)}[x.dtype]
        inp = (C.POINTER(ct) * rows)(*[x[r].ctypes.data_as(C.POINTER(ct)) for r in range(rows)])
        orow = self.info.n_streams * self.info.n_out_channels
        cap = n + self.info.max_latency
        out = np.zeros((max(orow, 1), cap), dtype=x.dtype)
        outp = (C.POINTER(ct) * max(orow, 1))(*[out[r].ctypes.data_as(C.POINTER(ct)) for r in range(max(orow, 1))])
        nout = C.c_int(0)
        capi.check(getattr(capi.lib(), fn)(self.handle, inp, C.c_int(n), outp if orow else None, C.c_int(cap), C.byref(nout)))
        return out[:orow, :nout.value]

    def process_device(self, d_in, in_pitch, nsamples, d_out=None, out_pitch=0):
        """device-resident variant: d_in / d_out are torch CUDA tensors (or raw pointers); asynchronous."""
        nout = C.c_int(0)
        capi.check(capi.lib().mcag_process_device_f32(self.handle, capi.vp(d_in), C.c_longlong(in_pitch), C.c_int(nsamples), capi.vp(d_out),
                                                      C.c_longlong(out_pitch), C.byref(nout)))
        return nout.value

    def synchronize(self):
        capi.check(capi.lib().mcag_synchronize(self.handle))

    def fetch(self, what, shape, out=None):
        """result `what` of the last call as a numpy array; `out`: a caller-owned C-contiguous array of that shape and dtype to fill instead
        of a fresh one (page-locked memory makes the copy run at the PCIe rate: bench.py reads 5.4 MB of lags in 0.1 ms instead of 0.5)"""
        dt = np.dtype(_OUT_DTYPE[what])
        if out is None:
            out = np.empty(shape, dtype=dt)
        elif out.dtype != dt or tuple(out.shape) != tuple(shape) or not out.flags.c_contiguous:
            raise ValueError(f"fetch: out must be a C-contiguous {dt} array of shape {tuple(shape)}")
        capi.check(capi.lib().mcag_fetch(self.handle, C.c_int(what), out.ctypes.data_as(C.c_void_p), C.c_longlong(out.nbytes)))
        return out

    # shaped accessors for the last call
    def _bt(self):
        return self.info.n_streams, self.frames_done

    def spectra(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_SPECTRA, (B, T, self.info.n_channels, self.info.spectrum_pitch))[..., : self.info.one_sided_length]

    def power_db(self):
        return self.fetch(capi.OUT_POWER_DB, self._bt())

    def active(self):
        return self.fetch(capi.OUT_ACTIVE, self._bt())


class _Localiser(Processor):
    def __init__(self, kind, sampleRate, mic_xyz, numOfSources, usePowerFloor, n_streams, max_frames_per_call, frame_size=None, emit=0,
                 srp_form=0):
        mic_xyz = np.ascontiguousarray(mic_xyz, dtype=np.float64).reshape(-1, 3)
        self.doa_step = np.float32(5 * np.pi / 180)                         # SteeringBeamforming.cpp:39
        N = frame_size or capi.frame_size(sampleRate, 0.025)                # _frameRate, SourceSeparationAndLocalisation.h:60
        tau = capi.pair_tau_reference(mic_xyz, sampleRate, self.doa_step)
        turns = capi.steer_turns_reference(mic_xyz, sampleRate, N, self.doa_step)
        super().__init__(kind=kind, sample_rate=sampleRate, frame_size=N, hop=N // 2, n_channels=len(mic_xyz), n_streams=n_streams,
                         max_frames_per_call=max_frames_per_call, n_dirs=tau.shape[1], pair_tau=tau, steer_turns=turns,
                         n_sources=numOfSources, use_power_floor=int(usePowerFloor), noise_margin_db=3.0, emit=emit, srp_form=srp_form)
        self._callback = None

    def setCallback(self, cb):
        """cb(doa_deg [S], prob [S], power, numOfSources) — LocalisationCallback::setDOA, SoundLocalisationCallback.h:53"""
        self._callback = cb

    def cells(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_CELL, (B, T, self.info.n_sources))

    def prob(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_PROB, (B, T, self.info.n_sources))

    def energy(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_ENERGY, (B, T, self.info.n_dirs))

    def corr(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_CORR, (B, T, self.info.n_pairs, self.info.n_dirs))

    def doa_deg(self, cells=None):
        cells = self.cells() if cells is None else cells
        ang = np.array([capi.cell_angle(i, self.doa_step) for i in range(self.info.n_dirs)] + [0.0])
        return ang[cells] * (180 / np.pi)                                   # toDegrees, microhponeArrayHelpers.cpp:91-98

    def process(self, x):
        y = super().process(x)
        if self._callback is not None and self.frames_done:
            doa, prob, power, act = self.doa_deg(), self.prob(), self.power_db(), self.active()
            for b in range(self.info.n_streams):                            # one stream after the other, frames in order
                for t in range(self.frames_done):
                    if act[b, t]:
                        self._callback(doa[b, t], prob[b, t], float(power[b, t]), self.info.n_sources)
        return y


class SourceSeparationAndLocalisation(_Localiser):
    def __init__(self, sampleRate, microphonePositions, numOfSources, usePowerFloor=True, n_streams=1, max_frames_per_call=256, **kw):
        super().__init__(capi.KIND_SSL, sampleRate, microphonePositions, numOfSources, usePowerFloor, n_streams, max_frames_per_call, **kw)


class SourceLocalisation(_Localiser):
    def __init__(self, sampleRate, microphonePositions, numOfSources, usePowerFloor=True, n_streams=1, max_frames_per_call=256, **kw):
        super().__init__(capi.KIND_SL, sampleRate, microphonePositions, numOfSources, usePowerFloor, n_streams, max_frames_per_call, **kw)


class FreqGCCBinauralLocalisation(Processor):
    def __init__(self, sampleRate, microphoneDistance, usePowerFloor=True, n_streams=1, max_frames_per_call=256, frame_size=None,
                 noise_preestimated=True, deterministic_tracker=False):
        self.doa_step = np.float32(3 * np.pi / 180)                         # BinauralLocalisation.cpp:328
        N = frame_size or capi.frame_size(sampleRate, 0.075)                # BinauralLocalisation.h:196
        xyz = np.array([[0.0, 0, 0], [microphoneDistance, 0, 0]])
        tau = capi.pair_tau_reference(xyz, sampleRate, self.doa_step)       # same helper chain as :363-366
        super().__init__(kind=capi.KIND_FREQGCC, sample_rate=sampleRate, frame_size=N, hop=N // 2, n_channels=2, n_streams=n_streams,
                         max_frames_per_call=max_frames_per_call, n_dirs=tau.shape[1], pair_tau=tau, use_power_floor=int(usePowerFloor),
                         noise_margin_db=6.0, floor_ccs_power=1, noise_preestimated=int(noise_preestimated), corr_memory=0.8,
                         doa_tracker=int(deterministic_tracker), doa_memory=0.6)   # _maxDoaMemoryFactor, BinauralLocalisation.h:199
        self._tracker = bool(deterministic_tracker)
        self._callback = None

    def setCallback(self, cb):
        """cb(doa_deg [1], prob [1], power, 1) — BinauralLocalisation.cpp:521"""
        self._callback = cb

    def curves(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_CURVES, (B, T, self.info.n_dirs))

    def cells(self):
        return self.fetch(capi.OUT_CELL, self._bt())

    def tracked_doa(self):
        """deterministic_tracker: _currentDOA (rad) after every frame, the `#else` branch of BinauralLocalisation.cpp:501-504"""
        return self.fetch(capi.OUT_TRACK_DOA, self._bt())

    def prob(self):
        """deterministic_tracker: setProbability (BinauralLocalisation.cpp:454,569-631) after every frame"""
        return self.fetch(capi.OUT_PROB, self._bt())

    def process(self, x):
        y = super().process(x)
        if self._callback is not None and self.frames_done:
            act, power = self.active(), self.power_db()
            if self._tracker:
                doa, prob = np.degrees(self.tracked_doa()), self.prob()
            else:
                cells = self.cells()
                ang = np.array([capi.cell_angle(i, self.doa_step) for i in range(self.info.n_dirs)])
                doa, prob = np.degrees(ang[cells]), np.ones(cells.shape)
            for b in range(self.info.n_streams):
                for t in range(self.frames_done):
                    if act[b, t]:
                        self._callback(doa[b, t:t + 1], prob[b, t:t + 1], float(power[b, t]), 1)
        return y


class MultibandBinarualLocalisation(Processor):
    """MultibandBinarualLocalisation(sampleRate, microphonePositions, nbins=15, usePowerFloor) — include/mcarray/MultibandBinarualLocalisation.h:40;
    here the two-microphone array is given by its spacing."""

    def __init__(self, sampleRate, microphoneDistance, nbins=15, usePowerFloor=True, n_streams=1, max_frames_per_call=256, frame_size=None,
                 noise_preestimated=False):
        self.doa_step = np.float32(5 * np.pi / 180)                         # MultibandBinarualLocalisation.cpp:62
        N = frame_size or capi.frame_size(sampleRate, 0.025)                # _frameRate, .h:41
        tau, H = capi.multiband_setup(sampleRate, microphoneDistance, N, nbins)
        super().__init__(kind=capi.KIND_MULTIBAND, sample_rate=sampleRate, frame_size=N, hop=N // 2, n_channels=2, n_streams=n_streams,
                         max_frames_per_call=max_frames_per_call, n_dirs=len(tau), pair_tau=tau, n_bands=nbins, band_coefs=H,
                         use_power_floor=int(usePowerFloor), noise_margin_db=3.0, noise_preestimated=int(noise_preestimated),
                         corr_memory=float(np.float32(0.4)))               # _corrMemoryFactor, .h:44
        self.nbins, self.H, self.tau = nbins, H, tau
        self._callback = None

    def setCallback(self, cb):
        self._callback = cb

    def cells(self):
        return self.fetch(capi.OUT_CELL, self._bt())

    def prob(self):
        return self.fetch(capi.OUT_PROB, self._bt())

    def histogram(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_ENERGY, (B, T, self.info.n_dirs))

    def band_cells(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_BAND_CELL, (B, T, self.nbins))

    def band_curves(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_CURVES, (B, T, self.nbins, self.info.n_dirs))

    def doa_deg(self, cells=None):
        cells = self.cells() if cells is None else cells
        ang = np.array([capi.cell_angle(i, self.doa_step) for i in range(self.info.n_dirs)] + [0.0])   # cell -1: the initial DOA of 0 rad
        return ang[cells] * (180 / np.pi)

    def process(self, x):
        y = super().process(x)
        if self._callback is not None and self.frames_done:
            doa, prob, power, act = self.doa_deg(), self.prob(), self.power_db(), self.active()
            for b in range(self.info.n_streams):
                for t in range(self.frames_done):
                    if act[b, t]:                                           # _ptrCallback->setDOA(toDegrees(_currentDOA,1), _prob, power, 1)  (:245)
                        self._callback(doa[b, t:t + 1], prob[b, t:t + 1], float(power[b, t]), 1)
        return y


class FastBinauralMasking(Processor):
    def __init__(self, samplerate, microDistance, lowFreq, highFreq, mmethod="RELATIVE", algorithm="BOTH", n_streams=1,
                 max_frames_per_call=256, frame_size=None, n_bands=45, emit_spectra=False, emit_trace=False):
        """emit_spectra keeps the analysis / masked spectra fetchable (staged kernels); emit_trace keeps Q() / decisions() of the last call.
        Without them the whole chain is one fused kernel and only the audio (and the frame powers) come back."""
        N = frame_size or capi.frame_size(samplerate, 0.050)                # FastBinauralMasking.h:112
        H, fc, thr = capi.mel_bank(N, n_bands, samplerate, lowFreq, highFreq, microDistance)
        m = MaskingMethod[mmethod] if isinstance(mmethod, str) else mmethod
        a = MaskingAlg[algorithm] if isinstance(algorithm, str) else algorithm
        super().__init__(kind=capi.KIND_MASK, sample_rate=samplerate, frame_size=N, hop=N // 2, n_channels=2, n_streams=n_streams,
                         max_frames_per_call=max_frames_per_call, mask_method=m, mask_alg=a, n_bands=n_bands, band_coefs=H, band_thresholds=thr,
                         emit=(capi.EMIT_SPECTRA if emit_spectra else 0) | (capi.EMIT_MASK_TRACE if emit_trace else 0))
        self.H, self.fc, self.thresholds = H, fc, thr

    def masked_spectra(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_BEAMS, (B, T, 2, self.info.spectrum_pitch))[..., : self.info.one_sided_length]

    def Q(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_MASK_Q, (B, T, self.cfg.n_bands))

    def decisions(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_MASK_DEC, (B, T, self.cfg.n_bands))


class TdoaEstimator(Processor):
    """BASELINE config 2: integer-lag GCC-PHAT on all pairs (tau-vector mode of dsp::GeneralisedCrossCorrelation with integer taus)."""

    def __init__(self, sampleRate, n_channels, frame_size, max_lag, n_streams=1, max_frames_per_call=256, emit_curves=False, hop=None,
                 emit_spectra=False):
        super().__init__(kind=capi.KIND_TDOA, sample_rate=sampleRate, frame_size=frame_size, hop=hop or frame_size // 2, n_channels=n_channels,
                         n_streams=n_streams, max_frames_per_call=max_frames_per_call, max_lag=max_lag,
                         emit=(capi.EMIT_CURVES if emit_curves else 0) | (capi.EMIT_SPECTRA if emit_spectra else 0))
        self.max_lag = max_lag

    def lags(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_LAGS, (B, T, self.info.n_pairs))

    def curves(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_CURVES, (B, T, self.info.n_pairs, 2 * self.max_lag + 1))


class DelayAndSumFan(Processor):
    """BASELINE config 3: Beamformer::processFrame steered to D azimuths per frame."""

    def __init__(self, sampleRate, mic_xyz, frame_size, doas, n_streams=1, max_frames_per_call=64):
        mic_xyz = np.ascontiguousarray(mic_xyz, dtype=np.float64).reshape(-1, 3)
        turns = capi.steer_turns(mic_xyz, sampleRate, frame_size, doas)
        super().__init__(kind=capi.KIND_DSFAN, sample_rate=sampleRate, frame_size=frame_size, hop=frame_size // 2, n_channels=len(mic_xyz),
                         n_streams=n_streams, max_frames_per_call=max_frames_per_call, n_dirs=len(doas), steer_turns=turns)

    def beams(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_BEAMS, (B, T, self.info.n_dirs, self.info.beams_pitch))[..., : self.info.one_sided_length]


class FilterAndSumFan(Processor):
    """Filter-and-sum beamformer (BASELINE.json north_star; no reference class): D beams with loaded per-bin complex weights
    W [D][M][N/2+1], Y[d][k] = 1/M sum_c X_c[k] W[d][c][k].  With W = exp(j k phi_c(d)) it is DelayAndSumFan."""

    def __init__(self, sampleRate, n_channels, frame_size, weights, n_streams=1, max_frames_per_call=64):
        w = np.ascontiguousarray(weights, dtype=np.complex128)
        D, M, K = w.shape
        assert M == n_channels and K == frame_size // 2 + 1
        super().__init__(kind=capi.KIND_DSFAN, sample_rate=sampleRate, frame_size=frame_size, hop=frame_size // 2, n_channels=M, n_streams=n_streams,
                         max_frames_per_call=max_frames_per_call, n_dirs=D, fs_weights=w.view(np.float64).reshape(-1))

    def beams(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_BEAMS, (B, T, self.info.n_dirs, self.info.beams_pitch))[..., : self.info.one_sided_length]


class SrpPhat(Processor):
    """BASELINE config 4: SRP-PHAT energy map over a direction grid (channel form), smoothing + selectDOA as SteeringBeamforming."""

    def __init__(self, sampleRate, mic_xyz, frame_size, dirs, numOfSources=1, n_streams=1, max_frames_per_call=64):
        mic_xyz = np.ascontiguousarray(mic_xyz, dtype=np.float64).reshape(-1, 3)
        mt = capi.mic_tau(mic_xyz, sampleRate, dirs)
        super().__init__(kind=capi.KIND_SRP, sample_rate=sampleRate, frame_size=frame_size, hop=frame_size // 2, n_channels=len(mic_xyz),
                         n_streams=n_streams, max_frames_per_call=max_frames_per_call, n_dirs=mt.shape[1], mic_tau=mt, n_sources=numOfSources)

    def energy(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_ENERGY, (B, T, self.info.n_dirs))

    def cells(self):
        B, T = self._bt()
        return self.fetch(capi.OUT_CELL, (B, T, self.info.n_sources))

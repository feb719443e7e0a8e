"""Multi-GPU partitioning of the hot path.  One process per GPU, torch.distributed for the plumbing.

* Independent array streams (BASELINE configs 1, 2, 3, 5): contiguous blocks of streams per rank, no data-path collective
  (`stream_block`).  Each reference object is self-contained (no shared state, SURVEY.md §8b "Threading").
* One array whose direction grid is split over GPUs (BASELINE config 4): every rank computes the SRP-PHAT map of its slice of
  directions; per frame the local (maximum, first arg-max) is packed into one int64 and a single MAX all-reduce (NCCL over
  NVLink on GPUs, gloo in the CPU tests) yields the global arg-max cell (`ShardedSrpPhat`).  There is no other collective.
"""
import ctypes as C

import numpy as np
import torch

_IDX_BITS = 31
_IDX_MASK = (1 << _IDX_BITS) - 1


def stream_block(n_streams, rank, world):
    """[begin, end) of the contiguous block of streams rank `rank` of `world` owns."""
    return n_streams * rank // world, n_streams * (rank + 1) // world


def direction_block(n_dirs, rank, world):
    return n_dirs * rank // world, n_dirs * (rank + 1) // world


def pack_max(values, indices):
    """values float32 [...], global indices int [...]: int64 keys whose MAX is the largest value, lowest index on ties."""
    u = values.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    u = torch.where((u & 0x80000000) != 0, (~u) & 0xFFFFFFFF, u | 0x80000000)
    return (u << _IDX_BITS) | (_IDX_MASK - indices.to(torch.int64))


def unpack_max(packed):
    """int64 keys -> (float32 maxima, int64 indices)."""
    idx = _IDX_MASK - (packed & _IDX_MASK)
    u = (packed >> _IDX_BITS) & 0xFFFFFFFF
    u = torch.where((u & 0x80000000) != 0, u & 0x7FFFFFFF, (~u) & 0xFFFFFFFF)
    val = (u.to(torch.int64) - ((u >> 31) << 32)).to(torch.int32).view(torch.float32)
    return val, idx


def local_argmax_packed(energy, d_offset):
    """torch reference of mcag_k_argmax_pack: energy [rows][D_local] float32 -> packed int64 [rows]."""
    v, i = energy.max(dim=1)
    first = (energy == v[:, None]).to(torch.int64).argmax(dim=1)   # first maximum, as wipp::maxidx
    return pack_max(v, first + d_offset)


def allreduce_argmax(packed, group=None):
    """The single collective of the sharded-grid path."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
    return packed


class ShardedSrpPhat:
    """SRP-PHAT over a direction grid split across the ranks of a process group: same input on every rank, D/G directions each."""

    def __init__(self, sampleRate, mic_xyz, frame_size, dirs, n_streams=1, max_frames_per_call=64, group=None):
        import torch.distributed as dist
        from . import SrpPhat
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dirs = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 3)
        self.n_dirs = len(dirs)
        self.d0, self.d1 = direction_block(self.n_dirs, self.rank, self.world)
        self.local = SrpPhat(sampleRate, mic_xyz, frame_size, dirs[self.d0:self.d1], numOfSources=1, n_streams=n_streams,
                             max_frames_per_call=max_frames_per_call)
        dev = torch.device("cuda", self.local.cfg.device)
        self._packed = torch.empty(n_streams * max_frames_per_call, dtype=torch.int64, device=dev)
        from . import capi
        # torch view of the handle's own CUDA stream: the pack kernel, the all-reduce and the unpack are stream-ordered behind the
        # SRP kernels, so a step needs no host synchronisation
        self.stream = torch.cuda.ExternalStream(capi.lib().mcag_stream(self.local.handle), device=dev)
        self.time_allreduce, self._ar_events = False, []     # bench.py: CUDA events around the collective on the handle's stream

    def _reduce(self):
        from . import capi
        lib = capi.lib()
        p = self.local
        rows = p.info.n_streams * p.frames_done
        packed = self._packed[:rows]
        with torch.cuda.stream(self.stream):
            if rows:
                capi.check(lib.mcag_k_argmax_pack(C.c_void_p(lib.mcag_device_ptr(p.handle, capi.OUT_ENERGY)), C.c_longlong(rows), p.info.n_dirs, self.d0,
                                                  capi.vp(packed), C.c_void_p(lib.mcag_stream(p.handle))))
                if self.time_allreduce:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(self.stream)
                allreduce_argmax(packed, self.group)
                if self.time_allreduce:
                    e1.record(self.stream)
                    self._ar_events.append((e0, e1))
            val, idx = unpack_max(packed)
        return val.view(p.info.n_streams, p.frames_done), idx.view(p.info.n_streams, p.frames_done)   # explicit shape: a call may complete 0 frames

    def allreduce_ms(self):
        """mean device time of the all-reduce since the last call (needs time_allreduce); synchronises the stream"""
        self.stream.synchronize()
        ev, self._ar_events = self._ar_events, []
        return float(np.mean([a.elapsed_time(b) for a, b in ev])) if ev else None

    def process(self, x):
        """x [B*M][n] host array (identical on every rank) -> (peak energy [B][T], global direction cell [B][T]) on every rank."""
        self.local.process(x)
        return self._reduce()

    def process_device(self, d_in, pitch, nsamples):
        """device-resident input; asynchronous on the handle's stream (results are torch tensors ordered on `self.stream`)."""
        self.local.process_device(d_in, pitch, nsamples)
        return self._reduce()

"""Multi-GPU partitioning logic on CPU: world_size-2 gloo process groups (the N > 1 path of bench.py and ShardedSrpPhat)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_pack_unpack_round_trip_and_order():
    from mcarray_b200 import sharding
    v = torch.tensor([-3.5, -0.0, 0.0, 1e-30, 2.0, 2.0, 1e30, -1e30, float("-inf")], dtype=torch.float32)
    i = torch.arange(len(v))
    p = sharding.pack_max(v, i)
    assert (p >= 0).all()
    val, idx = sharding.unpack_max(p)
    assert torch.equal(val.view(torch.int32), v.view(torch.int32)) and torch.equal(idx, i)
    order = torch.argsort(p)                       # keys sort like the floats; equal floats: lower index = larger key
    assert torch.all(v[order][1:] >= v[order][:-1])
    assert p[4] > p[5]                             # tie 2.0 == 2.0 -> the lower index wins the MAX


def test_blocks_cover_everything():
    from mcarray_b200 import sharding
    for n in (1, 7, 64, 1024, 3600):
        for w in (1, 2, 3, 8):
            blocks = [sharding.stream_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n and all(blocks[r][1] == blocks[r + 1][0] for r in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, D, T, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcarray_b200 import sharding
    g = torch.Generator().manual_seed(7)
    energy = torch.randn(T, D, generator=g)
    energy[3, 10] = energy[3, D - 5] = 9.0         # a cross-shard tie: the lower cell must win
    energy[5] = -2.0                               # a flat, negative frame: cell 0
    d0, d1 = sharding.direction_block(D, rank, world)
    packed = sharding.local_argmax_packed(energy[:, d0:d1].contiguous(), d0)
    sharding.allreduce_argmax(packed)              # the one collective
    val, idx = sharding.unpack_max(packed)
    # stream sharding: every rank handles its block, results gathered without a data-path collective
    b0, b1 = sharding.stream_block(5, rank, world)
    torch.save({"val": val, "idx": idx, "streams": list(range(b0, b1))}, os.path.join(tmp, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_argmax_equals_unsharded_gloo(tmp_path, world):
    D, T = 101, 8
    port = 29650 + world
    mp.spawn(_worker, args=(world, port, D, T, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(7)
    energy = torch.randn(T, D, generator=g)
    energy[3, 10] = energy[3, D - 5] = 9.0
    energy[5] = -2.0
    want_val = energy.max(dim=1).values
    want_idx = torch.tensor([int(np.flatnonzero(energy[t].numpy() == want_val[t].item())[0]) for t in range(T)])
    streams = []
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(res["val"], want_val) and torch.equal(res["idx"], want_idx)
        streams += res["streams"]
    assert streams == list(range(5))
    assert want_idx[3] == 10 and want_idx[5] == 0


@pytest.mark.gpu
def test_argmax_pack_kernel_matches_torch():
    import ctypes as C
    from mcarray_b200 import capi, sharding
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = torch.Generator(device="cuda").manual_seed(3)
    e = torch.randn(300, 451, device="cuda", generator=g)
    e[7, 3] = e[7, 400] = 50.0
    e[9] = -1.0
    packed = torch.empty(300, dtype=torch.int64, device="cuda")
    capi.check(capi.lib().mcag_k_argmax_pack(capi.vp(e), C.c_longlong(300), 451, 1000, capi.vp(packed), None))
    torch.cuda.synchronize()
    assert torch.equal(packed.cpu(), sharding.local_argmax_packed(e.cpu(), 1000))
    val, idx = sharding.unpack_max(packed.cpu())
    assert idx[7] == 1003 and idx[9] == 1000 and torch.equal(val, e.max(dim=1).values.cpu())

import os, sys, time, ctypes as C
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import mcarray_b200 as mb
from mcarray_b200 import capi, sharding, scenes
import bench
mb.set_default_device(local)
wl=bench.Cfg4Sharded(); B,T=4,256; n=wl.N+(T-1)*wl.hop
x=torch.randn(B*wl.M, n, device="cuda")*1000
sp=wl.make(mb,B,T); p=sp.local
lib=capi.lib()
def run(mode, steps=20):
    for it in range(3+steps):
        if it==3:
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); t0=time.perf_counter()
        p.flush_input()
        p.process_device(x, n, n)
        if mode>=1:
            rows=B*T; packed=sp._packed[:rows]
            with torch.cuda.stream(sp.stream):
                capi.check(lib.mcag_k_argmax_pack(C.c_void_p(lib.mcag_device_ptr(p.handle, capi.OUT_ENERGY)), C.c_longlong(rows), p.info.n_dirs, sp.d0, capi.vp(packed), C.c_void_p(lib.mcag_stream(p.handle))))
                if mode>=2: dist.all_reduce(packed, op=dist.ReduceOp.MAX)
                if mode>=3: val,idx=sharding.unpack_max(packed)
    p.synchronize(); torch.cuda.synchronize()
    dt=(time.perf_counter()-t0)/steps*1e3
    if rank==0: print("mode",mode,"ms/step",round(dt,3),flush=True)
for m in (0,1,2,3,0): run(m)
dist.destroy_process_group()

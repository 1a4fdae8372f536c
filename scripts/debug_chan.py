import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_channel import Actors
from ray_b200 import _native, build
build.build(); _native.load()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for host_sync in (True, False):
    a = Actors(world, host_sync=host_sync)
    shape = (4 * world, 6)
    xs = [(1.0 + 0.1 * torch.randn(shape, generator=torch.Generator().manual_seed(r))).to(torch.float16) for r in range(world)]
    bad = 0
    for rep in range(20):
        for op in range(5):
            def f(r, c):
                s = xs[r].to(a.dev(r))
                out = torch.empty_like(s)
                c.allreduce(s, out, op)
                rs = torch.empty((shape[0] // world, shape[1]), dtype=s.dtype, device=s.device)
                c.reducescatter(s, rs, op)
                return out.cpu(), rs.cpu()
            res = a.run(f)
            for r in range(world):
                if not torch.equal(res[r][0], res[0][0]) or (res[r][1] == 0).all() or (res[r][0] == 0).all():
                    bad += 1
                    print("host_sync", host_sync, "rep", rep, "op", op, "rank", r, "ar zero", bool((res[r][0] == 0).all()), "rs zero", bool((res[r][1] == 0).all()))
    print("host_sync", host_sync, "bad", bad)
    a.close()

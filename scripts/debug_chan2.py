import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_channel import Actors
from ray_b200 import _native, build
build.build(); _native.load()
world = 3
a = Actors(world, host_sync=True)
shape = (4 * world, 6)
xs = [(1.0 + 0.1 * torch.randn(shape, generator=torch.Generator().manual_seed(r))).to(torch.float16) for r in range(world)]
dev_s = [xs[r].to(a.dev(r)) for r in range(world)]
dev_out = [torch.empty_like(dev_s[r]) for r in range(world)]
dev_rs = [torch.zeros((shape[0] // world, shape[1]), dtype=torch.float16, device=a.dev(r)) for r in range(world)]
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else "prealloc"
for rep in range(6):
    for op in range(5):
        t0 = time.time()
        def f(r, c):
            tt = [time.time()]
            if mode == "prealloc":
                s, out, rs = dev_s[r], dev_out[r], dev_rs[r]
            else:
                s = xs[r].to(a.dev(r)); out = torch.empty_like(s)
            tt.append(time.time())
            c.allreduce(s, out, op)
            tt.append(time.time())
            if mode != "prealloc":
                rs = torch.empty((shape[0] // world, shape[1]), dtype=s.dtype, device=s.device)
            tt.append(time.time())
            c.reducescatter(s, rs, op)
            tt.append(time.time())
            z = bool((rs == 0).all().item())
            return z, [round(b - a_, 3) for a_, b in zip(tt, tt[1:])]
        try:
            res = a.run(f)
        except Exception as e:
            print("rep", rep, "op", op, "EXC", type(e).__name__, str(e)[:80]); raise SystemExit
        print("rep", rep, "op", op, "dt %.2f" % (time.time() - t0), res, flush=True)
a.close()

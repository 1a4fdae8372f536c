"""First-contact probe: what does the GPU box offer, and does the basic path work?

    python scripts/gpu_probe.py [world_size ...]
"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ray_b200 import _native as N
from ray_b200.testing import LocalGroup


def main():
    print("torch", torch.__version__, "cuda", torch.version.cuda, "devices", torch.cuda.device_count())
    for i in range(torch.cuda.device_count()):
        p = torch.cuda.get_device_properties(i)
        print(f"  dev{i}: {p.name} sm_{p.major}{p.minor} SMs={p.multi_processor_count} mem={p.total_memory >> 30} GiB")
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    print(N.load().b200_version())
    worlds = [int(a) for a in sys.argv[1:]] or [2, 4, 8]
    for n in worlds:
        t0 = time.time()
        try:
            with LocalGroup(n, timeout_ms=8000) as g:
                print(f"[n={n}] group up in {time.time() - t0:.2f}s devices={g.devices} shared={g.shared_gpu} "
                      f"multicast={g.has_multicast}")
                for numel in (1000, 1 << 20):
                    for algo, name in ((N.ALGO_ONESHOT, "oneshot"), (N.ALGO_TWOSHOT, "twoshot"),
                                       (N.ALGO_NVLS, "nvls")):
                        if algo == N.ALGO_NVLS and not g.has_multicast:
                            continue
                        xs = [torch.randn(numel, device=g.device(r), generator=None) for r in range(n)]
                        ref = xs[0].clone().to("cuda:0")
                        for x in xs[1:]:
                            ref += x.to("cuda:0")
                        g.run(lambda c, r: c.allreduce(xs[r], N.SUM, algo=algo))
                        err = max((x.to("cuda:0") - ref).abs().max().item() for x in xs)
                        same = all(torch.equal(xs[0].to("cuda:0"), x.to("cuda:0")) for x in xs[1:])
                        print(f"[n={n}] allreduce f32 {name} numel={numel}: max_err={err:.3e} replicas_equal={same}")
                # timing of a mid-size allreduce
                numel = 4 << 20
                xs = [torch.randn(numel, device=g.device(r)) for r in range(n)]
                for algo, name in ((N.ALGO_TWOSHOT, "twoshot"), (N.ALGO_NVLS, "nvls")):
                    if algo == N.ALGO_NVLS and not g.has_multicast:
                        continue
                    for _ in range(3):
                        g.run(lambda c, r: c.allreduce(xs[r], N.SUM, algo=algo))
                    torch.cuda.synchronize()
                    t1 = time.time()
                    iters = 10
                    for _ in range(iters):
                        for r, c in enumerate(g.comms):
                            with torch.cuda.device(g.devices[r]), torch.cuda.stream(g.streams[r]):
                                c.allreduce(xs[r], N.SUM, algo=algo)
                    g.synchronize()
                    dt = (time.time() - t1) / iters
                    print(f"[n={n}] {name} 16 MiB: {dt * 1e6:.1f} us/iter  algbw={numel * 4 / dt / 1e9:.1f} GB/s")
        except Exception:
            print(f"[n={n}] FAILED")
            traceback.print_exc()


if __name__ == "__main__":
    main()

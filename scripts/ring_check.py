"""Chunk ring on/off for messages larger than the staging slot (n >= 3, NVLS pipeline)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ray_b200 import _native as N
from ray_b200.testing import LocalGroup
from scripts.bw_sweep import time_graphs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
MiB = 1 << 20
g = LocalGroup(n, timeout_ms=20000, staging_bytes=256 << 20, inbox_bytes=8 << 20)
print("world", n, "multicast", g.has_multicast)
sizes = (1024 * MiB,) if len(sys.argv) > 2 else (512 * MiB, 1024 * MiB)
chunks = (-1,) if len(sys.argv) > 2 else (4, 8)
# correctness of the ring at this world size: exact integer-valued pattern, 1 GiB through one launch
chk = [(torch.arange(sizes[-1] // 4, dtype=torch.float32, device=g.device(r)) % 509) * (r + 1) for r in range(n)]
before = g.comms[0].launch_count
g.run(lambda c, r: c.allreduce(chk[r], N.SUM))
want = (torch.arange(sizes[-1] // 4, dtype=torch.float32, device=g.device(0)) % 509) * (n * (n + 1) // 2)
print("ring result exact:", all(torch.equal(x.to(g.device(0)), want) for x in chk), "launches", g.comms[0].launch_count - before, flush=True)
del chk, want
for size in sizes:
    xs = [torch.ones(size // 4, device=g.device(r)) for r in range(n)]
    for chunk in chunks:
        for ring in (0, -1):
            for c in g.comms:
                c.set_param(N.PARAM_PIPE_CHUNK_BYTES, chunk * MiB if chunk > 0 else -1)
                c.set_param(N.PARAM_PIPE_RING, ring)
            us = time_graphs(g, lambda c, r: c.allreduce(xs[r], N.SUM), 6)
            print(f"allreduce n={n} {size >> 20} MiB chunk={chunk}MiB ring={'on' if ring else 'off'}: {us:9.1f} us busbw {size / us / 1e3 * 2 * (n - 1) / n:7.1f} GB/s", flush=True)
    del xs
g.destroy()

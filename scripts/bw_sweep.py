"""Device-timed bandwidth / latency sweep of the collective kernels (one process, one rank per
GPU, launches replayed from CUDA graphs so host launch overhead does not pollute the numbers).

    python scripts/bw_sweep.py --world 2 [--algos oneshot,twoshot,nvls] [--blocks 0,32,64] \
        [--min 1024 --max 1073741824] [--op allreduce|allgather|reducescatter|broadcast|sendrecv|grad]

Prints one line per (size, algo, blocks): us per launch, algbw, busbw (nccl-tests convention).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ray_b200 import _native as N
from ray_b200.testing import LocalGroup

ALGOS = {"auto": N.ALGO_AUTO, "oneshot": N.ALGO_ONESHOT, "twoshot": N.ALGO_TWOSHOT, "nvls": N.ALGO_NVLS,
         "ll": N.ALGO_LL}


def time_graphs(g, make_call, iters, reps=3):
    """Capture `iters` launches per rank into one graph per rank; replay; return us per launch
    (max over ranks, best of reps)."""
    graphs = []
    g.run(lambda c, r: make_call(c, r))  # one eager warm-up outside capture (module load, lazy init)
    for r, c in enumerate(g.comms):
        torch.cuda.set_device(g.devices[r])
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=g.streams[r]):
            for _ in range(iters):
                make_call(c, r)
        graphs.append(gr)
    best = None
    for _ in range(reps + 1):
        starts, ends = [], []
        for r in range(g.world_size):
            torch.cuda.set_device(g.devices[r])
            with torch.cuda.stream(g.streams[r]):
                st = torch.cuda.Event(enable_timing=True)
                en = torch.cuda.Event(enable_timing=True)
                st.record()
                graphs[r].replay()
                en.record()
                starts.append(st)
                ends.append(en)
        g.synchronize()
        us = max(s.elapsed_time(e) for s, e in zip(starts, ends)) * 1e3 / iters
        best = us if best is None else min(best, us)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--algos", default="auto")
    ap.add_argument("--blocks", default="0")
    ap.add_argument("--min", type=int, default=1 << 10)
    ap.add_argument("--max", type=int, default=1 << 30)
    ap.add_argument("--step", type=int, default=4, help="size multiplier between points")
    ap.add_argument("--op", default="allreduce")
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--symm", action="store_true", help="operands in the symmetric heap (zero copy)")
    ap.add_argument("--nvls-min-world", type=int, default=-1)
    ap.add_argument("--nvls-ctas", default="-1", help="comma list of CTA counts for the NVLS reduce phase")
    args = ap.parse_args()
    n = args.world
    dtype = getattr(torch, args.dtype)
    heap = (args.max + (4 << 20)) if args.symm else 0
    g = LocalGroup(n, timeout_ms=20000, staging_bytes=512 << 20, heap_bytes=heap, inbox_bytes=64 << 20)
    print(f"# world={n} devices={g.devices} shared={g.shared_gpu} multicast={g.has_multicast} op={args.op} "
          f"dtype={args.dtype} symm={args.symm}")
    for c in g.comms:
        c.set_param(N.PARAM_NVLS_MIN_WORLD, args.nvls_min_world)
    size = args.min
    es = torch.empty((), dtype=dtype).element_size()
    while size <= args.max:
        numel = size // es
        iters = 50 if size <= (1 << 20) else (20 if size <= (64 << 20) else 6)
        for blocks in [int(b) for b in args.blocks.split(",")]:
            if not g.shared_gpu:
                for c in g.comms:
                    c.set_blocks(blocks)
            for aname, nctas in [(a, int(x)) for a in args.algos.split(",") for x in args.nvls_ctas.split(",")]:
                algo = ALGOS[aname]
                for c in g.comms:
                    c.set_param(N.PARAM_NVLS_CTAS, nctas)
                if algo == N.ALGO_NVLS and not g.has_multicast:
                    continue
                if algo == N.ALGO_ONESHOT and size > (8 << 20):
                    continue
                if algo == N.ALGO_LL and size > (64 << 10):
                    continue
                if args.symm:
                    for c in g.comms:
                        c.symm_reset()
                    xs = [g.comms[r].symm_empty((numel,), dtype) for r in range(n)]
                    for x in xs:
                        x.fill_(1)
                else:
                    xs = [torch.ones(numel, dtype=dtype, device=g.device(r)) for r in range(n)]
                if args.op == "allreduce":
                    call = lambda c, r: c.allreduce(xs[r], N.SUM, algo=algo)  # noqa: E731
                    factor = 2 * (n - 1) / n
                elif args.op == "grad":
                    call = lambda c, r: c.grad_allreduce(xs[r], 1.0 / n, torch.bfloat16)  # noqa: E731
                    factor = 2 * (n - 1) / n
                elif args.op == "allgather":
                    outs = [torch.empty(n * numel, dtype=dtype, device=g.device(r)) for r in range(n)]
                    call = lambda c, r: c.allgather_into(outs[r], xs[r])  # noqa: E731
                    factor = (n - 1)  # S_total = n*size; busbw = n*size/t*(n-1)/n
                elif args.op == "reducescatter":
                    ins = [torch.ones(n * numel, dtype=dtype, device=g.device(r)) for r in range(n)]
                    call = lambda c, r: c.reducescatter_from(xs[r], ins[r], N.SUM)  # noqa: E731
                    factor = (n - 1)
                elif args.op == "broadcast":
                    call = lambda c, r: c.broadcast(xs[r], 0)  # noqa: E731
                    factor = 1.0
                elif args.op == "sendrecv":
                    call = lambda c, r: (c.send(xs[0], 1) if r == 0 else (c.recv(xs[1], 0) if r == 1 else None))  # noqa: E731
                    factor = 1.0
                else:
                    raise SystemExit(f"unknown op {args.op}")
                torch.cuda.synchronize()
                us = time_graphs(g, call, iters)
                algbw = size / us / 1e3
                print(f"{args.op} {size:>11d} B  algo={aname:8s} blocks={blocks:3d} nvls_ctas={nctas:3d} {us:10.2f} us  "
                      f"algbw={algbw:8.1f} GB/s  busbw={algbw * factor:8.1f} GB/s", flush=True)
                del xs
        size *= args.step
    g.destroy()


if __name__ == "__main__":
    main()

"""The reference's CPU path, executed: torch.distributed gloo driven through the exact call
sequence of ``TorchGLOOGroup``.  TEST INFRASTRUCTURE ONLY (see collective_oracle.py).

``ray`` cannot be imported here (no wheel, ``ray._raylet`` needs the Bazel-built C++ core),
but ``TorchGLOOGroup`` is a thin wrapper over ``torch.distributed``; the functions below
issue the same c10d calls in the same order, citing the wrapper line they restate:

    allreduce      dist.all_reduce(t, op)                         torch_gloo_collective_group.py:208-217
    reduce         root: dist.reduce(t); others: on a clone        :222-240
    allgather      dist.all_gather(list, t)                        :242-252
    broadcast      dist.broadcast(t, src)                          :254-258
    reducescatter  all_reduce every list member, then copy [rank]  :260-282
    send / recv    dist.send / dist.recv                           :284-290

``run(world_size, jobs)`` spawns one process per rank (rendezvous over a temp file, the
stand-in for the GCS-KV rendezvous at :128-150) and returns every rank's results.
"""
from __future__ import annotations

import os
import tempfile
import time
from typing import Any, Dict, List

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

# ray.util.collective.types.ReduceOp value -> torch op (torch_gloo_collective_group.py:40-45)
TORCH_REDUCE_OP = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.PRODUCT, 2: dist.ReduceOp.MIN, 3: dist.ReduceOp.MAX}


def gloo_allreduce(t: torch.Tensor, op: int) -> None:
    dist.all_reduce(t, op=TORCH_REDUCE_OP[op])


def gloo_reduce(t: torch.Tensor, root: int, op: int) -> None:
    if dist.get_rank() == root:
        dist.reduce(t, dst=root, op=TORCH_REDUCE_OP[op])
    else:
        tmp = t.detach().clone()
        dist.reduce(tmp, dst=root, op=TORCH_REDUCE_OP[op])


def gloo_allgather(outs: List[torch.Tensor], t: torch.Tensor) -> None:
    dist.all_gather(outs, t)


def gloo_broadcast(t: torch.Tensor, root: int) -> None:
    dist.broadcast(t, src=root)


def gloo_reducescatter(out: torch.Tensor, ins: List[torch.Tensor], op: int) -> None:
    rank = dist.get_rank()
    if out.shape != ins[rank].shape:
        raise ValueError("Output tensor has wrong shape")
    for t in ins:
        dist.all_reduce(t, op=TORCH_REDUCE_OP[op])
    if out.data_ptr() != ins[rank].data_ptr():
        out.copy_(ins[rank])


def _to_torch(a: np.ndarray) -> torch.Tensor:
    if a.dtype.name == "bfloat16":
        return torch.from_numpy(a.view(np.uint16).copy()).view(torch.bfloat16)
    return torch.from_numpy(a.copy())


def _to_numpy(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        import ml_dtypes

        return t.view(torch.uint16).numpy().view(ml_dtypes.bfloat16)
    return t.numpy()


def _worker(rank: int, world: int, init_file: str, jobs: List[Dict[str, Any]], out_dir: str) -> None:
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    results = []
    for job in jobs:
        kind = job["kind"]
        if kind == "allreduce":
            t = _to_torch(job["inputs"][rank])
            gloo_allreduce(t, job["op"])
            results.append(_to_numpy(t))
        elif kind == "reduce":
            t = _to_torch(job["inputs"][rank])
            gloo_reduce(t, job["root"], job["op"])
            results.append(_to_numpy(t))
        elif kind == "broadcast":
            t = _to_torch(job["inputs"][rank])
            gloo_broadcast(t, job["root"])
            results.append(_to_numpy(t))
        elif kind == "allgather":
            t = _to_torch(job["inputs"][rank])
            outs = [torch.empty_like(t) for _ in range(world)]
            gloo_allgather(outs, t)
            results.append(np.stack([_to_numpy(o) for o in outs]))
        elif kind == "reducescatter":
            ins = [_to_torch(a) for a in job["inputs"][rank]]
            out = torch.empty_like(ins[rank])
            gloo_reducescatter(out, ins, job["op"])
            results.append(_to_numpy(out))
        elif kind == "sendrecv":
            t = _to_torch(job["inputs"][rank])
            if rank == job["src"]:
                dist.send(t, dst=job["dst"])
            elif rank == job["dst"]:
                dist.recv(t, src=job["src"])
            results.append(_to_numpy(t))
        elif kind == "time_allreduce":
            # wall-clock timing of the reference CPU path (bench.py's cpu_baseline leg)
            t = torch.ones(job["numel"], dtype=torch.float32) * (rank + 1)
            for _ in range(job["warmup"]):
                gloo_allreduce(t, 0)
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(job["iters"]):
                gloo_allreduce(t, 0)
            dist.barrier()
            results.append(np.array([(time.perf_counter() - t0) / job["iters"]]))
        else:
            raise ValueError(kind)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array(results, dtype=object), allow_pickle=True)
    dist.destroy_process_group()


def run(world: int, jobs: List[Dict[str, Any]]) -> List[List[np.ndarray]]:
    """Execute ``jobs`` on ``world`` gloo ranks; returns results[rank][job]."""
    with tempfile.TemporaryDirectory(prefix="b200_gloo_") as d:
        init_file = os.path.join(d, "rendezvous")
        mp.spawn(_worker, args=(world, init_file, jobs, d), nprocs=world, join=True)
        return [list(np.load(os.path.join(d, f"rank{r}.npy"), allow_pickle=True)) for r in range(world)]

"""Generates tests/golden/collective_golden.npz from the REFERENCE.

Run in the build container (needs /root/reference and 8 host processes):

    python tests/golden/make_golden.py

Two independent sources pin the oracle:

  A. ``CPUCommBarrier._apply_op`` -- the reference's own explicit reduction
     (python/ray/experimental/channel/cpu_communicator.py:69-89).  ``ray`` is not importable
     (no ray._raylet), so the method's source is cut out of the reference file with ``ast``
     and executed against the reference's real ``ReduceOp`` enum
     (python/ray/experimental/util/types.py:11-17, also extracted from source).
  B. torch.distributed gloo driven through TorchGLOOGroup's exact call sequence
     (oracle/gloo_reference.py; python/ray/util/collective/collective_group/
     torch_gloo_collective_group.py:208-290) on world sizes 2, 3, 4, 8.

Inputs are stored next to the outputs so the fixtures do not depend on any RNG.
"""
from __future__ import annotations

import ast
import os
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/python/ray"

from oracle import gloo_reference  # noqa: E402


def _extract(path: str, name: str) -> str:
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == name:
            return textwrap.dedent(ast.get_source_segment(src, node))
    raise KeyError(name)


def load_reference_apply_op():
    """Returns (apply_op(op, tensors), ReduceOp) built from the reference sources."""
    from enum import Enum

    ns = {"Enum": Enum, "torch": torch, "List": list}
    enum_src = _extract(f"{REF}/experimental/util/types.py", "ReduceOp")
    enum_src = "\n".join(l for l in enum_src.splitlines() if not l.strip().startswith("@"))
    exec(enum_src, ns)  # noqa: S102 - reference source, executed only to generate fixtures
    fn_src = _extract(f"{REF}/experimental/channel/cpu_communicator.py", "_apply_op")
    exec(fn_src, ns)  # noqa: S102
    fn = ns["_apply_op"]
    return (lambda op, tensors: fn(None, op, tensors)), ns["ReduceOp"]


DTYPES = {
    "uint8": torch.uint8, "int8": torch.int8, "int32": torch.int32, "int64": torch.int64,
    "float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32, "float64": torch.float64,
}
OPS = {"SUM": 0, "PRODUCT": 1, "MIN": 2, "MAX": 3}  # ray.util.collective.types.ReduceOp
NUMEL = 257  # odd on purpose: exercises the 16-byte tail path of the kernels


def make_inputs(world: int, dtype: torch.dtype, op: str, seed: int) -> list:
    g = torch.Generator().manual_seed(seed)
    outs = []
    for r in range(world):
        if dtype.is_floating_point:
            x = torch.randn(NUMEL, generator=g, dtype=torch.float64)
            if op == "PRODUCT":
                x = 1.0 + 0.1 * x  # keep products finite in fp16
            outs.append(x.to(dtype))
        else:
            info = torch.iinfo(dtype)
            lo, hi = (0, 7) if op == "PRODUCT" else (max(info.min, -1000), min(info.max, 1000))
            outs.append(torch.randint(lo, hi + 1, (NUMEL,), generator=g, dtype=torch.int64).to(dtype))
    return outs


def to_np(t: torch.Tensor) -> np.ndarray:
    # bf16 is stored as its uint16 bit pattern (npz has no bfloat16)
    return t.view(torch.uint16).numpy() if t.dtype == torch.bfloat16 else t.numpy()


def main():
    apply_op, RefReduceOp = load_reference_apply_op()
    ref_op = {"SUM": RefReduceOp.SUM, "PRODUCT": RefReduceOp.PRODUCT, "MIN": RefReduceOp.MIN, "MAX": RefReduceOp.MAX}
    store = {}
    index = []
    seed = 20260921
    for world in (2, 3, 4, 8):
        jobs, keys = [], []
        for dname, dtype in DTYPES.items():
            for oname, ocode in OPS.items():
                seed += 1
                xs = make_inputs(world, dtype, oname, seed)
                key = f"allreduce/w{world}/{dname}/{oname}"
                store[key + "/in"] = np.stack([to_np(x) for x in xs])
                store[key + "/apply_op"] = to_np(apply_op(ref_op[oname], [x.clone() for x in xs]))
                jobs.append({"kind": "allreduce", "op": ocode,
                             "inputs": [gloo_reference._to_numpy(x) for x in xs]})
                keys.append(key)
                index.append(key)
        # data-movement ops and reduce / reducescatter on fp32 + int32
        for dname in ("float32", "int32"):
            dtype = DTYPES[dname]
            seed += 1
            xs = make_inputs(world, dtype, "SUM", seed)
            arr = [gloo_reference._to_numpy(x) for x in xs]
            root = world - 1
            for kind, extra in (("broadcast", {"root": root}), ("allgather", {}),
                                ("reduce", {"root": root, "op": 0}),
                                ("sendrecv", {"src": 0, "dst": world - 1})):
                key = f"{kind}/w{world}/{dname}"
                store[key + "/in"] = np.stack(arr)
                jobs.append({"kind": kind, "inputs": arr, **extra})
                keys.append(key)
                index.append(key)
            seed += 1
            lists = [[gloo_reference._to_numpy(make_inputs(world, dtype, "SUM", seed * 100 + q * 10 + i)[0])
                      for i in range(world)] for q in range(world)]
            key = f"reducescatter/w{world}/{dname}"
            store[key + "/in"] = np.stack([np.stack(l) for l in lists])  # [rank q][slot i][numel]
            jobs.append({"kind": "reducescatter", "op": 0, "inputs": lists})
            keys.append(key)
            index.append(key)
        results = gloo_reference.run(world, jobs)  # results[rank][job]
        for j, key in enumerate(keys):
            per_rank = [results[r][j] for r in range(world)]
            per_rank = [a.view(np.uint16) if a.dtype.name == "bfloat16" else a for a in per_rank]
            store[key + "/gloo"] = np.stack(per_rank)
        print(f"world {world}: {len(keys)} cases")
    store["index"] = np.array(index)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "collective_golden.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()

"""Generates tests/golden/collective_golden_large.npz: one >= 1 MiB NON-INTEGER (seeded randn fp32)
case per collective, computed by the reference's CPU backend call sequence on real gloo
(oracle/gloo_reference.py; torch_gloo_collective_group.py:208-290), at world sizes 2 and 4.

    python tests/golden/make_golden_large.py          # build container: needs 4 host processes

A 1 MiB-per-rank case would cost several MiB of fixtures if inputs and outputs were stored, so
the fixture stores the RECIPE of the inputs (torch CPU generator seeds -- deterministic for the
pinned torch build of this image) and, of the gloo outputs, a strided sample (every 997th element
plus the first and last 64) and the sha256 of the full byte image.  ``recipe()`` below is the
single definition of the inputs; the tests import it.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

NUMEL = (1 << 18) + 3  # 1 MiB + 12 bytes of fp32: not a multiple of 16 bytes on purpose
WORLDS = (2, 4)
BASE_SEED = 20260921


def recipe(world: int, kind: str):
    """The inputs of one case, as numpy fp32 arrays.  allreduce / reduce / broadcast / allgather:
    one [NUMEL] tensor per rank; reducescatter: per rank a list of `world` [NUMEL] tensors."""
    salt = {"allreduce": 1, "reduce": 2, "broadcast": 3, "allgather": 4, "reducescatter": 5}[kind]

    def one(tag):
        g = torch.Generator().manual_seed(BASE_SEED + 1000 * salt + 100 * world + tag)
        return torch.randn(NUMEL, generator=g).numpy()

    if kind == "reducescatter":
        return [[one(10 * q + i) for i in range(world)] for q in range(world)]
    return [one(r) for r in range(world)]


def sample_index() -> np.ndarray:
    idx = np.concatenate([np.arange(64), np.arange(64, NUMEL - 64, 997), np.arange(NUMEL - 64, NUMEL)])
    return np.unique(idx).astype(np.int64)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from oracle import gloo_reference

    store = {"sample_index": sample_index()}
    idx = store["sample_index"]
    for world in WORLDS:
        root = world - 1
        jobs = [{"kind": "allreduce", "op": 0, "inputs": recipe(world, "allreduce")},
                {"kind": "reduce", "op": 0, "root": root, "inputs": recipe(world, "reduce")},
                {"kind": "broadcast", "root": root, "inputs": recipe(world, "broadcast")},
                {"kind": "allgather", "inputs": recipe(world, "allgather")},
                {"kind": "reducescatter", "op": 0, "inputs": recipe(world, "reducescatter")}]
        res = gloo_reference.run(world, jobs)  # res[rank][job]
        for j, job in enumerate(jobs):
            key = f"{job['kind']}/w{world}"
            per_rank = [np.asarray(res[r][j]) for r in range(world)]
            if job["kind"] == "allgather":  # [rank][source][NUMEL]
                store[key + "/sample"] = np.stack([np.stack([s[idx] for s in pr]) for pr in per_rank])
                store[key + "/sha256"] = np.array([digest(np.stack(pr)) for pr in per_rank])
            else:
                store[key + "/sample"] = np.stack([pr[idx] for pr in per_rank])
                store[key + "/sha256"] = np.array([digest(pr) for pr in per_rank])
        print(f"world {world}: {len(jobs)} cases")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "collective_golden_large.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()

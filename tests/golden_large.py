"""Shared checker for the >= 1 MiB non-integer fixtures (tests/golden/collective_golden_large.npz,
generated from real gloo by tests/golden/make_golden_large.py).  Used by the CPU test (oracle) and
the GPU test (CUDA kernels through the C ABI) with the same tolerances:
  copies (all-gather, broadcast): bit exact, sha256 of the full output;
  sums at world 2 (a single fp32 add): bit exact, sha256;
  sums at world 4: |out - gloo| <= 1e-6 * sum_r |x_r| on the sampled elements."""
import os

import numpy as np

from tests.golden.make_golden_large import NUMEL, WORLDS, digest, recipe  # noqa: F401

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collective_golden_large.npz")
KINDS = ("allreduce", "reduce", "broadcast", "allgather", "reducescatter")


def load():
    return np.load(PATH, allow_pickle=False)


def check(fix, kind: str, world: int, outputs):
    """``outputs[rank]`` = what rank produced for this case (allgather: array [world, NUMEL]);
    for ``reduce`` only the root's buffer is compared (the others must equal their inputs)."""
    idx = fix["sample_index"]
    key = f"{kind}/w{world}"
    sample, sha = fix[key + "/sample"], fix[key + "/sha256"]
    ins = recipe(world, kind)
    root = world - 1
    for r in range(world):
        out = np.asarray(outputs[r])
        if kind in ("allgather", "broadcast"):
            assert digest(out) == str(sha[r]), (key, r, "copy must be bit exact")
            continue
        if kind == "reduce" and r != root:
            assert np.array_equal(out, ins[r]), (key, r, "non-root buffers must be untouched")
            continue
        if world == 2:
            assert digest(out) == str(sha[r]), (key, r, "a single fp32 add must be bit exact")
            continue
        if kind == "reducescatter":
            sum_abs = np.sum([np.abs(ins[q][r].astype(np.float64)) for q in range(world)], axis=0)
        else:
            sum_abs = np.sum([np.abs(x.astype(np.float64)) for x in ins], axis=0)
        err = np.abs(out[idx].astype(np.float64) - sample[r].astype(np.float64))
        assert np.all(err <= 1e-6 * sum_abs[idx]), (key, r, float((err / sum_abs[idx]).max()))

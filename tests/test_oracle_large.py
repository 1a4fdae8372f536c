"""The oracle against the >= 1 MiB non-integer gloo fixtures (CPU): pins the rank-ascending
restatement at BASELINE-like message sizes, and validates the fixture recipe the GPU test uses."""
import numpy as np
import pytest

from oracle import collective_oracle as O
from tests import golden_large as G


@pytest.mark.parametrize("world", G.WORLDS)
@pytest.mark.parametrize("kind", G.KINDS)
def test_oracle_matches_large_gloo_fixtures(kind, world):
    fix = G.load()
    ins = G.recipe(world, kind)
    root = world - 1
    if kind == "allreduce":
        red = O.reduce_rank_ascending(ins, O.SUM)
        outs = [red] * world
    elif kind == "reduce":
        outs = [O.reduce_rank_ascending(ins, O.SUM) if r == root else ins[r] for r in range(world)]
    elif kind == "broadcast":
        outs = [ins[root]] * world
    elif kind == "allgather":
        outs = [np.stack(ins)] * world
    else:
        outs = [O.reduce_rank_ascending([ins[q][r] for q in range(world)], O.SUM) for r in range(world)]
    G.check(fix, kind, world, outs)

"""Property tests of the CPU oracle (hypothesis): the algebra every parity test leans on."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

from oracle import collective_oracle as O

INT_DTYPES = [np.uint8, np.int8, np.int32, np.int64]


def stacks(dtype, min_world=1, max_world=8):
    """A list of `world` equally shaped 1-D arrays."""
    return st.integers(min_world, max_world).flatmap(
        lambda w: st.integers(1, 40).flatmap(
            lambda n: st.lists(hnp.arrays(dtype, n, elements=hnp.from_dtype(np.dtype(dtype), allow_nan=False,
                                                                          allow_infinity=False)),
                               min_size=w, max_size=w)))


@settings(max_examples=60, deadline=None)
@given(st.sampled_from(INT_DTYPES).flatmap(stacks), st.sampled_from([O.SUM, O.PRODUCT, O.MIN, O.MAX]),
       st.randoms(use_true_random=False))
def test_integer_reductions_do_not_depend_on_rank_order(xs, op, rnd):
    """Wrapping integer SUM/PRODUCT and MIN/MAX are commutative and associative: the reason the
    integer parity bar is bit-exactness against ANY backend order (gloo ring, NCCL tree, NVLS)."""
    want = O.reduce_rank_ascending(xs, op)
    perm = list(xs)
    rnd.shuffle(perm)
    assert np.array_equal(O.reduce_rank_ascending(perm, op), want)
    assert want.dtype == xs[0].dtype


@settings(max_examples=40, deadline=None)
@given(stacks(np.float32, min_world=2, max_world=8))
def test_fp32_sum_meets_the_north_star_tolerance_against_fp64(xs):
    xs = [np.clip(x, -1e30, 1e30) for x in xs]
    got = O.reduce_rank_ascending(xs, O.SUM).astype(np.float64)
    ref = np.sum([x.astype(np.float64) for x in xs], axis=0)
    bound = 1e-6 * np.sum([np.abs(x.astype(np.float64)) for x in xs], axis=0)
    assert np.all(np.abs(got - ref) <= bound + 1e-300)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 8), st.integers(1, 6), st.integers(0, 2 ** 31))
def test_allgather_then_local_reduce_equals_allreduce_and_reducescatter(world, d, seed):
    rng = np.random.default_rng(seed)
    xs = [rng.integers(-1000, 1000, (world * d, 3)).astype(np.int64) for _ in range(world)]
    ar = O.cgraph_allreduce(xs, 0)[0]
    cat = O.cgraph_allgather(xs)[0].reshape(world, world * d, 3)
    assert np.array_equal(cat.sum(0), ar)
    rs = O.cgraph_reducescatter(xs, 0)
    assert np.array_equal(np.concatenate(rs, axis=0), ar)
    # util.collective flavour: reducescatter of per-rank lists == slices of the all-reduce
    lists = [[x[i * d:(i + 1) * d].copy() for i in range(world)] for x in xs]
    outs = [np.empty((d, 3), np.int64) for _ in range(world)]
    O.reducescatter(outs, lists, O.SUM)
    assert np.array_equal(np.concatenate(outs, axis=0), ar)


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 8), st.integers(1, 50), st.integers(0, 2 ** 31))
def test_half_precision_accumulate_modes_agree_at_world_two_only(world, n, seed):
    rng = np.random.default_rng(seed)
    xs = [rng.standard_normal(n).astype(np.float16) for _ in range(world)]
    native = O.reduce_rank_ascending(xs, O.SUM, accumulate="native")
    wide = O.reduce_rank_ascending(xs, O.SUM, accumulate="fp32")
    if world == 2:
        assert np.array_equal(native, wide)  # one add: fp32-then-round == fp16 add
    err = np.abs(native.astype(np.float64) - wide.astype(np.float64))
    assert np.all(err <= 2.0 ** -9 * np.sum([np.abs(x.astype(np.float64)) for x in xs], axis=0) + 1e-12)


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 8), st.integers(1, 64), st.integers(0, 2 ** 31))
def test_ddp_grad_sync_is_the_mean_and_power_of_two_worlds_commute_with_the_cast(world, n, seed):
    rng = np.random.default_rng(seed)
    grads = [rng.standard_normal(n).astype(np.float32) for _ in range(world)]
    mean = np.mean([g.astype(np.float64) for g in grads], axis=0)
    f32 = O.ddp_grad_sync(grads, "f32")[0]
    assert np.all(np.abs(f32 - mean) <= 1e-6 * np.mean([np.abs(g) for g in grads], axis=0) * world + 1e-12)
    bf = O.ddp_grad_sync(grads, "bf16")[0]
    assert np.all(np.abs(bf - mean) <= 2.0 ** -6 * (np.abs(mean) + np.mean([np.abs(g) for g in grads], axis=0)))

"""Pins the CPU oracle against fixtures generated FROM THE REFERENCE
(tests/golden/make_golden.py): the reference's own ``_apply_op`` and real gloo runs of
TorchGLOOGroup's call sequence."""
import numpy as np
import pytest

from oracle import collective_oracle as O

try:
    import ml_dtypes

    BF16 = np.dtype(ml_dtypes.bfloat16)
except Exception:  # pragma: no cover
    BF16 = None

DTYPES = ["uint8", "int8", "int32", "int64", "float16", "bfloat16", "float32", "float64"]
OPS = {"SUM": O.SUM, "PRODUCT": O.PRODUCT, "MIN": O.MIN, "MAX": O.MAX}
WORLDS = [2, 3, 4, 8]


def _view(a, dname):
    return a.view(BF16) if dname == "bfloat16" else a


@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("dname", DTYPES)
@pytest.mark.parametrize("oname", list(OPS))
def test_reduction_matches_reference_apply_op_bit_exact(golden, world, dname, oname):
    key = f"allreduce/w{world}/{dname}/{oname}"
    xs = [_view(x, dname) for x in golden[key + "/in"]]
    want = _view(golden[key + "/apply_op"], dname)
    got = O.reduce_rank_ascending(xs, OPS[oname])
    assert got.dtype == want.dtype
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), key


@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("dname", DTYPES)
@pytest.mark.parametrize("oname", list(OPS))
def test_allreduce_matches_gloo(golden, world, dname, oname):
    key = f"allreduce/w{world}/{dname}/{oname}"
    xs = [_view(x, dname).copy() for x in golden[key + "/in"]]
    gloo = [_view(g, dname) for g in golden[key + "/gloo"]]
    # every gloo rank holds the same result
    for g in gloo[1:]:
        assert np.array_equal(g.view(np.uint8), gloo[0].view(np.uint8))
    ins = [x.copy() for x in xs]
    O.allreduce(xs, OPS[oname])
    is_float = dname.startswith("float") or dname == "bfloat16"
    order_free = oname in ("MIN", "MAX") or not is_float or world == 2
    if order_free:
        # integers, min/max, and a single fp add are order independent: bit exact
        assert np.array_equal(xs[0].view(np.uint8), gloo[0].view(np.uint8)), key
    elif dname in ("float32", "float64") and oname == "SUM":
        # north_star tolerance: |out - ref| <= 1e-6 * sum_r |x_r|
        bound = 1e-6 * np.sum([np.abs(i.astype(np.float64)) for i in ins], axis=0)
        err = np.abs(xs[0].astype(np.float64) - gloo[0].astype(np.float64))
        assert np.all(err <= bound + 1e-300), (key, err.max())
    else:
        # products and 16-bit floats: gloo's ring order differs from rank-ascending
        rtol = {"float16": 2e-2, "bfloat16": 1e-1}.get(dname, 1e-5)
        np.testing.assert_allclose(xs[0].astype(np.float64), gloo[0].astype(np.float64), rtol=rtol, atol=rtol)
    for x in xs[1:]:
        assert np.array_equal(x.view(np.uint8), xs[0].view(np.uint8))


@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("dname", ["float32", "int32"])
def test_data_movement_and_rooted_ops_match_gloo(golden, world, dname):
    root = world - 1
    # broadcast
    xs = [x.copy() for x in golden[f"broadcast/w{world}/{dname}/in"]]
    O.broadcast(xs, root)
    assert np.array_equal(np.stack(xs), golden[f"broadcast/w{world}/{dname}/gloo"])
    # allgather: gloo fixture is [rank][p][numel]
    xs = [x.copy() for x in golden[f"allgather/w{world}/{dname}/in"]]
    outs = [[np.empty_like(xs[0]) for _ in range(world)] for _ in range(world)]
    O.allgather(outs, xs)
    assert np.array_equal(np.stack([np.stack(o) for o in outs]), golden[f"allgather/w{world}/{dname}/gloo"])
    # reduce: non-root ranks keep their input
    xs = [x.copy() for x in golden[f"reduce/w{world}/{dname}/in"]]
    O.reduce(xs, root, O.SUM)
    want = golden[f"reduce/w{world}/{dname}/gloo"]
    for r in range(world):
        if r != root or dname == "int32" or world == 2:
            assert np.array_equal(xs[r], want[r])
        else:
            np.testing.assert_allclose(xs[r], want[r], rtol=1e-5, atol=1e-5)
    # send/recv 0 -> world-1
    xs = [x.copy() for x in golden[f"sendrecv/w{world}/{dname}/in"]]
    O.sendrecv(xs[0], xs[world - 1])
    assert np.array_equal(np.stack(xs), golden[f"sendrecv/w{world}/{dname}/gloo"])
    # reducescatter: fixture input is [rank q][slot i][numel]
    lists = [[t.copy() for t in per_rank] for per_rank in golden[f"reducescatter/w{world}/{dname}/in"]]
    outs = [np.empty_like(lists[0][0]) for _ in range(world)]
    O.reducescatter(outs, lists, O.SUM)
    want = golden[f"reducescatter/w{world}/{dname}/gloo"]
    if dname == "int32" or world == 2:
        assert np.array_equal(np.stack(outs), want)
    else:
        np.testing.assert_allclose(np.stack(outs), want, rtol=1e-5, atol=1e-5)


def test_cgraph_layouts_and_op_numbering():
    xs = [np.arange(8, dtype=np.float32).reshape(4, 2) * (r + 1) for r in range(2)]
    cat = O.cgraph_allgather(xs)[0]
    assert cat.shape == (8, 2) and np.array_equal(cat[:4], xs[0]) and np.array_equal(cat[4:], xs[1])
    rs = O.cgraph_reducescatter(xs, 0)
    assert rs[0].shape == (2, 2) and np.array_equal(rs[1], (xs[0] + xs[1])[2:])
    # cgraph enum: MAX is 2, MIN is 3 (experimental/util/types.py:11-17)
    assert np.array_equal(O.cgraph_allreduce(xs, 2)[0], np.maximum(xs[0], xs[1]))
    assert np.array_equal(O.cgraph_allreduce(xs, 3)[0], np.minimum(xs[0], xs[1]))
    assert np.array_equal(O.cgraph_allreduce(xs, 4)[0], (xs[0] + xs[1]) / 2)
    with pytest.raises(ValueError):
        O.cgraph_reducescatter([np.zeros((3, 2), np.float32)] * 2, 0)


def test_ddp_grad_sync_matches_torch_hooks():
    """The gradient oracle against torch's own arithmetic: reducer mean and bf16_compress_hook."""
    import torch

    g = torch.Generator().manual_seed(7)
    n = 4
    grads = [torch.randn(1000, generator=g) for _ in range(n)]
    want = torch.stack([x / n for x in grads]).sum(0)  # order-free reference
    got = O.ddp_grad_sync([x.numpy() for x in grads], "f32")[0]
    np.testing.assert_allclose(got, want.numpy(), rtol=0, atol=1e-6)
    comp = [x.to(torch.bfloat16).div_(n) for x in grads]  # bf16_compress_hook: cast, then divide
    acc = comp[0].float()
    for c in comp[1:]:
        acc = acc + c.float()
    want16 = acc.to(torch.bfloat16).float()
    got16 = O.ddp_grad_sync([x.numpy() for x in grads], "bf16")[0]
    assert np.array_equal(got16, want16.numpy())

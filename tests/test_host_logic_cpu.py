"""More CPU tests of the host-side mirrors: declarative / env-var group creation
(collective.py:188-261,760-770), enum translations, operand checks that must fire
before anything reaches the GPU."""
import enum

import numpy as np
import pytest
import torch

from ray_b200 import _native as N
from ray_b200 import collective as col
from ray_b200.channel.communicator import _cgraph_op_code
from ray_b200.collective.b200_group import _as_cuda_tensor, _check_same_shape_dtype, _op_code, _unwrap_one
from ray_b200.store import DictStore


class Recorder(col.BaseGroup):
    calls = []

    @classmethod
    def backend(cls):
        return "REC"

    @classmethod
    def check_backend_availability(cls):
        return True

    def _log(name):  # noqa: N805
        def f(self, *a, **k):
            type(self).calls.append((name, self.rank, self.world_size, self.group_name))

        return f

    allreduce, barrier, reduce, allgather = _log("allreduce"), _log("barrier"), _log("reduce"), _log("allgather")
    broadcast, reducescatter, send, recv = _log("broadcast"), _log("reducescatter"), _log("send"), _log("recv")


@pytest.fixture()
def rec_backend():
    from ray_b200.collective.registry import _global_registry

    if not _global_registry.is_registered("REC"):
        col.register_collective_backend("rec", Recorder)
    Recorder.calls = []
    return Recorder


def test_declarative_group_creation_and_validation(rec_backend):
    store = DictStore()
    mgr_a, mgr_b = col.GroupManager(store), col.GroupManager(store)
    with pytest.raises(RuntimeError, match="Each actor should correspond to one rank"):
        col.create_collective_group(["a", "b"], 2, [0], backend="rec", group_name="g1", store=store)
    with pytest.raises(RuntimeError, match="permutation"):
        col.create_collective_group(["a", "b"], 2, [0, 2], backend="rec", group_name="g1", store=store)
    with pytest.raises(RuntimeError, match="greater than zero"):
        col.create_collective_group([], 0, [], backend="rec", group_name="g1", store=store)
    with pytest.raises(RuntimeError, match="not registered"):
        col.create_collective_group(["a"], 1, [0], backend="nope", group_name="g1", store=store)
    col.create_collective_group(["a", "b"], 2, [1, 0], backend="rec", group_name="g1", store=store)
    with pytest.raises(RuntimeError, match="twice"):
        col.create_collective_group(["a", "b"], 2, [1, 0], backend="rec", group_name="g1", store=store)
    with col.use_manager(mgr_a):
        col.set_member_id("a")
        col.allreduce(np.ones(2, np.float32), "g1")  # lazily created from the record: rank 1
        assert col.get_rank("g1") == 1 and col.get_collective_group_size("g1") == 2
    with col.use_manager(mgr_b):
        col.set_member_id("b")
        col.barrier("g1")
        assert col.get_rank("g1") == 0
        col.set_member_id("stranger")
        with pytest.raises(RuntimeError, match="not initialized"):
            col.barrier("unknown-group")
    assert [c[0] for c in rec_backend.calls] == ["allreduce", "barrier"]
    with col.use_manager(mgr_a):
        col.destroy_collective_group("g1")
        assert not col.is_group_initialized("g1")
        col.destroy_collective_group("g1")  # destroying twice only warns
    col.set_member_id(None)


def test_env_var_group_creation(rec_backend, monkeypatch):
    monkeypatch.setenv("collective_group_name", "envgroup")
    monkeypatch.setenv("collective_rank", "3")
    monkeypatch.setenv("collective_world_size", "4")
    monkeypatch.setenv("collective_backend", "rec")
    with col.use_manager(col.GroupManager(DictStore())):
        with pytest.raises(RuntimeError, match="not initialized"):
            col.barrier("othergroup")
        col.barrier("envgroup")
        assert col.get_rank("envgroup") == 3 and col.get_collective_group_size("envgroup") == 4
        with pytest.raises(ValueError, match="must be less than world size"):
            col.send(torch.ones(1), 4, "envgroup")
        with pytest.raises(ValueError, match="negative"):
            col.recv(torch.ones(1), -1, "envgroup")
        with pytest.raises(RuntimeError, match="is self"):
            col.send(torch.ones(1), 3, "envgroup")
        with pytest.raises(ValueError, match="needs to be a string"):
            col.init_collective_group(1, 0, backend="rec", group_name="")


def test_reduce_op_translations():
    assert [_op_code(o) for o in (col.ReduceOp.SUM, col.ReduceOp.PRODUCT, col.ReduceOp.MIN, col.ReduceOp.MAX)] == \
        [N.SUM, N.PROD, N.MIN, N.MAX]

    class ForeignCollectiveOp(enum.Enum):  # ray.util.collective.types.ReduceOp from a real Ray
        SUM = 0
        PRODUCT = 1
        MIN = 2
        MAX = 3

    assert _op_code(ForeignCollectiveOp.MIN) == N.MIN and _op_code(ForeignCollectiveOp.MAX) == N.MAX
    with pytest.raises(RuntimeError, match="Unsupported reduce op"):
        _op_code("sum")

    class CgraphOp(enum.Enum):  # ray.experimental.util.types.ReduceOp: MAX and MIN swapped (Q2)
        SUM = 0
        PRODUCT = 1
        MAX = 2
        MIN = 3
        AVG = 4

    assert [_cgraph_op_code(o) for o in CgraphOp] == [N.SUM, N.PROD, N.MAX, N.MIN, N.AVG]
    assert [_cgraph_op_code(v) for v in range(5)] == [N.SUM, N.PROD, N.MAX, N.MIN, N.AVG]  # raw nccl values
    with pytest.raises(ValueError):
        _cgraph_op_code(9)


def test_operand_checks_fire_before_the_gpu_is_touched():
    with pytest.raises(RuntimeError, match="must be on GPU"):
        _as_cuda_tensor(torch.ones(2))
    with pytest.raises(ValueError, match="Unsupported tensor type"):
        _as_cuda_tensor(np.ones(2))
    with pytest.raises(RuntimeError, match="1-element tensor list"):
        _unwrap_one([torch.ones(1), torch.ones(1)])
    with pytest.raises(RuntimeError, match="1-element tensor list"):
        _unwrap_one(torch.ones(1))
    a = torch.ones(2, 3)
    _check_same_shape_dtype(a, [torch.zeros(2, 3), torch.zeros(2, 3)])
    with pytest.raises(RuntimeError, match="same shape"):
        _check_same_shape_dtype(a, [torch.zeros(3, 2)])
    with pytest.raises(RuntimeError, match="same dtype"):
        _check_same_shape_dtype(a, [torch.zeros(2, 3, dtype=torch.float64)])
    assert col.B200Group.backend() == "B200"


def test_rdt_heap_arena_first_fit_and_coalescing():
    from ray_b200.rdt import _HeapArena

    a = _HeapArena(4096, 1 << 20)
    x, y, z = a.alloc(1000), a.alloc(70000), a.alloc(1)
    assert x == 4096 and y == 4096 + 1024 and z % 256 == 0 and len({x, y, z}) == 3
    a.free(y)
    assert a.alloc(60000) == y  # first fit reuses the hole
    a.free(x)
    a.free(z)
    a.free(y)
    assert a._free == [(4096, 1 << 20)]  # everything coalesced back
    import pytest

    with pytest.raises(MemoryError):
        a.alloc((1 << 20) + 1)


def test_stream_switch_between_collectives_is_ordered_on_the_device(monkeypatch):
    """two collectives of one communicator issued from different streams must not
    race on the device-resident launch counter: the second stream waits for the tail of the first."""
    import torch

    from ray_b200.comm import B200Comm

    log = []

    class FakeStream:
        def __init__(self, h):
            self.cuda_stream = h

        def wait_event(self, ev):
            log.append(("wait", self.cuda_stream, ev.recorded_on))

    class FakeEvent:
        def record(self, stream):
            self.recorded_on = stream.cuda_stream

    cur = {"s": FakeStream(11)}
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: cur["s"])
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    comm = object.__new__(B200Comm)
    comm.device = 0
    assert comm._stream() == 11 and log == []          # first op: nothing to order against
    assert comm._stream() == 11 and log == []          # same stream: free
    cur["s"] = FakeStream(22)
    assert comm._stream() == 22 and log == [("wait", 22, 11)]   # switch: 22 waits for the tail of 11
    assert comm._stream(FakeStream(22)) == 22 and len(log) == 1
    assert comm._stream(FakeStream(33)) == 33 and log[-1] == ("wait", 33, 22)
    comm._closed = True  # keep __del__ quiet

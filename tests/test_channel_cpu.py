"""CPU tests of the Compiled-Graph boundary's host logic: communicator bookkeeping before
``initialize``, pickling, metadata header codec, collective-operation shape / dtype rules
(driven with a CPU stand-in communicator that follows CPUCommBarrier semantics,
python/ray/experimental/channel/cpu_communicator.py:17-89)."""
import pickle

import pytest
import torch

from ray_b200.channel import (AllGatherOp, AllReduceOp, B200Communicator, Communicator, RayChannelError,
                              ReduceScatterOp, execute_collective)
from ray_b200.channel import tensor_channel as tc


def test_communicator_bookkeeping_and_pickle_without_gpu():
    handles = ["a", "b", "c"]
    c = B200Communicator(3, actor_handles=handles)
    assert isinstance(c, Communicator)
    assert c.get_world_size() == 3 and c.get_actor_handles() == handles
    assert [c.get_rank(h) for h in handles] == [0, 1, 2] and c.get_self_rank() is None
    with pytest.raises(ValueError, match="not in the B200 group"):
        c.get_rank("zzz")
    assert c.get_transport_name() == "accelerator"
    d = pickle.loads(pickle.dumps(c))
    assert d._comm_id == c._comm_id and d.get_world_size() == 3 and d._comm is None
    ids = {B200Communicator.generate_communicator_id() for _ in range(4)}
    assert len(ids) == 4
    with pytest.raises(RayChannelError):
        c.send(torch.ones(1), 1)  # not initialised -> closed semantics
    with pytest.raises(ValueError):
        c.initialize(7)
    c.destroy()
    c.destroy()  # idempotent


@pytest.mark.parametrize("shape,dtype", [((), torch.float32), ((7,), torch.int64), ((2, 3, 5), torch.bfloat16),
                                         ((1, 1, 1, 1, 9), torch.uint8)])
def test_metadata_header_roundtrip(shape, dtype):
    t = torch.zeros(shape, dtype=dtype)
    raw = tc._encode(t)
    assert len(raw) == tc._DESC_BYTES
    got_shape, got_dtype = tc._decode(raw)
    assert tuple(got_shape) == tuple(shape) and got_dtype == dtype
    with pytest.raises(RayChannelError):
        tc._decode(b"\\0" * tc._HEADER_BYTES)
    with pytest.raises(ValueError):
        tc._encode(torch.zeros([1] * 13))


class _CpuComm:
    """Single-process stand-in: world_size identical ranks (rank-ascending reduce of copies)."""

    def __init__(self, world):
        self.world = world
        self.multi_calls = 0

    def get_world_size(self):
        return self.world

    def allreduce(self, send, recv, op):
        recv.copy_(send * self.world)

    def allgather(self, send, recv):
        recv.copy_(torch.cat([send] * self.world, dim=0))

    def reducescatter(self, send, recv, op):
        recv.copy_(send[: send.shape[0] // self.world] * self.world)

    def allreduce_multi(self, tensors, op):
        self.multi_calls += 1
        for t in tensors:
            t.mul_(self.world)


def test_collective_operation_shapes_and_errors():
    comm = _CpuComm(4)
    x = torch.arange(24, dtype=torch.float32).reshape(8, 3)
    assert execute_collective(comm, AllGatherOp(), x).shape == (32, 3)
    assert torch.equal(execute_collective(comm, AllReduceOp(), x), x * 4)
    assert execute_collective(comm, ReduceScatterOp(), x).shape == (2, 3)
    with pytest.raises(ValueError, match="divisible"):
        execute_collective(comm, ReduceScatterOp(), torch.zeros(6, 3))
    a, b = torch.ones(5), torch.ones(2, 2)
    outs = execute_collective(comm, AllReduceOp(), a, b)
    assert comm.multi_calls == 1 and isinstance(outs, tuple)
    assert torch.all(outs[0] == 4) and outs[1].shape == (2, 2) and torch.all(a == 1)  # inputs untouched
    with pytest.raises(ValueError, match="same dtype"):
        execute_collective(comm, AllReduceOp(), a, b.half())
    with pytest.raises(ValueError, match="torch tensor"):
        execute_collective(comm, AllReduceOp(), [1, 2])
    with pytest.raises(ValueError, match="unsupported"):
        execute_collective(comm, object(), a)

    class Foreign(_CpuComm):  # a Communicator without the multi-tensor entry: flatten path
        allreduce_multi = None

    outs = execute_collective(Foreign(2), AllReduceOp(), a, b)
    assert torch.all(outs[0] == 2) and torch.all(outs[1] == 2)


def test_rdt_metadata_without_gpu():
    from ray_b200.rdt import B200TensorTransport

    tr = B200TensorTransport()
    meta = tr.extract_tensor_transport_metadata("o", [torch.ones(2, 3), torch.ones(4, dtype=torch.int32)])
    assert [tuple(s) for s, _ in meta.tensor_meta] == [(2, 3), (4,)] and meta.tensor_device == "cpu"
    assert tr.extract_tensor_transport_metadata("o", []).tensor_device is None
    with pytest.raises(ValueError, match="No communicators"):
        tr.get_communicator_metadata("a", "b")

"""The Compiled-Graph collective operation: allocate the output, call the communicator.

Restates ``_CollectiveOperation.execute`` (python/ray/dag/collective_node.py:176-248), the
COMPUTE step a ``ray.experimental.collective.{allreduce,allgather,reducescatter}.bind(...)``
node runs inside each actor's execution loop:

* all-gather  -> ``[d0 * n, ...]`` output (:198-206)
* all-reduce  -> ``empty_like`` for one tensor (:207-211); several tensors must share a dtype
  (``ValueError``) and are reduced as one message (:212-232).  The reference flattens them with
  ``parameters_to_vector`` (a device cat kernel) and returns views of the flat buffer; here the
  list goes to the single-launch multi-tensor kernel and the results are written
  straight into freshly allocated outputs, so no flatten / unflatten copies exist.
* reduce-scatter -> ``[d0 / n, ...]`` output, ``d0 % n == 0`` required (``ValueError``) (:233-247)
"""
from __future__ import annotations

from typing import Tuple, Union

import torch

from .communicator import Communicator, _cgraph_op_code


class AllGatherOp:
    """ray.experimental.util.types.AllGatherOp"""


class AllReduceOp:
    """ray.experimental.util.types.AllReduceOp(reduceOp=ReduceOp.SUM)"""

    def __init__(self, reduceOp=0):  # noqa: N803 - reference spelling
        self.reduceOp = reduceOp


class ReduceScatterOp:
    """ray.experimental.util.types.ReduceScatterOp(reduceOp=ReduceOp.SUM)"""

    def __init__(self, reduceOp=0):  # noqa: N803
        self.reduceOp = reduceOp


def _kind(op) -> str:
    name = type(op).__name__
    if name in ("AllGatherOp", "AllReduceOp", "ReduceScatterOp"):
        return name
    raise ValueError(f"unsupported collective operation {op!r}")


def execute_collective(communicator: Communicator, op, *send_buf: torch.Tensor
                       ) -> Union[torch.Tensor, Tuple[torch.Tensor, ...]]:
    """Run ``op`` on ``send_buf`` through ``communicator``; outputs are allocated and returned."""
    if not all(isinstance(t, torch.Tensor) for t in send_buf):
        raise ValueError("Expected a torch tensor for each input node")
    world = communicator.get_world_size()
    kind = _kind(op)
    if kind == "AllGatherOp":
        assert len(send_buf) == 1
        t = send_buf[0]
        recv = torch.empty((t.shape[0] * world, *t.shape[1:]), dtype=t.dtype, device=t.device)
        communicator.allgather(t, recv)
        return recv
    if kind == "AllReduceOp":
        if len(send_buf) == 1:
            t = send_buf[0]
            recv = torch.empty_like(t)
            communicator.allreduce(t, recv, op.reduceOp)
            return recv
        if not all(t.dtype == send_buf[0].dtype for t in send_buf):
            raise ValueError("Expected all input tensors to have the same dtype, "
                             f"but got {[t.dtype for t in send_buf]}")
        outs = tuple(t.clone() for t in send_buf)
        multi = getattr(communicator, "allreduce_multi", None)
        if multi is not None:
            multi(list(outs), op.reduceOp)
        else:  # a foreign Communicator: the reference's flatten path
            flat = torch.nn.utils.parameters_to_vector(outs)
            communicator.allreduce(flat, flat, op.reduceOp)
            off = 0
            for o in outs:
                o.copy_(flat[off:off + o.numel()].view(o.shape))
                off += o.numel()
        return outs
    t = send_buf[0]
    assert len(send_buf) == 1
    if t.shape[0] % world != 0:
        raise ValueError("Expected the first dimension of the input tensor to be divisible "
                         f"by the world size {world}")
    recv = torch.empty((t.shape[0] // world, *t.shape[1:]), dtype=t.dtype, device=t.device)
    communicator.reducescatter(t, recv, op.reduceOp)
    return recv


__all__ = ["AllGatherOp", "AllReduceOp", "ReduceScatterOp", "execute_collective", "_cgraph_op_code"]

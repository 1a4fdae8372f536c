"""``GlooOracleGroup``: the reference's CPU backend restated as a ``BaseGroup``.
TEST INFRASTRUCTURE ONLY (see collective_oracle.py) -- never registered by the product.

Restates ``TorchGLOOGroup`` (python/ray/util/collective/collective_group/
torch_gloo_collective_group.py:58-290): one default gloo process group per process,
MASTER address published by rank 0 through the rendezvous store (the reference uses the GCS
internal KV, :128-150), operand unwrapping that turns numpy arrays into zero-copy torch views
(:172-184) so results land in the caller's array, reduce-on-a-clone for non-root ranks
(:229-240), reduce-scatter emulated by all-reducing every list member (:260-282).

It serves BASELINE config 1 ("ray.util.collective.allreduce fp32 world_size=2 gloo backend on
CPU") and lets the CPU test-suite drive ``ray_b200.collective``'s host logic with two real
processes.
"""
from __future__ import annotations

import datetime
import socket
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ray_b200.collective import types
from ray_b200.collective.base_group import BaseGroup
from ray_b200.store import Store, default_store

from .gloo_reference import TORCH_REDUCE_OP


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class GlooOracleGroup(BaseGroup):
    store: Optional[Store] = None

    def __init__(self, world_size, rank, group_name, gloo_timeout: Optional[int] = None):
        super().__init__(world_size, rank, group_name)
        st = type(self).store or default_store()
        key = f"gloo_oracle/{group_name}/master"
        if rank == 0:
            addr = f"127.0.0.1:{_free_port()}"
            st.set(key, addr.encode())
        else:
            addr = st.get(key, timeout_s=(gloo_timeout or 30000) / 1000.0).decode()
        host, port = addr.split(":")
        self._owns_default = not dist.is_initialized()
        if self._owns_default:
            dist.init_process_group(
                "gloo", init_method=f"tcp://{host}:{port}", rank=rank, world_size=world_size,
                timeout=datetime.timedelta(milliseconds=gloo_timeout or 30000))
            self._pg = dist.group.WORLD
        else:
            self._pg = dist.new_group(list(range(world_size)), backend="gloo")
        if rank == 0:
            dist.barrier(group=self._pg)
            st.delete(key)
        else:
            dist.barrier(group=self._pg)

    @classmethod
    def backend(cls):
        return types.Backend.GLOO

    @classmethod
    def check_backend_availability(cls) -> bool:
        return dist.is_available() and dist.is_gloo_available()

    def destroy_group(self):
        if self._owns_default:
            dist.destroy_process_group()
        elif self._pg is not None:
            dist.destroy_process_group(self._pg)
        self._pg = None

    @staticmethod
    def _one(wrapped) -> torch.Tensor:
        assert isinstance(wrapped, list) and len(wrapped) == 1
        t = wrapped[0]
        if isinstance(t, torch.Tensor):
            return t
        if isinstance(t, np.ndarray):
            return torch.from_numpy(t)
        raise ValueError(f"torch_gloo group only accepts torch.Tensor or numpy.ndarray, received {type(t)}")

    @classmethod
    def _many(cls, wrapped) -> List[torch.Tensor]:
        assert isinstance(wrapped, list) and len(wrapped) == 1
        return [cls._one([t]) for t in wrapped[0]]

    @staticmethod
    def _op(opts) -> "dist.ReduceOp":
        return TORCH_REDUCE_OP[opts.reduceOp.value]

    def allreduce(self, tensor, allreduce_options=None):
        opts = allreduce_options or types.AllReduceOptions()
        dist.all_reduce(self._one(tensor), op=self._op(opts), group=self._pg)

    def barrier(self, barrier_options=None):
        dist.barrier(group=self._pg)

    def reduce(self, tensor, reduce_options=None):
        opts = reduce_options or types.ReduceOptions()
        t = self._one(tensor)
        if self._rank != opts.root_rank:
            t = t.detach().clone()
        dist.reduce(t, dst=opts.root_rank, op=self._op(opts), group=self._pg)

    def allgather(self, tensor_list, tensor, allgather_options=None):
        dist.all_gather(self._many(tensor_list), self._one(tensor), group=self._pg)

    def broadcast(self, tensor, broadcast_options=None):
        opts = broadcast_options or types.BroadcastOptions()
        dist.broadcast(self._one(tensor), src=opts.root_rank, group=self._pg)

    def reducescatter(self, tensor, tensor_list, reducescatter_options=None):
        opts = reducescatter_options or types.ReduceScatterOptions()
        ins = self._many(tensor_list)
        out = self._one(tensor)
        if out.shape != ins[self._rank].shape:
            raise ValueError(f"Output tensor has wrong shape {out.shape}, expected {ins[self._rank].shape}")
        for t in ins:
            dist.all_reduce(t, op=self._op(opts), group=self._pg)
        if out.data_ptr() != ins[self._rank].data_ptr():
            out.copy_(ins[self._rank])

    def send(self, tensor, send_options):
        dist.send(self._one(tensor), dst=send_options.dst_rank)

    def recv(self, tensor, recv_options):
        dist.recv(self._one(tensor), src=recv_options.src_rank)

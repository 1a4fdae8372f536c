"""The backend contract (boundary B1).

When Ray is installed the real ``ray.util.collective.collective_group.base_collective_group.
BaseGroup`` is used, so ``B200Group`` can be handed to ``register_collective_backend``
unchanged.  Without Ray (this build container, the GPU box) an equivalent ABC with the same
constructor, properties and abstract methods is defined here
(python/ray/util/collective/collective_group/base_collective_group.py:16-91).
"""
from __future__ import annotations

import abc

try:  # pragma: no cover - Ray is not installable in the build environment
    from ray.util.collective.collective_group.base_collective_group import BaseGroup  # type: ignore

    HAVE_RAY_BASEGROUP = True
except Exception:  # ModuleNotFoundError, or ray._raylet missing
    HAVE_RAY_BASEGROUP = False

    class BaseGroup(abc.ABC):
        """One process's membership of a collective group."""

        def __init__(self, world_size: int, rank: int, group_name: str):
            self._world_size, self._rank, self._group_name = world_size, rank, group_name

        rank = property(lambda self: self._rank, doc="rank of this process")
        world_size = property(lambda self: self._world_size, doc="number of processes in the group")
        group_name = property(lambda self: self._group_name, doc="name of the group")

        def destroy_group(self):
            """Release communicator resources."""

        @classmethod
        def backend(cls):
            raise NotImplementedError()

        @classmethod
        @abc.abstractmethod
        def check_backend_availability(cls) -> bool: ...

        # Operands arrive list-wrapped exactly as ray.util.collective passes them
        # (collective.py:331,392,450,503,557,607,670).
        @abc.abstractmethod
        def allreduce(self, tensor, allreduce_options=None): ...

        @abc.abstractmethod
        def barrier(self, barrier_options=None): ...

        @abc.abstractmethod
        def reduce(self, tensor, reduce_options=None): ...

        @abc.abstractmethod
        def allgather(self, tensor_list, tensor, allgather_options=None): ...

        @abc.abstractmethod
        def broadcast(self, tensor, broadcast_options=None): ...

        @abc.abstractmethod
        def reducescatter(self, tensor, tensor_list, reducescatter_options=None): ...

        @abc.abstractmethod
        def send(self, tensor, send_options): ...

        @abc.abstractmethod
        def recv(self, tensor, recv_options): ...

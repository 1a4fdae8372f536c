"""RDT tensor transport (boundary B3): ``@ray.method(tensor_transport="B200")``.

Implements ``TensorTransportManager`` (python/ray/experimental/rdt/tensor_transport_manager.py:
37-224) the way the reference's ``CollectiveTensorTransport`` does for NCCL / GLOO
(python/ray/experimental/rdt/collective_tensor_transport.py:34-203): a two-sided transport whose
``__ray_send__`` / ``__ray_recv__`` halves (run on the ``_ray_system`` concurrency-group thread,
rdt_manager.py:655-681) map to ``collective.send`` / ``collective.recv`` of a collective group
that contains both actors -- here a B200 group, so the payload moves through the
sender-push NVLink kernel.  Two differences from the NCCL transport:

* ``can_abort_transport()`` is True: the device-side waits poll an abort word, so a stuck
  transfer is cancelled instead of Ray having to kill both actors (tensor_transport_manager.py:
  75-92);
* sends are eager up to the inbox ring size, so the sender does not block on the receiver
  having posted its recv.

Register with ``register_tensor_transport("B200", ["cuda"], B200TensorTransport, torch.Tensor)``
(python/ray/experimental/rdt/util.py:46-84).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional, Tuple

import torch

try:  # pragma: no cover - Ray is not installable in the build environment
    import ray.util.collective as _collective  # type: ignore
    from ray.experimental.rdt.tensor_transport_manager import (  # type: ignore
        CommunicatorMetadata,
        TensorTransportManager,
        TensorTransportMetadata,
    )

    HAVE_RAY_RDT = True
except Exception:
    import abc

    from . import collective as _collective

    HAVE_RAY_RDT = False

    @dataclass
    class CommunicatorMetadata:
        """Metadata for the communicator."""

    @dataclass
    class TensorTransportMetadata:
        """(shape, dtype) per tensor plus the common device type."""

        tensor_meta: List[Tuple[Any, Any]] = field(default_factory=list)
        tensor_device: Optional[str] = None

    class TensorTransportManager(abc.ABC):
        @abc.abstractmethod
        def tensor_transport_backend(self) -> str: ...

        @staticmethod
        @abc.abstractmethod
        def is_one_sided() -> bool: ...

        @staticmethod
        @abc.abstractmethod
        def can_abort_transport() -> bool: ...

        @abc.abstractmethod
        def actor_has_tensor_transport(self, actor) -> bool: ...

        @abc.abstractmethod
        def extract_tensor_transport_metadata(self, obj_id, rdt_object): ...

        @abc.abstractmethod
        def get_communicator_metadata(self, src_actor, dst_actor, backend=None): ...

        @abc.abstractmethod
        def recv_multiple_tensors(self, obj_id, tensor_transport_metadata, communicator_metadata,
                                  target_buffers=None): ...

        @abc.abstractmethod
        def send_multiple_tensors(self, tensors, tensor_transport_metadata, communicator_metadata): ...

        @abc.abstractmethod
        def garbage_collect(self, obj_id, tensor_transport_meta, tensors): ...

        @abc.abstractmethod
        def abort_transport(self, obj_id, communicator_metadata): ...


@dataclass
class B200CommunicatorMetadata(CommunicatorMetadata):
    """Which group and which ranks a transfer uses (collective_tensor_transport.py:19-31)."""

    communicator_name: str = ""
    src_rank: Optional[int] = None
    dst_rank: Optional[int] = None


@dataclass
class B200TransportMetadata(TensorTransportMetadata):
    pass


class B200TensorTransport(TensorTransportManager):
    """Two-sided RDT transport over a B200 collective group."""

    #: resolves (src_actor, dst_actor) -> (group name, src rank, dst rank) on the driver; with Ray
    #: this is ray.experimental.collective.get_collective_groups, injected here for harnesses
    group_resolver = None

    def tensor_transport_backend(self) -> str:
        return "B200"

    @staticmethod
    def is_one_sided() -> bool:
        return False

    @staticmethod
    def can_abort_transport() -> bool:
        return True

    def actor_has_tensor_transport(self, actor) -> bool:
        if HAVE_RAY_RDT:  # pragma: no cover
            from ray.experimental.collective import get_collective_groups

            return len(get_collective_groups([actor], backend=self.tensor_transport_backend())) > 0
        return self.group_resolver is not None

    def extract_tensor_transport_metadata(self, obj_id: str, rdt_object: List[torch.Tensor]) -> B200TransportMetadata:
        meta, device = [], None
        for t in rdt_object or []:
            device = device or t.device
            if t.device.type != device.type:
                raise ValueError("All tensors in an RDT object must have the same device type.")
            meta.append((t.shape, t.dtype))
        return B200TransportMetadata(tensor_meta=meta, tensor_device=device.type if device else None)

    def get_communicator_metadata(self, src_actor, dst_actor, backend: Optional[str] = None) -> B200CommunicatorMetadata:
        if HAVE_RAY_RDT:  # pragma: no cover
            from ray.experimental.collective import get_collective_groups

            groups = get_collective_groups([src_actor, dst_actor], backend=backend)
            if len(groups) == 0:
                raise ValueError(f"No communicators found for actors {src_actor} and {dst_actor}. Create a "
                                 "communicator with `ray.experimental.collective.create_collective_group` before "
                                 "calling actor tasks. with non-default tensor_transport.")
            if len(groups) > 1:
                raise ValueError(f"There are {len(groups)} possible communicators that contain actors {src_actor} "
                                 f"and {dst_actor}. Currently, RDT objects only support one communicator.")
            g = groups[0]
            name, src, dst = g.name, g.get_rank(src_actor), g.get_rank(dst_actor)
        else:
            if self.group_resolver is None:
                raise ValueError(f"No communicators found for actors {src_actor} and {dst_actor}.")
            name, src, dst = self.group_resolver(src_actor, dst_actor)
        if src == -1 or dst == -1:
            raise ValueError("Sender and receiver must be in the same communicator.")
        return B200CommunicatorMetadata(communicator_name=name, src_rank=src, dst_rank=dst)

    def recv_multiple_tensors(self, obj_id, tensor_transport_metadata, communicator_metadata,
                              target_buffers: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        assert isinstance(communicator_metadata, B200CommunicatorMetadata)
        tensors = target_buffers or [torch.empty(tuple(shape), dtype=dtype, device="cuda")
                                     for shape, dtype in tensor_transport_metadata.tensor_meta]
        for t in tensors:
            _collective.recv(t, communicator_metadata.src_rank, communicator_metadata.communicator_name)
        return tensors

    def send_multiple_tensors(self, tensors, tensor_transport_metadata, communicator_metadata) -> None:
        assert isinstance(communicator_metadata, B200CommunicatorMetadata)
        device = tensors[0].device if tensors else None
        for t in tensors:
            if t.device.type != device.type:
                raise ValueError(f"tensor device {t.device} does not match device {device}")
            _collective.send(t, communicator_metadata.dst_rank, communicator_metadata.communicator_name)

    def garbage_collect(self, obj_id, tensor_transport_meta, tensors) -> None:
        """Nothing is registered per object: the inbox ring is owned by the communicator."""

    def abort_transport(self, obj_id, communicator_metadata) -> None:
        group = _collective.get_group_handle(communicator_metadata.communicator_name)
        comm = getattr(group, "comm", None)
        if comm is not None:
            comm.abort()

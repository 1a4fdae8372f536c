"""``ray_b200.channel`` -- the Compiled-Graph communicator and GPU tensor channel
(python/ray/experimental/channel/__init__.py:1-43)."""
from .communicator import B200Communicator, Communicator, RayChannelError
from .collective_op import AllGatherOp, AllReduceOp, ReduceScatterOp, execute_collective
from .tensor_channel import TorchTensorAcceleratorChannel

__all__ = ["B200Communicator", "Communicator", "RayChannelError", "TorchTensorAcceleratorChannel",
           "AllGatherOp", "AllReduceOp", "ReduceScatterOp", "execute_collective"]

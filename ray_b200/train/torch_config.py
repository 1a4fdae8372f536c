"""``B200TorchConfig``: selects the b200 c10d backend for TorchTrainer / LearnerGroup.

Reference entry points (python/ray/train/torch/config.py): ``TorchConfig`` dataclass
(:42-83: ``backend``, ``init_method`` "env"|"tcp", ``timeout_s``), ``_TorchBackend.on_start``
(:186-233: choose nccl for GPU workers else gloo, publish MASTER_ADDR/PORT, run
``_setup_torch_process_group`` on every worker) and ``_setup_torch_process_group`` (:95-150).
RLlib reuses the same classes (rllib/core/learner/learner_group.py:57-76).

With Ray installed::

    trainer = TorchTrainer(loop, scaling_config=ScalingConfig(num_workers=8, use_gpu=True),
                           torch_config=B200TorchConfig())

``B200TorchConfig`` then *is* a ``ray.train.torch.TorchConfig`` whose backend class registers
the b200 backend on every worker before the stock ``on_start`` runs.  Without Ray (this
repository's harness, torchrun) the same dataclass drives ``setup_torch_process_group``,
the restatement of ``_setup_torch_process_group``.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass
from datetime import timedelta
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist

from .process_group import BACKEND_NAME, register_b200_backend

logger = logging.getLogger(__name__)

#: ranks of one b200 group: the GPUs of one HGX host (B200_MAX_RANKS in include/b200_collective.h)
MAX_B200_WORLD = 8

#: CUDA tensors -> hand-written kernels; CPU tensors (object collectives) -> gloo
DEFAULT_GPU_BACKEND = f"cpu:gloo,cuda:{BACKEND_NAME}"


def uses_b200(backend: str) -> bool:
    """Same containment rule the reference applies to nccl (config.py:86-92)."""
    return backend == BACKEND_NAME or any(
        item.split(":")[1] == BACKEND_NAME for item in backend.split(",") if item.startswith("cuda:"))


def resolve_backend(backend: Optional[str], use_gpu: bool) -> str:
    """``None`` -> b200 when the workers have GPUs, else gloo (config.py:189-196 picks nccl/gloo)."""
    if backend is not None:
        return backend
    return DEFAULT_GPU_BACKEND if use_gpu else "gloo"


def setup_torch_process_group(backend: str, world_rank: int, world_size: int, init_method: str = "env://",
                              timeout_s: int = 1800) -> None:
    """Connect this worker's default process group (config.py:95-150)."""
    level = logging.INFO if world_rank == 0 else logging.DEBUG
    logger.log(level, "Setting up process group for: %s [rank=%d, world_size=%d] using %s", init_method,
               world_rank, world_size, backend)
    if uses_b200(backend) and world_size > MAX_B200_WORLD:
        # one NVSwitch domain of one host: beyond that the reference's own default applies
        # (config.py:189-196 picks nccl for GPU workers)
        logger.warning("b200 groups span at most %d ranks on one host; %d workers requested -> falling back to "
                       "the nccl backend", MAX_B200_WORLD, world_size)
        backend = "nccl"
    if uses_b200(backend):
        register_b200_backend()
    dist.init_process_group(backend=backend, init_method=init_method, rank=world_rank, world_size=world_size,
                            timeout=timedelta(seconds=timeout_s))


def shutdown_torch(destroy_process_group: bool = False) -> None:
    """config.py:153-164."""
    if destroy_process_group and dist.is_initialized():
        dist.destroy_process_group()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


try:  # pragma: no cover - Ray is not installable in the build environment
    from ray.train.torch.config import TorchConfig as _RayTorchConfig
    from ray.train.torch.config import _TorchBackend as _RayTorchBackend

    class _B200TorchBackend(_RayTorchBackend):
        """Registers the b200 c10d backend on every worker, then defers to Ray's own on_start."""

        def on_start(self, worker_group, backend_config):
            worker_group.execute(register_b200_backend)
            if backend_config.backend is None:
                gpus = worker_group.get_resources_per_worker().get("GPU", 0)
                backend_config = type(backend_config)(
                    backend=resolve_backend(None, gpus > 0), init_method=backend_config.init_method,
                    timeout_s=backend_config.timeout_s)
            super().on_start(worker_group, backend_config)

    @dataclass
    class B200TorchConfig(_RayTorchConfig):
        @property
        def backend_cls(self):
            return _B200TorchBackend

    HAVE_RAY_TRAIN = True
except Exception:
    HAVE_RAY_TRAIN = False

    @dataclass
    class B200TorchConfig:
        """Field-compatible with ``ray.train.torch.TorchConfig`` (config.py:42-83)."""

        backend: Optional[str] = None
        init_method: str = "env"
        timeout_s: int = 1800

        def to_dict(self) -> Dict[str, Any]:
            return {"backend": self.backend, "init_method": self.init_method, "timeout_s": self.timeout_s}

        def init_url(self, master_addr: str, master_port: int) -> str:
            if self.init_method == "env":
                os.environ["MASTER_ADDR"] = master_addr
                os.environ["MASTER_PORT"] = str(master_port)
                return "env://"
            if self.init_method == "tcp":
                return f"tcp://{master_addr}:{master_port}"
            raise ValueError(f"The provided init_method ({self.init_method}) is not supported. Must be either "
                             "'env' or 'tcp'.")

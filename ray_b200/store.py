"""Rendezvous stores: where each rank publishes its 256-byte bootstrap handle.

The reference exchanges its ncclUniqueId through Ray's control plane -- a detached named
actor (util/collective/collective_group/nccl_collective_group.py:36-125, util/collective/
util.py:10-52), the GCS internal KV (torch_gloo_collective_group.py:128-150) or an
``__ray_call__`` (experimental/channel/torch_tensor_accelerator_channel.py:794-831).
This package needs exactly the same service -- ``set(key, bytes)`` and a blocking
``get(key)`` -- so the store is a small interface with one adapter per control plane.
"""
from __future__ import annotations

import os
import tempfile
import threading
import time
from pathlib import Path
from typing import Dict, Optional


class Store:
    """Minimal key/value rendezvous interface."""

    def set(self, key: str, value: bytes) -> None:  # pragma: no cover - interface
        raise NotImplementedError

    def get(self, key: str, timeout_s: float = 180.0) -> bytes:  # pragma: no cover - interface
        raise NotImplementedError

    def delete(self, key: str) -> None:
        pass


class DictStore(Store):
    """In-process store (several ranks as threads of one process: the single-GPU harness)."""

    def __init__(self):
        self._d: Dict[str, bytes] = {}
        self._cv = threading.Condition()

    def set(self, key, value):
        with self._cv:
            self._d[key] = bytes(value)
            self._cv.notify_all()

    def get(self, key, timeout_s=180.0):
        deadline = time.monotonic() + timeout_s
        with self._cv:
            while key not in self._d:
                left = deadline - time.monotonic()
                if left <= 0:
                    raise TimeoutError(f"rendezvous key {key!r} never appeared")
                self._cv.wait(left)
            return self._d[key]

    def delete(self, key):
        with self._cv:
            self._d.pop(key, None)


class FileStore(Store):
    """Directory-backed store for independent processes on one host."""

    def __init__(self, path: Optional[str] = None):
        self._dir = Path(path or tempfile.mkdtemp(prefix="b200_store_"))
        self._dir.mkdir(parents=True, exist_ok=True)

    @property
    def path(self) -> str:
        return str(self._dir)

    def _file(self, key: str) -> Path:
        safe = "".join(ch if ch.isalnum() or ch in "-_." else f"%{ord(ch):02x}" for ch in key)
        return self._dir / safe

    def set(self, key, value):
        f = self._file(key)
        tmp = f.with_name(f.name + f".tmp{os.getpid()}")
        tmp.write_bytes(bytes(value))
        os.replace(tmp, f)

    def get(self, key, timeout_s=180.0):
        f = self._file(key)
        deadline = time.monotonic() + timeout_s
        while True:
            try:
                return f.read_bytes()
            except FileNotFoundError:
                if time.monotonic() > deadline:
                    raise TimeoutError(f"rendezvous key {key!r} never appeared in {self._dir}")
                time.sleep(0.005)

    def delete(self, key):
        try:
            self._file(key).unlink()
        except FileNotFoundError:
            pass


class TorchDistStore(Store):
    """Adapter over a ``torch.distributed`` Store (TCPStore / FileStore / PrefixStore),
    e.g. the one behind the default process group of a torchrun / TorchTrainer worker."""

    def __init__(self, store):
        self._s = store

    def set(self, key, value):
        self._s.set(key, bytes(value))

    def get(self, key, timeout_s=180.0):
        import datetime

        try:
            self._s.wait([key], datetime.timedelta(seconds=max(timeout_s, 0.001)))
        except Exception as exc:  # torch raises DistStoreError / RuntimeError on expiry
            raise TimeoutError(f"rendezvous key {key!r} never appeared in the torch store") from exc
        return bytes(self._s.get(key))

    def delete(self, key):
        try:
            self._s.delete_key(key)
        except Exception:
            pass


class RayInternalKVStore(Store):
    """Adapter over Ray's GCS internal KV -- the same channel TorchGLOOGroup uses for its
    rendezvous (util/collective/collective_group/torch_gloo_collective_group.py:128-150).
    Only importable inside a Ray worker."""

    def __init__(self, namespace: str = "b200_collective"):
        from ray.experimental import internal_kv  # noqa: WPS433 - optional dependency

        self._kv = internal_kv
        self._ns = namespace.encode()

    def set(self, key, value):
        self._kv._internal_kv_put(key.encode(), bytes(value), overwrite=True, namespace=self._ns)

    def get(self, key, timeout_s=180.0):
        deadline = time.monotonic() + timeout_s
        while True:
            v = self._kv._internal_kv_get(key.encode(), namespace=self._ns)
            if v is not None:
                return bytes(v)
            if time.monotonic() > deadline:
                raise TimeoutError(f"rendezvous key {key!r} never appeared in the GCS KV")
            time.sleep(0.05)

    def delete(self, key):
        self._kv._internal_kv_del(key.encode(), namespace=self._ns)


_default_store: Optional[Store] = None
_default_lock = threading.Lock()


def set_default_store(store: Optional[Store]) -> None:
    """Install the store used when a group is created without an explicit one."""
    global _default_store
    with _default_lock:
        _default_store = store


def default_store() -> Store:
    """Resolve the process-wide store: explicit > B200_STORE_DIR > torch.distributed > Ray KV."""
    global _default_store
    with _default_lock:
        if _default_store is not None:
            return _default_store
        path = os.environ.get("B200_STORE_DIR")
        if path:
            _default_store = FileStore(path)
            return _default_store
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                from torch.distributed.distributed_c10d import _get_default_store

                _default_store = TorchDistStore(_get_default_store())
                return _default_store
        except Exception:
            pass
        try:
            _default_store = RayInternalKVStore()
            return _default_store
        except Exception as exc:
            raise RuntimeError(
                "no rendezvous store available: call ray_b200.store.set_default_store(), set "
                "B200_STORE_DIR, initialise torch.distributed, or run inside a Ray worker"
            ) from exc

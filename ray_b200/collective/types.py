"""Enums and per-call option holders of the collective API.

API-compatible with ``ray.util.collective.types`` (python/ray/util/collective/types.py:33-122):
the same public names, defaults and ``ReduceOp`` numbering (SUM 0, PRODUCT 1, MIN 2, MAX 3),
so backends and call sites written against the reference work unchanged.  Two behaviours a
drop-in must keep:

* option holders carry their fields as *class attributes*; ``collective.allreduce`` mutates
  and passes the class object itself (collective.py:329-331), so a backend only ever reads
  ``opts.reduceOp`` and friends and never assumes an instance;
* ``Backend("name")`` is a lookup, not a constructor: it upper-cases, maps ``torch_gloo`` to
  GLOO, raises ``ValueError`` for unknown names, and backends registered later show up as
  attributes (backend_registry.py:118-122).
"""
from __future__ import annotations

import enum
from datetime import timedelta

unset_timeout_ms = timedelta(milliseconds=-1)


class ReduceOp(enum.Enum):
    SUM = 0
    PRODUCT = 1
    MIN = 2
    MAX = 3


class _BackendMeta(type):
    def __call__(cls, name: str):  # Backend("b200") -> "B200"
        key = str(name).upper()
        if key == "TORCH_GLOO":
            return cls.GLOO
        found = cls.__dict__.get(key)
        if not isinstance(found, str) or key == "UNRECOGNIZED":
            known = ", ".join(sorted(k for k, v in cls.__dict__.items()
                                     if isinstance(v, str) and k.isupper() and k != "UNRECOGNIZED"))
            raise ValueError(f"Unrecognized backend: '{name}'. Known backends: {known}")
        return found


class Backend(metaclass=_BackendMeta):
    """Backend names are plain upper-case strings."""

    NCCL = "NCCL"
    GLOO = "GLOO"
    B200 = "B200"
    UNRECOGNIZED = "unrecognized"


def _options(name: str, doc: str, **fields):
    """A holder whose fields live on the class: readable from the class or an instance,
    assignable on either (``opts = ReduceOptions(); opts.root_rank = 3`` and
    ``AllReduceOptions.reduceOp = op`` both occur in collective.py)."""
    cls = type(name, (), {"__doc__": doc, **fields})
    cls.__module__ = __name__
    return cls


AllReduceOptions = _options(
    "AllReduceOptions", "reduceOp for allreduce.", reduceOp=ReduceOp.SUM, timeout_ms=unset_timeout_ms)
BarrierOptions = _options("BarrierOptions", "barrier options.", timeout_ms=unset_timeout_ms)
ReduceOptions = _options(
    "ReduceOptions", "reduce-to-root options; root_tensor only matters for the legacy multi-GPU API.",
    reduceOp=ReduceOp.SUM, root_rank=0, root_tensor=0, timeout_ms=unset_timeout_ms)
AllGatherOptions = _options("AllGatherOptions", "allgather options.", timeout_ms=unset_timeout_ms)
BroadcastOptions = _options(
    "BroadcastOptions", "broadcast options.", root_rank=0, root_tensor=0, timeout_ms=unset_timeout_ms)
ReduceScatterOptions = _options(
    "ReduceScatterOptions", "reducescatter options.", reduceOp=ReduceOp.SUM, timeout_ms=unset_timeout_ms)
SendOptions = _options(
    "SendOptions", "send options.", dst_rank=0, dst_gpu_index=0, n_elements=0, timeout_ms=unset_timeout_ms)
RecvOptions = _options(
    "RecvOptions", "recv options (the reference spells its timeout field unset_timeout_ms).",
    src_rank=0, src_gpu_index=0, n_elements=0, unset_timeout_ms=unset_timeout_ms)


def torch_available() -> bool:
    return True


def cupy_available() -> bool:
    try:
        import cupy  # noqa: F401

        return True
    except ImportError:
        return False

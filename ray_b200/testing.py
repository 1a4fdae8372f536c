"""Test / benchmark harness: several ranks of a group inside ONE process.

``LocalGroup(n)`` creates n communicators that exchange handles through an in-process
``DictStore``.  With >= n visible GPUs every rank gets its own device (real NVLink /
NVSwitch traffic, NVLS when available); with fewer GPUs the ranks share devices and the
"peer" loads and stores resolve to the same HBM -- every kernel, flag protocol and code
path except the multicast instructions is exercised exactly as in production, which is
what lets the parity suite run on a single-GPU box.

Collective calls only enqueue kernels, so one host thread can issue rank 0's call, then
rank 1's, ... on per-rank streams; the kernels meet on the device.  When ranks share a
GPU the CTA count is capped so all n grids are co-resident (a spinning grid that fills
the GPU would starve its peers).
"""
from __future__ import annotations

import threading
from typing import Callable, List, Optional, Sequence

import torch

from .comm import B200Comm
from .store import DictStore

_group_serial = 0


class LocalGroup:
    def __init__(self, world_size: int, devices: Optional[Sequence[int]] = None, timeout_ms: int = 10000,
                 staging_bytes: int = 32 << 20, heap_bytes: int = 0, inbox_bytes: int = 8 << 20,
                 enable_multicast: bool = True):
        global _group_serial
        ndev = torch.cuda.device_count()
        if ndev == 0:
            raise RuntimeError("LocalGroup needs at least one CUDA device")
        if devices is None:
            devices = [r % ndev for r in range(world_size)] if ndev < world_size else list(range(world_size))
        self.world_size = world_size
        self.devices = list(devices)
        self.shared_gpu = len(set(self.devices)) < world_size
        _group_serial += 1
        store = DictStore()
        name = f"local{_group_serial}"
        self.comms: List[Optional[B200Comm]] = [None] * world_size
        errors: List[BaseException] = []

        def make(rank: int) -> None:
            try:
                self.comms[rank] = B200Comm(
                    world_size, rank, self.devices[rank], store=store, group_name=name,
                    staging_bytes=staging_bytes, heap_bytes=heap_bytes, inbox_bytes=inbox_bytes,
                    enable_multicast=enable_multicast, timeout_ms=timeout_ms, rendezvous_timeout_s=60.0)
            except BaseException as exc:  # noqa: BLE001 - surfaced below
                errors.append(exc)

        threads = [threading.Thread(target=make, args=(r,)) for r in range(world_size)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            self.destroy()
            raise errors[0]
        self.streams = [torch.cuda.Stream(device=d) for d in self.devices]
        if self.shared_gpu:
            # all grids of the ranks sharing a device must be resident at the same time
            per_dev = max(self.devices.count(d) for d in set(self.devices))
            sms = torch.cuda.get_device_properties(self.devices[0]).multi_processor_count
            blocks = max(1, (sms - 4) // per_dev)
            for c in self.comms:
                c.set_blocks(blocks)

    @property
    def has_multicast(self) -> bool:
        return all(c.has_multicast for c in self.comms)

    def device(self, rank: int) -> torch.device:
        return torch.device("cuda", self.devices[rank])

    def run(self, fn: Callable[[B200Comm, int], None]) -> None:
        """Issue ``fn(comm, rank)`` for every rank on that rank's stream, then wait."""
        for r, c in enumerate(self.comms):
            # operands are usually produced on the device's default stream: order after it
            self.streams[r].wait_stream(torch.cuda.current_stream(self.devices[r]))
            with torch.cuda.device(self.devices[r]), torch.cuda.stream(self.streams[r]):
                fn(c, r)
        self.synchronize()

    def synchronize(self, check: bool = True) -> None:
        for r, s in enumerate(self.streams):
            s.synchronize()
        if check:
            for c in self.comms:
                c.check_status()

    def destroy(self) -> None:
        for c in self.comms:
            if c is not None:
                c.abort()
        for c in self.comms:
            if c is not None:
                c.destroy()
        self.comms = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

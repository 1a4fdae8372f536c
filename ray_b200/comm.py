"""B200Comm: one rank's communicator -- a thin torch.Tensor <-> C-ABI adapter.

All arithmetic and data movement happens in libb200_collective.so; this file only turns
tensors into (pointer, count, dtype, stream) tuples, mirrors the argument checks of the
reference backends, and performs the one-time handle exchange through a ``Store``.
"""
from __future__ import annotations

import ctypes
import threading
from typing import List, Optional, Sequence

import torch

from . import _native as N
from .store import Store, default_store

# torch dtype -> b200_dtype_t.  Same coverage as the reference's TORCH_NCCL_DTYPE_MAP
# (util/collective/collective_group/nccl_util.py:51-71); torch.bool travels as int8 there.
TORCH_DTYPE_MAP = {
    torch.bool: N.I8,
    torch.uint8: N.U8,
    torch.int8: N.I8,
    torch.int32: N.I32,
    torch.int64: N.I64,
    torch.float16: N.F16,
    torch.bfloat16: N.BF16,
    torch.float32: N.F32,
    torch.float64: N.F64,
}
for _name, _code in (("uint32", N.U32), ("uint64", N.U64)):
    if hasattr(torch, _name):
        TORCH_DTYPE_MAP[getattr(torch, _name)] = _code


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return TORCH_DTYPE_MAP[dtype]
    except KeyError:
        raise ValueError(f"dtype {dtype} is not supported by the B200 collective backend") from None


def _check_cuda_contiguous(t: torch.Tensor, what: str = "tensor") -> None:
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{what} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        # same wording as nccl_util.get_tensor_ptr (nccl_util.py:170-173)
        raise RuntimeError("Torch tensor must be on GPU when using B200 collectives.")
    if not t.is_contiguous():
        raise RuntimeError(f"{what} must be contiguous")


class SymmetricTensorHolder:
    """Exposes a slice of the symmetric heap through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self._owner = owner  # keeps the communicator alive
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 3,
            "strides": None,
        }


class B200Comm:
    """One rank of a B200 collective group.

    Args:
        world_size, rank: group geometry (<= 8 ranks: one NVSwitch domain).
        device: CUDA device ordinal of this rank in this process.
        store: rendezvous store shared by all ranks (default: ``default_store()``).
        group_name: namespaces the store keys, so several groups can coexist
            (the reference tests create 5 at once).
        staging_bytes / heap_bytes / inbox_bytes / enable_multicast / timeout_ms:
            see ``b200_config_t`` in include/b200_collective.h.
    """

    def __init__(
        self,
        world_size: int,
        rank: int,
        device: int,
        store: Optional[Store] = None,
        group_name: str = "default",
        staging_bytes: int = 0,
        heap_bytes: int = 0,
        inbox_bytes: int = 0,
        enable_multicast: bool = True,
        timeout_ms: int = 0,
        rendezvous_timeout_s: float = 180.0,
    ):
        self._lib = N.load()
        self._h = ctypes.c_void_p()
        self.world_size = int(world_size)
        self.rank = int(rank)
        self.device = int(device)
        self.group_name = group_name
        self._closed = False
        self._lock = threading.Lock()
        cfg = N.B200Config(int(staging_bytes), int(heap_bytes), int(inbox_bytes),
                           1 if enable_multicast else 0, int(timeout_ms))
        N.check(self._lib.b200_comm_create(self.world_size, self.rank, self.device,
                                           ctypes.byref(cfg), ctypes.byref(self._h)))
        try:
            store = store if store is not None else default_store()
            blob = ctypes.create_string_buffer(N.HANDLE_BYTES)
            N.check(self._lib.b200_comm_export_handle(self._h, blob))
            prefix = f"b200/{group_name}/handle/"
            store.set(prefix + str(self.rank), blob.raw)
            blobs = b"".join(store.get(prefix + str(p), rendezvous_timeout_s) for p in range(self.world_size))
            N.check(self._lib.b200_comm_connect(self._h, blobs))
            # every rank has read every handle once connect() returned on all ranks; the
            # owner removes its key so the name can be reused (destroy / re-init)
            store.delete(prefix + str(self.rank))
        except BaseException:
            self._lib.b200_comm_destroy(self._h)
            self._h = ctypes.c_void_p()
            self._closed = True
            raise

    # ------------------------------------------------------------------ helpers
    def _stream(self, stream: Optional[torch.cuda.Stream] = None) -> int:
        """Raw handle of the stream an op is enqueued on (default: this device's current stream).

        The kernels read the communicator's launch counter at entry and rely on STREAM ORDER for
        their epoch and slot parity (include/b200_collective.h: "collectives of one communicator
        must be stream-ordered").  A caller that switches streams between two ops of the same
        communicator is therefore ordered here on the device: the new stream waits for
        an event recorded at the tail of the previous one.  Costs nothing while the stream stays
        the same."""
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        last = getattr(self, "_last_stream", None)
        if last is not None and last.cuda_stream != st.cuda_stream:
            ev = torch.cuda.Event()
            ev.record(last)
            st.wait_event(ev)
        self._last_stream = st
        return st.cuda_stream

    @property
    def has_multicast(self) -> bool:
        return bool(self._lib.b200_comm_has_multicast(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.b200_comm_launch_count(self._h))

    def set_blocks(self, nblocks: int) -> None:
        N.check(self._lib.b200_comm_set_blocks(self._h, int(nblocks)))

    def set_param(self, param: int, value: int) -> None:
        """Tuning knob (``N.PARAM_*``); must be set identically on every rank."""
        N.check(self._lib.b200_comm_set_param(self._h, int(param), int(value)))

    def trace_enable(self, capacity: int) -> None:
        """Profiling aid: let instrumented kernels record up to ``capacity`` timestamped events."""
        N.check(self._lib.b200_comm_trace_enable(self._h, int(capacity)))

    def trace_read(self, max_events: int = 1 << 20, reset: bool = True):
        """-> list of (ns, cta, event, arg) recorded since the last reset (synchronises the device)."""
        buf = (ctypes.c_ulonglong * (2 * max_events))()
        n = self._lib.b200_comm_trace_read(self._h, buf, max_events, 1 if reset else 0)
        if n < 0:
            N.check(n)
        return [(buf[2 * i], buf[2 * i + 1] >> 40, (buf[2 * i + 1] >> 32) & 0xFF, buf[2 * i + 1] & 0xFFFFFFFF)
                for i in range(n)]

    def status(self) -> int:
        return int(self._lib.b200_comm_status(self._h))

    def check_status(self) -> None:
        """Raise if a kernel of this communicator gave up (abort / watchdog).  Only
        meaningful after the stream was synchronised."""
        st = self.status()
        if st == N.ERR_ABORTED:
            raise N.B200AbortedError(st, "communicator aborted")
        if st == N.ERR_TIMEOUT:
            raise N.B200TimeoutError(st, "a peer did not arrive before the device watchdog expired")
        if st != 0:
            raise N.B200Error(st, N.last_error())

    # ------------------------------------------------------------------ symmetric heap
    def symm_empty(self, shape, dtype=torch.float32) -> torch.Tensor:
        """Collectively allocate a tensor in the symmetric heap (zero-copy operand)."""
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)))
        numel = 1
        for s in shape:
            numel *= s
        nbytes = max(numel * torch.empty((), dtype=dtype).element_size(), 1)
        ptr = ctypes.c_void_p()
        N.check(self._lib.b200_symm_alloc(self._h, nbytes, ctypes.byref(ptr)))
        holder = SymmetricTensorHolder(ptr.value, nbytes, self)
        raw = torch.as_tensor(holder, device=torch.device("cuda", self.device))
        t = raw.view(dtype)[:numel].view(shape)
        t._b200_holder = holder  # noqa: SLF001 - keep the mapping alive with the tensor
        return t

    def mem_pool(self):
        """A ``torch.cuda.MemPool`` whose memory is this communicator's symmetric heap: tensors
        created under ``with torch.cuda.use_mem_pool(comm.mem_pool()):`` are ordinary torch
        tensors, yet collectives reduce them in place with no staging copies (what
        ``ncclMemAlloc`` + buffer registration buys on the NCCL side).  Every rank must create
        the same tensors in the same order.  One communicator per process can back the pool."""
        if getattr(self, "_pool", None) is None:
            from torch.cuda.memory import CUDAPluggableAllocator

            N.check(self._lib.b200_pool_bind(self._h))
            self._allocator = CUDAPluggableAllocator(str(N.LIB_PATH), "b200_pool_alloc", "b200_pool_free")
            self._pool = torch.cuda.MemPool(self._allocator.allocator())
        return self._pool

    def symm_reset(self) -> None:
        N.check(self._lib.b200_symm_reset(self._h))

    def symm_contains(self, t: torch.Tensor) -> bool:
        return bool(self._lib.b200_symm_contains(self._h, t.data_ptr(), t.numel() * t.element_size()))

    # ------------------------------------------------------------------ collectives
    def allreduce(self, tensor: torch.Tensor, op: int = N.SUM, out: Optional[torch.Tensor] = None,
                  algo: int = N.ALGO_AUTO) -> None:
        _check_cuda_contiguous(tensor)
        out = tensor if out is None else out
        if out is not tensor:
            _check_cuda_contiguous(out, "output tensor")
            if out.dtype != tensor.dtype or out.numel() != tensor.numel():
                raise RuntimeError("allreduce output must match the input's dtype and size")
        N.check(self._lib.b200_allreduce(self._h, tensor.data_ptr(), out.data_ptr(), tensor.numel(),
                                         dtype_code(tensor.dtype), int(op), int(algo), self._stream()))

    def allgather(self, outs: Sequence[torch.Tensor], tensor: torch.Tensor) -> None:
        _check_cuda_contiguous(tensor)
        if len(outs) != self.world_size:
            raise RuntimeError("The length of the tensor list operands to allgather must be equal to world_size.")
        arr = (ctypes.c_void_p * N.MAX_RANKS)()
        for i, o in enumerate(outs):
            _check_cuda_contiguous(o, "output tensor")
            if o.dtype != tensor.dtype or o.numel() != tensor.numel():
                raise RuntimeError("All tensor operands to allgather must have the same dtype and size.")
            arr[i] = o.data_ptr()
        N.check(self._lib.b200_allgather(self._h, tensor.data_ptr(), arr, tensor.numel(),
                                         dtype_code(tensor.dtype), self._stream()))

    def allgather_into(self, out: torch.Tensor, tensor: torch.Tensor) -> None:
        """out = concat over ranks along dim 0 (the Compiled-Graph layout, collective_node.py:198-206)."""
        _check_cuda_contiguous(tensor)
        _check_cuda_contiguous(out, "output tensor")
        if out.dtype != tensor.dtype or out.numel() != tensor.numel() * self.world_size:
            raise RuntimeError("allgather output must hold world_size copies of the input")
        arr = (ctypes.c_void_p * N.MAX_RANKS)()
        step = tensor.numel() * tensor.element_size()
        for p in range(self.world_size):
            arr[p] = out.data_ptr() + p * step
        N.check(self._lib.b200_allgather(self._h, tensor.data_ptr(), arr, tensor.numel(),
                                         dtype_code(tensor.dtype), self._stream()))

    def reducescatter(self, out: torch.Tensor, ins: Sequence[torch.Tensor], op: int = N.SUM) -> None:
        _check_cuda_contiguous(out, "output tensor")
        if len(ins) != self.world_size:
            raise RuntimeError("The length of the tensor list operands to reducescatter must be equal to world_size.")
        arr = (ctypes.c_void_p * N.MAX_RANKS)()
        for i, t in enumerate(ins):
            _check_cuda_contiguous(t)
            if t.dtype != out.dtype or t.numel() != out.numel():
                raise RuntimeError("All tensor operands to reducescatter must have the same dtype and size.")
            arr[i] = t.data_ptr()
        N.check(self._lib.b200_reducescatter(self._h, arr, out.data_ptr(), out.numel(),
                                             dtype_code(out.dtype), int(op), self._stream()))

    def reducescatter_from(self, out: torch.Tensor, tensor: torch.Tensor, op: int = N.SUM) -> None:
        """out = reduce over ranks of this rank's 1/world slice of ``tensor`` along dim 0
        (the Compiled-Graph layout, collective_node.py:207-219)."""
        _check_cuda_contiguous(tensor)
        _check_cuda_contiguous(out, "output tensor")
        if out.dtype != tensor.dtype or out.numel() * self.world_size != tensor.numel():
            raise RuntimeError("reducescatter input must hold world_size slices of the output size")
        arr = (ctypes.c_void_p * N.MAX_RANKS)()
        step = out.numel() * out.element_size()
        for p in range(self.world_size):
            arr[p] = tensor.data_ptr() + p * step
        N.check(self._lib.b200_reducescatter(self._h, arr, out.data_ptr(), out.numel(),
                                             dtype_code(out.dtype), int(op), self._stream()))

    def broadcast(self, tensor: torch.Tensor, root: int = 0) -> None:
        _check_cuda_contiguous(tensor)
        N.check(self._lib.b200_broadcast(self._h, tensor.data_ptr(), tensor.numel(),
                                         dtype_code(tensor.dtype), int(root), self._stream()))

    def reduce(self, tensor: torch.Tensor, root: int = 0, op: int = N.SUM) -> None:
        _check_cuda_contiguous(tensor)
        N.check(self._lib.b200_reduce(self._h, tensor.data_ptr(), tensor.numel(),
                                      dtype_code(tensor.dtype), int(op), int(root), self._stream()))

    def barrier(self) -> None:
        N.check(self._lib.b200_barrier(self._h, self._stream()))

    def send(self, tensor: torch.Tensor, peer: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Enqueue a send on ``stream`` (default: the current stream of this rank's device)."""
        _check_cuda_contiguous(tensor)
        st = self._lib.b200_send(self._h, tensor.data_ptr(), tensor.numel() * tensor.element_size(), int(peer),
                                 stream.cuda_stream if stream is not None else self._stream())
        if st:
            N.check(st)

    def recv(self, tensor: torch.Tensor, peer: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        _check_cuda_contiguous(tensor)
        st = self._lib.b200_recv(self._h, tensor.data_ptr(), tensor.numel() * tensor.element_size(), int(peer),
                                 stream.cuda_stream if stream is not None else self._stream())
        if st:
            N.check(st)

    def send_ptr(self, ptr: int, nbytes: int, peer: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        """send() from a raw device-visible address (e.g. pinned host memory under unified
        addressing: the Compiled-Graph channel keeps its metadata header there)."""
        st = self._lib.b200_send(self._h, ptr, int(nbytes), int(peer),
                                 stream.cuda_stream if stream is not None else self._stream())
        if st:
            N.check(st)

    def recv_ptr(self, ptr: int, nbytes: int, peer: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        st = self._lib.b200_recv(self._h, ptr, int(nbytes), int(peer),
                                 stream.cuda_stream if stream is not None else self._stream())
        if st:
            N.check(st)

    # ------------------------------------------------------------------ one-sided get
    def heap_range(self):
        """(base address, bytes) of this rank's symmetric heap."""
        base, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        N.check(self._lib.b200_symm_base(self._h, ctypes.byref(base), ctypes.byref(nbytes)))
        return int(base.value or 0), int(nbytes.value)

    def heap_view(self, offset: int, nbytes: int) -> torch.Tensor:
        """uint8 tensor over [offset, offset+nbytes) of this rank's heap (no allocation bookkeeping)."""
        base, size = self.heap_range()
        if offset < 0 or offset + nbytes > size:
            raise ValueError("range outside the symmetric heap")
        holder = SymmetricTensorHolder(base + offset, max(nbytes, 1), self)
        t = torch.as_tensor(holder, device=torch.device("cuda", self.device))[:nbytes]
        t._b200_holder = holder  # noqa: SLF001
        return t

    def get(self, dst: torch.Tensor, src_rank: int, src_heap_offset: int,
            stream: Optional[torch.cuda.Stream] = None) -> None:
        """Pull ``dst.nbytes`` bytes from ``src_rank``'s symmetric heap into ``dst``; only this rank
        runs a kernel."""
        _check_cuda_contiguous(dst)
        N.check(self._lib.b200_get(self._h, dst.data_ptr(), int(src_rank), int(src_heap_offset),
                                   dst.numel() * dst.element_size(),
                                   stream.cuda_stream if stream is not None else self._stream()))

    def grad_allreduce(self, grad: torch.Tensor, scale: float, wire_dtype: torch.dtype = torch.bfloat16) -> None:
        """Fused ``grad = sum_r wire(grad_r * scale)`` on a flat fp32 bucket."""
        _check_cuda_contiguous(grad)
        if grad.dtype != torch.float32:
            raise RuntimeError("grad_allreduce expects a float32 bucket")
        N.check(self._lib.b200_grad_allreduce(self._h, grad.data_ptr(), grad.numel(), float(scale),
                                              dtype_code(wire_dtype), self._stream()))

    def allreduce_multi(self, tensors: List[torch.Tensor], op: int = N.SUM) -> None:
        if not tensors:
            return
        dt = tensors[0].dtype
        ptrs = (ctypes.c_void_p * len(tensors))()
        counts = (ctypes.c_size_t * len(tensors))()
        for i, t in enumerate(tensors):
            _check_cuda_contiguous(t)
            if t.dtype != dt:
                raise ValueError("Expected all input tensors to have the same dtype")
            ptrs[i] = t.data_ptr()
            counts[i] = t.numel()
        N.check(self._lib.b200_allreduce_multi(self._h, ptrs, counts, len(tensors), dtype_code(dt),
                                               int(op), self._stream()))

    # ------------------------------------------------------------------ lifecycle
    def abort(self) -> None:
        if self._h:
            self._lib.b200_comm_abort(self._h)

    def destroy(self) -> None:
        with self._lock:
            if self._closed:
                return
            self._closed = True
        self._lib.b200_comm_destroy(self._h)
        self._h = ctypes.c_void_p()

    @property
    def closed(self) -> bool:
        return self._closed

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

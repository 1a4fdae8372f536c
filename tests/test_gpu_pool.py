"""Zero-copy operands through torch's pluggable allocator: tensors created under the
communicator's MemPool live in the symmetric heap and are reduced in place (no staging)."""
import os
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, store_dir, out_dir):
    sys.path.insert(0, ROOT)
    from ray_b200 import _native as N
    from ray_b200.comm import B200Comm
    from ray_b200.store import FileStore

    ndev = torch.cuda.device_count()
    dev = rank if ndev >= world else 0
    torch.cuda.set_device(dev)
    comm = B200Comm(world, rank, dev, store=FileStore(store_dir), group_name="pool", staging_bytes=8 << 20,
                    heap_bytes=64 << 20, inbox_bytes=2 << 20, timeout_ms=20000)
    if ndev < world:
        comm.set_blocks(32)
    pool = comm.mem_pool()
    with torch.cuda.use_mem_pool(pool):
        x = torch.full((1 << 20,), float(rank + 1), device=f"cuda:{dev}")
        y = torch.arange(1000, device=f"cuda:{dev}", dtype=torch.float32) * (rank + 1)
    plain = torch.ones(10, device=f"cuda:{dev}")
    assert comm.symm_contains(x) and comm.symm_contains(y) and not comm.symm_contains(plain)
    torch.cuda.synchronize()
    before = comm.launch_count
    comm.allreduce(x)           # 4 MiB, in place, zero copy: two-shot / NVLS directly on the heap
    comm.allreduce(y, algo=N.ALGO_TWOSHOT)
    torch.cuda.synchronize()
    comm.check_status()
    assert comm.launch_count == before + 2
    total = sum(range(1, world + 1))
    assert torch.all(x == total)
    assert torch.equal(y.cpu(), torch.arange(1000, dtype=torch.float32) * total)
    # freed blocks are recycled by size
    ptr = x.data_ptr()
    del x
    pool_again = comm.mem_pool()
    assert pool_again is pool
    comm.barrier()
    torch.cuda.synchronize()
    comm.destroy()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write(str(ptr))


def test_mem_pool_tensors_are_zero_copy_operands(native_lib):
    world = 2
    with tempfile.TemporaryDirectory() as store_dir, tempfile.TemporaryDirectory() as out_dir:
        mp.spawn(_worker, args=(world, store_dir, out_dir), nprocs=world, join=True)
        assert all(os.path.exists(os.path.join(out_dir, f"ok{r}")) for r in range(world))


def test_bootstrap_endpoint_requires_the_token_and_stops_serving_after_connect(native_lib):
    """the fd-passing endpoint must not hand GPU memory to whoever connects.
    Before connect: a request without the 128-bit token from the handle blob gets nothing.  After
    connect: the endpoint is closed."""
    import ctypes
    import socket
    import struct
    import time

    from ray_b200 import _native as N
    from ray_b200.testing import LocalGroup

    lib = native_lib
    h = ctypes.c_void_p()
    cfg = N.B200Config(2 << 20, 0, 2 << 20, 0, 1000)
    N.check(lib.b200_comm_create(2, 0, 0, ctypes.byref(cfg), ctypes.byref(h)))
    try:
        blob = ctypes.create_string_buffer(N.HANDLE_BYTES)
        N.check(lib.b200_comm_export_handle(h, blob))
        raw = blob.raw
        name = raw[raw.index(b"b200coll-"):].split(b"\0")[0]
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.settimeout(3)
        s.connect(b"\0" + name)
        s.sendall(struct.pack("<IIIi16s", 0xB200C011, 1, 0, 0, b"\0" * 16))  # GET_FD(data), wrong token
        msg, anc, _, _ = s.recvmsg(4, socket.CMSG_SPACE(4))
        assert msg == b"" and not anc, "a request without the token must be dropped"
        s.close()
    finally:
        lib.b200_comm_destroy(h)
    with LocalGroup(2, staging_bytes=2 << 20, inbox_bytes=2 << 20) as g:
        blob = ctypes.create_string_buffer(N.HANDLE_BYTES)
        N.check(lib.b200_comm_export_handle(g.comms[0]._h, blob))
        name = blob.raw[blob.raw.index(b"b200coll-"):].split(b"\0")[0]
        time.sleep(0.3)  # the server thread leaves its loop within one 100 ms poll
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        with pytest.raises((ConnectionRefusedError, FileNotFoundError)):
            s.connect(b"\0" + name)
        s.close()
        x = [torch.ones(10, device=g.device(r)) for r in range(2)]
        g.run(lambda c, r: c.allreduce(x[r]))  # the group works without its endpoint
        assert all(torch.all(t == 2) for t in x)

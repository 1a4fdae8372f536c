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

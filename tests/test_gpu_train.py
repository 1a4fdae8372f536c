"""The c10d backend + DDP gradient path (boundary B4), with real worker processes.

Each worker is its own process (as a Ray Train worker is); with fewer GPUs than workers the
processes share cuda:0 and the GPU time-slices between them, which is slow but exercises the
same cross-process fd-passing bootstrap, kernels and stream semantics.
"""
import os
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, out_dir, wire_name):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import torch.nn as nn

    from ray_b200 import train as T

    ndev = torch.cuda.device_count()
    os.environ["LOCAL_RANK"] = str(rank if ndev >= world else 0)
    device = T.get_device()
    torch.cuda.set_device(device)
    cfg = T.B200TorchConfig()
    backend = T.resolve_backend(cfg.backend, use_gpu=True)
    assert backend == "cpu:gloo,cuda:b200"
    T.setup_torch_process_group(backend, rank, world, f"file://{init_file}", timeout_s=120)
    pg = dist.distributed_c10d._get_default_group()
    assert isinstance(pg, T.B200ProcessGroup)
    if ndev < world:
        # co-resident grids when the workers share one GPU
        x = torch.zeros(1, device=device)
        dist.all_reduce(x)
        pg.comm.set_blocks(32)

    # --- plain c10d calls on CUDA tensors -------------------------------------------------
    t = torch.ones(1000, device=device) * (rank + 1)
    dist.all_reduce(t)
    assert torch.all(t == sum(range(1, world + 1)))
    t = torch.ones(7, device=device) * (rank + 1)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert torch.all(t == world)
    t = torch.full((5,), float(rank), device=device)
    dist.broadcast(t, src=world - 1)
    assert torch.all(t == world - 1)
    outs = [torch.zeros(3, device=device) for _ in range(world)]
    dist.all_gather(outs, torch.full((3,), float(rank), device=device))
    assert all(torch.all(outs[p] == p) for p in range(world))
    big = torch.zeros(3 * world, device=device)
    dist.all_gather_into_tensor(big, torch.full((3,), float(rank), device=device))
    assert torch.equal(big.cpu(), torch.arange(world).repeat_interleave(3).float())
    rs = torch.zeros(4, device=device)
    dist.reduce_scatter_tensor(rs, torch.arange(4 * world, device=device, dtype=torch.float32))
    assert torch.equal(rs.cpu(), torch.arange(4 * rank, 4 * rank + 4).float() * world)
    # CPU tensors are served by the gloo side
    c = torch.ones(3) * (rank + 1)
    dist.all_reduce(c)
    assert torch.all(c == sum(range(1, world + 1)))
    if rank == 0:
        dist.send(torch.arange(10, device=device, dtype=torch.float32), dst=1)
    elif rank == 1:
        r = torch.zeros(10, device=device)
        dist.recv(r, src=0)
        assert torch.equal(r.cpu(), torch.arange(10).float())
    dist.barrier()

    # --- DDP: TorchTrainer's gradient path -------------------------------------------------
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(64, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Linear(128, 10))
    ref = nn.Sequential(nn.Linear(64, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Linear(128, 10)).to(device)
    ref.load_state_dict(model.state_dict())
    wire = {"f32": torch.float32, "bf16": torch.bfloat16, "none": None}[wire_name]
    ddp = T.prepare_model(model, gradient_wire_dtype=wire)
    assert isinstance(ddp, torch.nn.parallel.DistributedDataParallel)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.1)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    loss_fn = nn.CrossEntropyLoss()
    for step in range(3):
        xs = [torch.randn(16, 64, generator=torch.Generator().manual_seed(100 * step + r)).to(device)
              for r in range(world)]
        ys = [torch.randint(0, 10, (16,), generator=torch.Generator().manual_seed(200 * step + r)).to(device)
              for r in range(world)]
        loss_fn(ddp(xs[rank]), ys[rank]).backward()
        opt.step()
        opt.zero_grad()
        # reference: the mean over ranks of the per-rank gradients, computed locally
        grads = None
        bn_state = {k: v.clone() for k, v in ref.state_dict().items() if "running" in k or "num_batches" in k}
        for r in range(world):
            ref.load_state_dict({**ref.state_dict(), **bn_state})
            ref.zero_grad()
            loss_fn(ref(xs[r]), ys[r]).backward()
            g = [p.grad.clone() for p in ref.parameters()]
            grads = g if grads is None else [a + b for a, b in zip(grads, g)]
        for p, g in zip(ref.parameters(), grads):
            p.grad = g / world
        ref_opt.step()
        tol = 1e-5 if wire_name != "bf16" else 2e-2
        for (n1, p1), (_, p2) in zip(ddp.module.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p1, p2, atol=tol, rtol=tol), (step, n1, (p1 - p2).abs().max().item())
    # replicas identical across ranks, bit for bit
    flat = torch.cat([p.detach().flatten() for p in ddp.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    for gth in gathered[1:]:
        assert torch.equal(gth, gathered[0])
    launches = pg.comm.launch_count
    torch.cuda.synchronize()
    pg.comm.check_status()
    dist.barrier()
    dist.destroy_process_group()
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write(str(launches))


@pytest.mark.parametrize("wire", ["bf16", "f32", "none"])
def test_c10d_backend_and_ddp_gradient_path(native_lib, wire):
    world = 2
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rdzv")
        mp.spawn(_worker, args=(world, init_file, d, wire), nprocs=world, join=True)
        launches = [int(open(os.path.join(d, f"ok{r}")).read()) for r in range(world)]
        assert all(l > 0 for l in launches), "no native kernels were launched"


def _fsdp_worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import torch.nn as nn
    from torch.distributed.fsdp import FullyShardedDataParallel as FSDP

    from ray_b200 import train as T

    ndev = torch.cuda.device_count()
    os.environ["LOCAL_RANK"] = str(rank if ndev >= world else 0)
    device = T.get_device()
    torch.cuda.set_device(device)
    T.setup_torch_process_group(T.DEFAULT_GPU_BACKEND, rank, world, f"file://{init_file}", timeout_s=120)
    pg = dist.distributed_c10d._get_default_group()
    if ndev < world:
        x = torch.zeros(1, device=device)
        dist.all_reduce(x)
        pg.comm.set_blocks(32)

    # --- rooted and all-to-all ops of the c10d surface ------------------------------------------
    mine = torch.full((5,), float(rank + 1), device=device)
    outs = [torch.zeros(5, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, outs, dst=0)
    if rank == 0:
        assert all(torch.all(outs[p] == p + 1) for p in range(world))
    got = torch.zeros(3, device=device)
    dist.scatter(got, [torch.full((3,), 10.0 * p, device=device) for p in range(world)] if rank == 1 else None, src=1)
    assert torch.all(got == 10.0 * rank)
    src = torch.arange(4 * world, device=device, dtype=torch.float32) + 100 * rank
    dst = torch.zeros(4 * world, device=device)
    dist.all_to_all_single(dst, src)
    want = torch.cat([torch.arange(4 * rank, 4 * rank + 4, dtype=torch.float32) + 100 * p for p in range(world)])
    assert torch.equal(dst.cpu(), want)
    big_src = torch.randn(world * (3 << 20), device=device)  # 12 MiB per peer: larger than the eager ring
    big_dst = torch.zeros_like(big_src)
    dist.all_to_all_single(big_dst, big_src)
    gathered = [torch.empty_like(big_src) for _ in range(world)]
    dist.all_gather(gathered, big_src)
    n = 3 << 20
    for p in range(world):
        assert torch.equal(big_dst[p * n:(p + 1) * n], gathered[p][rank * n:(rank + 1) * n])
    work = dist.all_reduce(mine, async_op=True)
    work.wait()
    torch.cuda.synchronize()
    assert work.is_completed() and work.is_success() and work.exception() is None

    # --- FSDP through prepare_model(parallel_strategy="fsdp") (train_loop_utils.py:153-190) -------
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 8))
    ref = nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 8)).to(device)
    ref.load_state_dict(model.state_dict())
    fsdp = T.prepare_model(model, parallel_strategy="fsdp")
    assert isinstance(fsdp, FSDP)
    opt = torch.optim.SGD(fsdp.parameters(), lr=0.1)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    loss_fn = nn.MSELoss()
    for step in range(3):
        xs = [torch.randn(8, 32, generator=torch.Generator().manual_seed(10 * step + r)).to(device) for r in range(world)]
        ys = [torch.randn(8, 8, generator=torch.Generator().manual_seed(50 * step + r)).to(device) for r in range(world)]
        loss_fn(fsdp(xs[rank]), ys[rank]).backward()
        opt.step()
        opt.zero_grad()
        ref.zero_grad()
        sum(loss_fn(ref(xs[r]), ys[r]) for r in range(world)).div(world).backward()
        ref_opt.step()
        with FSDP.summon_full_params(fsdp):
            for (n1, p1), (_, p2) in zip(fsdp.module.named_parameters(), ref.named_parameters()):
                assert torch.allclose(p1, p2, atol=1e-5, rtol=1e-5), (step, n1, (p1 - p2).abs().max().item())
    launches = pg.comm.launch_count
    torch.cuda.synchronize()
    pg.comm.check_status()
    dist.barrier()
    dist.destroy_process_group()
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write(str(launches))


def test_c10d_rooted_ops_work_semantics_and_fsdp(native_lib):
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_fsdp_worker, args=(world, os.path.join(d, "rdzv"), d), nprocs=world, join=True)
        assert all(int(open(os.path.join(d, f"ok{r}")).read()) > 0 for r in range(world))


def _late_rank_worker(rank, world, init_file, out_dir):
    """Rank 1 never joins the collective: rank 0's watchdog must surface as an exception through
    Work.wait(timeout) / barrier, and later launches must be refused (a timed-out
    kernel used to leave unreduced gradients with nobody told)."""
    sys.path.insert(0, ROOT)
    import time

    import torch.distributed as dist

    from ray_b200 import _native as N
    from ray_b200 import train as T

    ndev = torch.cuda.device_count()
    os.environ["LOCAL_RANK"] = str(rank if ndev >= world else 0)
    device = T.get_device()
    torch.cuda.set_device(device)
    T.setup_torch_process_group(T.DEFAULT_GPU_BACKEND, rank, world, f"file://{init_file}", timeout_s=2)
    x = torch.ones(10, device=device)
    dist.all_reduce(x)  # both ranks: creates the communicator (watchdog = the 2 s group timeout)
    torch.cuda.synchronize()
    ok = "skipped"
    if rank == 0:
        work = dist.all_reduce(x, async_op=True)  # rank 1 never issues this one
        t0 = time.time()
        try:
            work.wait(__import__("datetime").timedelta(seconds=30))
            torch.cuda.synchronize()
            ok = "no exception" if work.is_success() else "flagged"
        except N.B200Error:
            ok = "raised"
        assert time.time() - t0 < 20
        assert not work.is_success() and work.exception() is not None
        try:
            dist.all_reduce(x)
            ok = "later launch accepted"
        except N.B200Error:
            pass
    else:
        time.sleep(6)
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write(ok)
    os._exit(0)  # the group is broken by design: skip the collective teardown


def test_late_rank_surfaces_as_an_error_not_silent_divergence(native_lib):
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_late_rank_worker, args=(world, os.path.join(d, "rdzv"), d), nprocs=world, join=True)
        assert open(os.path.join(d, "ok0")).read() in ("raised", "flagged")

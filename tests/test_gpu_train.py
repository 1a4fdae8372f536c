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

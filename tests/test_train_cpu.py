"""CPU tests of the train-side host logic (boundary B4) with world_size-2 processes:
backend registration, ``B200TorchConfig`` / ``resolve_backend``, the process group's
CPU-tensor side (served by gloo), and loud failure of CUDA-only entry points without a GPU."""
import os
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from ray_b200 import train as T

    T.setup_torch_process_group("b200", rank, world, f"file://{init_file}", timeout_s=60)
    pg = dist.distributed_c10d._get_default_group()
    assert isinstance(pg, T.B200ProcessGroup) and pg.getBackendName() == "b200"
    assert dist.get_rank() == rank and dist.get_world_size() == world
    t = torch.ones(5) * (rank + 1)
    dist.all_reduce(t)
    assert torch.all(t == 3)
    t = torch.full((3,), float(rank))
    dist.broadcast(t, src=1)
    assert torch.all(t == 1)
    outs = [torch.zeros(2) for _ in range(world)]
    dist.all_gather(outs, torch.full((2,), float(rank)))
    assert [o[0].item() for o in outs] == [0.0, 1.0]
    objs = [None] * world
    dist.all_gather_object(objs, {"rank": rank})
    assert [o["rank"] for o in objs] == [0, 1]
    # rooted and all-to-all ops of the c10d surface: CPU tensors are served by the gloo side
    mine = torch.full((3,), float(rank + 1))
    gathered = [torch.zeros(3) for _ in range(world)] if rank == 1 else None
    dist.gather(mine, gathered, dst=1)
    if rank == 1:
        assert [g[0].item() for g in gathered] == [1.0, 2.0]
    got = torch.zeros(2)
    dist.scatter(got, [torch.full((2,), 5.0 + p) for p in range(world)] if rank == 0 else None, src=0)
    assert torch.all(got == 5.0 + rank)
    src = torch.arange(2 * world, dtype=torch.float32) + 10 * rank
    dst = torch.zeros(2 * world)
    dist.all_to_all_single(dst, src)
    assert dst.tolist() == [2.0 * rank, 2.0 * rank + 1, 10 + 2.0 * rank, 11 + 2.0 * rank]
    work = dist.all_reduce(torch.ones(2), async_op=True)
    work.wait()
    assert work.is_completed()
    dist.barrier()
    # subgroups: rank 0 alone first (rank 1 is not a member), then both -- the second group must
    # still rendezvous although the two ranks have created a different number of groups
    solo = dist.new_group([0], backend="b200")
    both = dist.new_group([0, 1], backend="b200")
    if rank == 0:
        s = torch.ones(2)
        dist.all_reduce(s, group=solo)
        assert torch.all(s == 1)
    t = torch.ones(4) * (rank + 1)
    dist.all_reduce(t, group=both)
    assert torch.all(t == 3)
    assert pg.comm is None, "no CUDA communicator may be created for CPU-only traffic"
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_b200_process_group_cpu_side_world2():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "rdzv"), d), nprocs=2, join=True)
        assert all(os.path.exists(os.path.join(d, f"ok{r}")) for r in range(2))


def test_torch_config_mirrors_reference_fields():
    from ray_b200 import train as T

    cfg = T.B200TorchConfig()
    assert (cfg.backend, cfg.init_method, cfg.timeout_s) == (None, "env", 1800)
    assert T.resolve_backend(None, use_gpu=True) == "cpu:gloo,cuda:b200"
    assert T.resolve_backend(None, use_gpu=False) == "gloo"
    assert T.resolve_backend("nccl", use_gpu=True) == "nccl"
    assert T.uses_b200("b200") and T.uses_b200("cpu:gloo,cuda:b200") and not T.uses_b200("cpu:gloo,cuda:nccl")
    if not T.torch_config.HAVE_RAY_TRAIN:
        assert cfg.to_dict() == {"backend": None, "init_method": "env", "timeout_s": 1800}
        assert cfg.init_url("127.0.0.1", 1234) == "env://" and os.environ["MASTER_PORT"] == "1234"
        assert T.B200TorchConfig(init_method="tcp").init_url("127.0.0.1", 5) == "tcp://127.0.0.1:5"
        with pytest.raises(ValueError, match="not supported"):
            T.B200TorchConfig(init_method="mpi").init_url("h", 1)


def test_grad_hook_has_ddp_compatible_annotations():
    import torch.distributed as dist

    from ray_b200.train import b200_grad_hook

    hook = b200_grad_hook(torch.bfloat16)
    assert hook.__annotations__["bucket"] is dist.GradBucket
    assert hook.__annotations__["return"] == torch.futures.Future[torch.Tensor]


def test_communicator_refuses_cpu_tensors_and_missing_gpu():
    from ray_b200.comm import _check_cuda_contiguous, dtype_code

    with pytest.raises(RuntimeError, match="must be on GPU"):
        _check_cuda_contiguous(torch.ones(3))
    with pytest.raises(ValueError, match="not supported"):
        dtype_code(torch.complex64)
    if not torch.cuda.is_available():
        from ray_b200 import _native as N
        from ray_b200.comm import B200Comm
        from ray_b200.store import DictStore

        with pytest.raises(N.B200Error):
            B200Comm(1, 0, 0, store=DictStore())


def test_more_than_eight_workers_fall_back_to_nccl(monkeypatch):
    """a b200 group is one NVSwitch domain (<= 8 ranks of one host); a larger
    TorchTrainer must get the reference's default backend instead of failing at the first CUDA op."""
    import torch.distributed as dist

    from ray_b200.train import torch_config as tc

    seen = {}
    monkeypatch.setattr(dist, "init_process_group", lambda **kw: seen.update(kw))
    tc.setup_torch_process_group(tc.DEFAULT_GPU_BACKEND, 0, 16, "env://", timeout_s=5)
    assert seen["backend"] == "nccl" and seen["world_size"] == 16
    tc.setup_torch_process_group(tc.DEFAULT_GPU_BACKEND, 0, 8, "env://", timeout_s=5)
    assert seen["backend"] == tc.DEFAULT_GPU_BACKEND

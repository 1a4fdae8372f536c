"""CPU tests of the host-side logic (no GPU, no compute through the C ABI).

* BASELINE config 1 -- ``collective.allreduce`` fp32, world_size=2, gloo backend: two real
  processes drive ``ray_b200.collective`` with the reference's CPU backend restated in
  ``oracle/gloo_group.py`` registered as "GLOO" (test infrastructure; the product never
  registers it).  Mirrors python/ray/util/collective/tests/single_node_cpu_tests/.
* Registry / Backend / option-holder behaviour (Q1, Q3).
* Rendezvous stores.
"""
import os
import sys
import tempfile
import threading

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, store_dir, out_dir):
    sys.path.insert(0, ROOT)
    from oracle.gloo_group import GlooOracleGroup
    from ray_b200 import collective as col
    from ray_b200.collective.registry import _global_registry
    from ray_b200.store import FileStore, set_default_store

    set_default_store(FileStore(store_dir))
    if not _global_registry.is_registered("GLOO"):
        col.register_collective_backend("GLOO", GlooOracleGroup)
    log = {}
    col.init_collective_group(world, rank, backend="gloo", group_name="default")
    log["rank"] = col.get_rank()
    log["size"] = col.get_collective_group_size()
    # test_allreduce.py (cpu): numpy and torch operands, sizes 2..2^20, results in place
    for size in (2, 2 ** 10, 2 ** 20):
        a = np.ones(size, dtype=np.float32) * (rank + 1)
        col.allreduce(a)
        assert np.all(a == sum(range(1, world + 1))), size
        t = torch.ones(size, dtype=torch.float32) * (rank + 1)
        col.allreduce(t, op=col.ReduceOp.MAX)
        assert torch.all(t == world)
    # reduce leaves non-root untouched (Q9)
    t = torch.ones(5) * (rank + 1)
    col.reduce(t, dst_rank=1)
    assert torch.all(t == (sum(range(1, world + 1)) if rank == 1 else rank + 1))
    # broadcast / allgather / reducescatter (gloo emulation overwrites the inputs, Q13)
    t = torch.ones(5) * (rank + 1)
    col.broadcast(t, src_rank=1)
    assert torch.all(t == 2)
    outs = [torch.zeros(4) for _ in range(world)]
    col.allgather(outs, torch.ones(4) * (rank + 1))
    assert all(torch.all(outs[p] == p + 1) for p in range(world))
    ins = [torch.ones(4) * (rank + 1) * (i + 1) for i in range(world)]
    out = torch.zeros(4)
    col.reducescatter(out, ins)
    assert torch.all(out == 3 * (rank + 1)) and torch.all(ins[0] == 3)
    if rank == 0:
        col.send(torch.arange(6, dtype=torch.float32), 1)
    else:
        r = torch.zeros(6)
        col.recv(r, 0)
        assert torch.equal(r, torch.arange(6, dtype=torch.float32))
    # errors
    for fn, exc in ((lambda: col.send(torch.ones(1), rank), RuntimeError),
                    (lambda: col.send(torch.ones(1), 7), ValueError),
                    (lambda: col.allgather([torch.ones(1)], torch.ones(1)), RuntimeError),
                    (lambda: col.init_collective_group(world, rank, backend="gloo"), RuntimeError),
                    (lambda: col.allreduce(torch.ones(1), "nope"), RuntimeError)):
        try:
            fn()
            raise AssertionError("expected an exception")
        except exc:
            pass
    col.barrier()
    col.destroy_collective_group()
    assert col.get_rank() == -1 and not col.is_group_initialized("default")
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write(repr(log))


def test_config1_allreduce_world2_gloo_through_the_api():
    with tempfile.TemporaryDirectory() as store_dir, tempfile.TemporaryDirectory() as out_dir:
        mp.spawn(_gloo_worker, args=(2, store_dir, out_dir), nprocs=2, join=True)
        logs = [eval(open(os.path.join(out_dir, f"ok{r}")).read()) for r in range(2)]
        assert [l["rank"] for l in logs] == [0, 1] and all(l["size"] == 2 for l in logs)


def test_backend_lookup_and_registry():
    from ray_b200 import collective as col
    from ray_b200.collective import types
    from ray_b200.collective.registry import BackendRegistry

    assert col.Backend("b200") == "B200" and col.Backend("NCCL") == "NCCL" and col.Backend("torch_gloo") == "GLOO"
    with pytest.raises(ValueError, match="Unrecognized backend"):
        col.Backend("unrecognized")
    reg = BackendRegistry()
    with pytest.raises(TypeError):
        reg.put("x", object)
    with pytest.raises(ValueError, match="not registered"):
        reg.get("x")
    assert not reg.check("x") and not reg.is_registered("X")

    class Dummy(col.BaseGroup):
        @classmethod
        def backend(cls):
            return "DUMMY"

        @classmethod
        def check_backend_availability(cls):
            return True

        def allreduce(self, tensor, allreduce_options=None):
            self.seen = allreduce_options.reduceOp

        barrier = reduce = allgather = broadcast = reducescatter = send = recv = lambda self, *a, **k: None

    reg.put("dummy", Dummy)
    with pytest.raises(ValueError, match="already registered"):
        reg.put("DUMMY", Dummy)
    assert reg.check("Dummy") and reg.get("dummy") is Dummy
    # registering through the public entry point makes Backend.<NAME> resolvable (Q3)
    col.register_collective_backend("unit_dummy", Dummy)
    assert types.Backend("unit_dummy") == "UNIT_DUMMY"
    # Q1: collective.allreduce hands the options CLASS to the backend with the op set on it
    mgr = col.GroupManager()
    with col.use_manager(mgr):
        col.init_collective_group(1, 0, backend="unit_dummy", group_name="g")
        col.allreduce(torch.ones(1), "g", col.ReduceOp.MAX)
        assert mgr.get_group_by_name("g").seen == col.ReduceOp.MAX
        assert types.AllReduceOptions.reduceOp == col.ReduceOp.MAX
        col.destroy_collective_group("g")
    types.AllReduceOptions.reduceOp = col.ReduceOp.SUM


def test_b200_backend_fails_loudly_without_a_gpu():
    from ray_b200 import collective as col

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert col.B200Group.check_backend_availability() is False
    with col.use_manager(col.GroupManager()):
        with pytest.raises(RuntimeError, match="not available"):
            col.init_collective_group(2, 0, backend="b200", group_name="x")


def test_stores_roundtrip_and_timeout():
    from ray_b200.store import DictStore, FileStore

    with tempfile.TemporaryDirectory() as d:
        for st in (DictStore(), FileStore(d)):
            with pytest.raises(TimeoutError):
                st.get("missing", timeout_s=0.05)
            threading.Timer(0.05, lambda s=st: s.set("k/1?x", b"\x00\x01payload")).start()
            assert st.get("k/1?x", timeout_s=5) == b"\x00\x01payload"
            st.delete("k/1?x")
            with pytest.raises(TimeoutError):
                st.get("k/1?x", timeout_s=0.01)


def test_torch_dist_store_adapter():
    import torch.distributed as dist

    from ray_b200.store import TorchDistStore

    st = TorchDistStore(dist.HashStore())
    st.set("a", b"xyz")
    assert st.get("a") == b"xyz"
    with pytest.raises(TimeoutError):
        st.get("missing", timeout_s=0.0)
    st.delete("a")
    with pytest.raises(TimeoutError):
        st.get("a", timeout_s=0.05)

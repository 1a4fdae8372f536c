"""BASELINE config 3: GPU tensor ping-pong between two workers (one process per GPU), 1 KB .. 1 GB.

Follows the reference microbenchmark's raw-communicator loop (``exec_nccl_gpu`` in
release/microbenchmark/experimental/compiled_graph_gpu_microbenchmark.py:380-408): rank 0 sends a
fp16 tensor, rank 1 receives it and sends it back.  The reference's ``_NcclGroup`` synchronises the
host after each recv; the B200 communicator hands tensors over by CUDA event instead (rank 0
waits once per round trip, rank 1 never).  Reports the round-trip time per size
for the B200 communicator and, as comparator, for torch.distributed NCCL send/recv.

    python scripts/channel_pingpong.py            # spawns 2 processes on GPUs 0 and 1
"""
import json
import os
import sys
import tempfile
import time

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES = [1 << s for s in range(10, 31, 2)]


def worker(rank, store_dir, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from ray_b200.channel import B200Communicator, TorchTensorAcceleratorChannel
    from ray_b200.store import FileStore

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    comm = B200Communicator(2, "pingpong", None, ["w0", "w1"], None, False, FileStore(store_dir), rank,
                            inbox_bytes=64 << 20, staging_bytes=16 << 20)
    comm.initialize(rank)
    dist.init_process_group("nccl", init_method=f"file://{store_dir}/nccl_rdzv", rank=rank, world_size=2,
                            device_id=dev)
    alloc = lambda shape, dtype: torch.empty(shape, dtype=dtype, device=dev)  # noqa: E731
    rows = []
    for nbytes in SIZES:
        numel = nbytes // 2
        x = torch.ones(numel, dtype=torch.float16, device=dev)
        iters = 200 if nbytes <= (1 << 20) else (30 if nbytes <= (64 << 20) else 8)
        res = {"bytes": nbytes}
        # _static_shape=True, _direct_return=True: metadata travels once per channel
        fwd = TorchTensorAcceleratorChannel(comm, 0, [1], static_shape=True, direct_return=True)
        back = TorchTensorAcceleratorChannel(comm, 1, [0], static_shape=True, direct_return=True)
        for name in ("b200_raw", "b200_channel", "nccl"):
            def one():
                if name == "b200_raw":
                    # event mode (default): no host sync per op.  Rank 1 forwards the tensor it
                    # received on the same stream, so its host never waits; rank 0 waits once per
                    # round trip (that wait IS the round-trip time being measured).
                    if rank == 0:
                        comm.send(x, 1)
                        comm.wait(comm.recv((numel,), torch.float16, 1, alloc))
                    else:
                        comm.send(comm.recv((numel,), torch.float16, 0, alloc), 0)
                elif name == "b200_channel":
                    if rank == 0:
                        fwd.write(x)
                        comm.wait(back.read())
                    else:
                        back.write(fwd.read())
                else:
                    if rank == 0:
                        dist.send(x, 1)
                        dist.recv(x, 1)
                    else:
                        dist.recv(x, 0)
                        dist.send(x, 0)
                    torch.cuda.current_stream().synchronize()

            for _ in range(5):
                one()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(iters):
                one()
            torch.cuda.synchronize()
            res[name + "_rtt_us"] = (time.perf_counter() - t0) / iters * 1e6
        res["b200_one_way_gbs"] = nbytes / (res["b200_raw_rtt_us"] / 2 * 1e-6) / 1e9
        res["nccl_one_way_gbs"] = nbytes / (res["nccl_rtt_us"] / 2 * 1e-6) / 1e9
        rows.append(res)
        if rank == 0:
            print(json.dumps(res), flush=True)
    dist.barrier()
    if rank == 0:
        json.dump(rows, open(out_path, "w"), indent=1)
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "channel_pingpong.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(worker, args=(d, out), nprocs=2, join=True)

"""Host-side ceiling of the end-to-end leg: N ranks (one per GPU) copy an observation-sized block
device -> pinned host (and bids host -> device) concurrently with cudaMemcpyAsync, nothing else.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/host_ceiling.py

Prints one JSON line (rank 0): aggregate GB/s for D2H alone and for D2H + H2D together, and the
keyword-auction-steps/s those rates would allow at 20 B / 14 B per unit out + 4 B in."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:  # noqa: BLE001
    pass


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


out = {}
for name, d2h_bytes, h2d_bytes in (("int32_arrays", 4096 * 100 * 20, 4096 * 100 * 4), ("compact_rows", 4096 * 1424, 4096 * 100 * 4)):
    d_src = torch.zeros(d2h_bytes, dtype=torch.uint8, device=dev)
    h_dst = torch.zeros(d2h_bytes, dtype=torch.uint8).pin_memory()
    h_src = torch.zeros(h2d_bytes, dtype=torch.uint8).pin_memory()
    d_dst = torch.zeros(h2d_bytes, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream(device=dev)
    for both in (False, True):
        n = 300
        for it in range(2):
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                h_dst.copy_(d_src, non_blocking=True)
                if both:
                    with torch.cuda.stream(s2):
                        d_dst.copy_(h_src, non_blocking=True)
                torch.cuda.synchronize(dev)  # a step returns host data: one sync per step
            barrier()
            dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        key = name + ("_d2h+h2d" if both else "_d2h")
        out[key] = {"ms_per_step": dt / n * 1e3, "aggregate_GBps": world * (d2h_bytes + (h2d_bytes if both else 0)) * n / dt / 1e9,
                    "units_per_s_allowed": world * 4096 * 100 * n / dt}
if rank == 0:
    print(json.dumps({"n_gpus": world, "copies": out}))
if world > 1:
    dist.destroy_process_group()

#!/usr/bin/env python
"""Host <-> device copy ceiling of this box with all ranks copying at once: every rank moves `MB` megabytes up and the same
amount down concurrently (two streams, pinned buffers) -- what the end-to-end (host-buffer) residual of bench.py does with the
state.  Prints the aggregate and per-GPU rate; the e2e figure of N GPUs cannot exceed state_bytes / aggregate rate.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py [MB]"""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = mb * 1024 * 1024 // 8
    hu, hd = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
    du, dd = torch.empty(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for mode in ("both", "h2d", "d2h"):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_event(e0); s2.wait_event(e0)
            if mode in ("both", "h2d"):
                with torch.cuda.stream(s1):
                    du.copy_(hu, non_blocking=True)
            if mode in ("both", "d2h"):
                with torch.cuda.stream(s2):
                    hd.copy_(dd, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = min(best, ms)
        out[mode + "_ms"] = best
        out[mode + "_gbs_per_gpu_per_direction"] = mb / 1024 / (best * 1e-3)
        out[mode + "_gbs_aggregate_per_direction"] = world * mb / 1024 / (best * 1e-3)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "MiB_per_gpu_per_direction": mb, **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

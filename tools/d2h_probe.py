#!/usr/bin/env python
"""Host-link probe for the end-to-end numbers: every rank copies a device buffer into its own pinned host buffer
at the same time (what bench.py's e2e leg does with the CSR values), alone and all together.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/d2h_probe.py

Prints the per-rank and aggregate GB/s, device to host and host to device, plus what the box says about its topology."""
import os
import subprocess
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = (1 << 30) // 8 * 2                       # 2 GiB
    dev = torch.ones(n, dtype=torch.float64, device="cuda")
    host = torch.empty(n, dtype=torch.float64).pin_memory()
    host.fill_(0.0)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, reps=4):
        fn(); sync_all()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return reps * n * 8 / dt / 1e9

    res = {}
    for name, fn in (("d2h", lambda: host.copy_(dev, non_blocking=True)), ("h2d", lambda: dev.copy_(host, non_blocking=True))):
        # all ranks at once
        sync_all()
        together = timed(fn)
        # one rank at a time
        alone = 0.0
        for r in range(world):
            sync_all()
            if r == rank:
                alone = timed(fn)
            sync_all()
        res[name] = (alone, together)
    t = torch.tensor([res["d2h"][0], res["d2h"][1], res["h2d"][0], res["h2d"][1]], dtype=torch.float64, device="cuda")
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    if rank == 0:
        rows = [x.tolist() for x in allt]
        print("rank  d2h alone  d2h together  h2d alone  h2d together  (GB/s)")
        for r, x in enumerate(rows):
            print(f"{r:4d}  {x[0]:9.1f}  {x[1]:12.1f}  {x[2]:9.1f}  {x[3]:12.1f}")
        print(f"aggregate together: d2h {sum(x[1] for x in rows):.1f} GB/s, h2d {sum(x[3] for x in rows):.1f} GB/s; "
              f"host cpus visible {len(os.sched_getaffinity(0))}")
        for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "ls /sys/devices/system/node/ | grep node; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c"]):
            try:
                print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout)
            except Exception as e:      # noqa: BLE001
                print(cmd, e)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

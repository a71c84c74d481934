#!/usr/bin/env python3
"""Bare host->device bandwidth probe per GPU count: every rank copies a pinned
800 MB buffer to its GPU, all ranks at once, with and without binding the rank to its GPU's NUMA node before the
buffer is allocated.  Launch with torchrun; rank 0 prints one JSON line (and `nvidia-smi topo -m` to stderr).

    torchrun --nproc-per-node N tools/h2d_probe.py
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import bind_to_gpu_numa_node  # noqa: E402


def measure(dev, nbytes, reps=8):
    h = torch.empty(nbytes // 4, dtype=torch.float32, pin_memory=True)
    h.fill_(1.0)
    d = torch.empty_like(h, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dist.barrier()
    return nbytes * reps / dt / 1e9


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nbytes = 800_000_000
    unbound = measure(dev, nbytes)
    numa = bind_to_gpu_numa_node(local)
    bound = measure(dev, nbytes)
    rows = [None] * world
    dist.all_gather_object(rows, {"rank": rank, "unbound_GBps": unbound, "numa_bound_GBps": bound, "numa": numa})
    if rank == 0:
        try:
            sys.stderr.write(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout)
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("nvidia-smi topo failed: %r\n" % (e,))
        print(json.dumps({"world": world, "bytes_per_copy": nbytes, "aggregate_unbound_GBps": sum(r["unbound_GBps"] for r in rows),
                          "aggregate_numa_bound_GBps": sum(r["numa_bound_GBps"] for r in rows), "ranks": rows,
                          "host_cpus": os.cpu_count()}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

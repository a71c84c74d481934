#!/usr/bin/env python3
"""Where does a psd() readout after EVERY 200e6-sample step go at N GPUs?  (VERDICT r01 weak #8.)
torchrun --nproc-per-node N tools/readout_probe.py : per-step time of the channel-sharded group with a
sspsd_group_psd_all readout per step, with and without the NVML clock sampler thread of bench.py running in
every rank; SSPSD_GROUP_TRACE=1 makes the library print its per-phase host times to stderr."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SSPSD_GROUP_TRACE", "1")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import N_FFT, SAMPLES_PER_STEP, ClockSampler  # noqa: E402
from stabilizer_stream_b200 import Group, MergeOpts, ShardMode, _lib  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream(dev)
    x = (torch.rand(SAMPLES_PER_STEP, device=dev) - 0.5) * (12 ** 0.5)
    out = {"world": world}
    for sampler_on in (True, False, True, False):
        ids = [Group.unique_id() if (rank == 0 and world > 1) else None]
        if world > 1:
            dist.broadcast_object_list(ids, src=0)
        g = Group(N_FFT, rank=rank, n_ranks=world, unique_id=ids[0], device=local, mode=ShardMode.CHANNELS,
                  stream=stream.cuda_stream or 1)
        for _ in range(3):
            g.process_raw(rank, x.data_ptr(), SAMPLES_PER_STEP, _lib.MEM_DEVICE)
            g.psd_all(world, MergeOpts())
        dist.barrier()
        torch.cuda.synchronize()
        smp = ClockSampler(local)
        if sampler_on:
            smp.start()
        steps = 20
        t0 = time.perf_counter()
        for _ in range(steps):
            g.process_raw(rank, x.data_ptr(), SAMPLES_PER_STEP, _lib.MEM_DEVICE)
            g.psd_all(world, MergeOpts())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if sampler_on:
            smp.stop()
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        key = "sampler_on" if sampler_on else "sampler_off"
        out.setdefault(key, []).append(float(t.item()) / steps * 1e3)
        del g
    if rank == 0:
        out["GSps_sampler_on"] = world * SAMPLES_PER_STEP / (min(out["sampler_on"]) * 1e-3) / 1e9
        out["GSps_sampler_off"] = world * SAMPLES_PER_STEP / (min(out["sampler_off"]) * 1e-3) / 1e9
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

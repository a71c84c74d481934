#!/usr/bin/env python3
"""One rank of a multi-process sspsd_group (ncclCommInitRank inside the library): launched by torchrun, checks
the group's result against the sequential cascade on rank 0 and prints one JSON line.

    torchrun --nproc-per-node N tools/group_rank.py --mode time|time_preset|channels [--total 6e6] [--n 512]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="time")
    ap.add_argument("--total", type=float, default=6e6)
    ap.add_argument("--n", type=int, default=512)
    args = ap.parse_args()
    import stabilizer_stream_b200 as sp
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")  # only to hand the 128-byte NCCL id around
    torch.cuda.set_device(local)
    ids = [sp.Group.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    n, total = args.n, int(args.total)
    preset = args.mode == "time_preset"
    out = {"world": world, "mode": args.mode}
    if args.mode.startswith("time"):
        g = sp.Group(n, rank=rank, n_ranks=world, unique_id=ids[0], device=local, mode=sp.ShardMode.TIME)
        if preset:
            g.set_detrend(sp.Detrend.MEAN)
            g.set_avg(sp.AvgOpts(999, 2 ** 32 - 2))
        g.time_plan(total)
        g.time_process_noise(0, sp.Source.SEED)
        g.time_finish()
        p, b = g.psd(0)
        if rank == 0:
            seq = sp.PsdCascade(n, device=local)
            if preset:
                seq.set_detrend(sp.Detrend.MEAN)
                seq.set_avg(sp.AvgOpts(999, 2 ** 32 - 2))
            src = sp.Source.noise(0, device=local)
            pos = 0
            while pos < total:
                m = min(1 << 26, total - pos)
                seq.process_source(src, m)
                pos += m
            ps, bs = seq.psd()
            same = [(k.count, k.pending, k.processed, k.bins) for k in b] == [(k.count, k.pending, k.processed, k.bins) for k in bs]
            head = 2 if preset else 0
            rel = float(np.max(np.abs(p[head:] - ps[head:]) / np.maximum(ps[head:], 1e-30))) if p.size == ps.size else float("inf")
            out.update(breaks_equal=same, max_rel_diff_vs_sequential=rel, ok=bool(same and rel < 5e-5), reduce=g.info()["reduce"],
                       n_local=g.time_chunk(0).n_local)
        else:
            assert p.size == 0 and b == []
    else:
        n_ch = 2 * world + 1
        g = sp.Group(n, rank=rank, n_ranks=world, unique_id=ids[0], device=local, mode=sp.ShardMode.CHANNELS)
        xs = []
        for c in range(n_ch):
            rng = np.random.default_rng(900 + c)
            xs.append(((rng.random(150 * n + 31 * c, dtype=np.float32) - np.float32(0.5)) * np.float32(1 + c)).astype(np.float32))
            g.process(c, xs[c])      # a no-op on the ranks that do not own channel c
        res = g.psd_all(n_ch)
        if rank == 0:
            ok = True
            for c in range(n_ch):
                seq = sp.PsdCascade(n, device=local)
                seq.process(xs[c])
                ps, bs = seq.psd()
                p, b = res[c]
                ok = ok and p.size == ps.size and [k.count for k in b] == [k.count for k in bs] and \
                    bool(np.max(np.abs(p - ps) / np.maximum(ps, 1e-30)) < 1e-5)
            out.update(ok=bool(ok), channels=n_ch)
        else:
            assert all(p.size == 0 for p, _ in res)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

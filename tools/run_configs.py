#!/usr/bin/env python3
"""Run the BASELINE.json parity/throughput configurations that are not the bench.py workload and print one
JSON line per config (rank 0).  Launch with python (1 GPU) or torchrun (N GPUs, config 5 only).

  --config 1   raw f32, single stage N=4096 Hann (Psd handle), 2^28 samples
  --config 2   PsdCascade N=4096, 200e6 samples per call, default options and the `psd` binary's preset
               (Detrend::Mean, AvgOpts{limit 999, count u32::MAX-1}, bin/psd.rs:60-78); bench.py measures the
               default-options case with the full contract, this adds the preset beside it
  --config 3   AdcDac frames (22 batches) -> decode + loss + 4 cascades, 2^20 frames
  --config 5   4.8e9-sample capture, N=4096, time-chunked over WORLD_SIZE GPUs (checked against the
               sequential 1-GPU run of the same stream when --check is given); --preset uses the binary's
               options (EWMA over the global segment order)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

BLOCK = 1 << 26


def noise_block(b, dev):
    """Block b of the synthetic stream: uniform, zero mean, unit variance; counter-based per block so that
    every rank generates identical samples for the same stream positions."""
    g = torch.Generator(device=dev).manual_seed(0x7654321 + b)
    return (torch.rand(BLOCK, device=dev, generator=g) - 0.5) * (12 ** 0.5)


def feed_stream(lo, hi, sink, dev):
    pos = lo
    while pos < hi:
        b = pos // BLOCK
        blk = noise_block(b, dev)
        a = pos - b * BLOCK
        e = min(BLOCK, hi - b * BLOCK)
        sink(blk[a:e])
        pos = b * BLOCK + e


def config1(dev):
    from oracle import binding as orc
    from stabilizer_stream_b200 import Psd
    n, total = 4096, 1 << 28
    x = torch.cat([noise_block(b, dev) for b in range(total // BLOCK)])
    s = Psd(n)
    ybuf = torch.empty(total // 8 + n, device=dev)
    s.process(x, out=ybuf)   # first pass: allocates the handle's buffers at their final size
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y = s.process(x, out=ybuf)   # timed: second pass over the same 2^28 samples (the stream simply continues)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    p = s.spectrum() / s.gain()
    flat = bool(np.all(np.abs(p * 0.5 - 1.0) < 10.0 / np.sqrt(s.count())))
    xs = x[:1 << 24].cpu().numpy()
    o = orc.Stage(n)
    t0 = time.perf_counter()
    o.process(xs)
    dtc = time.perf_counter() - t0
    return {"config": 1, "samples": total, "gpu_MSps_single_stage_incl_decimated_output": total / dt / 1e6,
            "cpu_port_MSps_1core": xs.size / dtc / 1e6, "segments": s.count(), "flat_10sigma": flat,
            "decimated_len": int(y.numel())}


def config3(dev):
    from frames_util import make_frames, oracle_decode_stream
    from oracle import binding as orc
    from stabilizer_stream_b200 import Detrend, FrameDecoder, Loss, PsdCascade
    n, batches, nfr = 4096, 22, 1 << 20
    small, flen, stride, _ = make_frames(1, batches, 4096, seed=3, drop_every=1009, start_seq=0xFFFFFFFF - 9 * 1009 * batches)
    reps = nfr // 4096
    # the big stream repeats the small one (sequence numbers then jump back: counted as huge wrapping gaps,
    # exactly like Loss::update would) -- what matters is bit-exact agreement with the oracle on the same bytes
    data = small * reps
    fr = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    cas = [PsdCascade(n) for _ in range(4)]
    for c in cas:
        c.set_detrend(Detrend.MIDPOINT)
    dec = FrameDecoder()
    loss = Loss()
    dec.process_frames(cas, fr, flen, Loss())   # warm-up at full size: decoder and cascade buffers reach their final size
    for c in cas:
        c.sync()
        c.reset()                                # keeps the buffers (the GUI's Cmd::Reset)
        c.set_detrend(Detrend.MIDPOINT)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = dec.process_frames(cas, fr, flen, loss)
    for c in cas:
        c.sync()
    dt = time.perf_counter() - t0
    # decode + loss only, traces written to device buffers (K1 alone: 2 B read + 4 B written per sample)
    outs = [torch.empty(nfr * batches * 8, device=dev) for _ in range(4)]
    dec.decode_device(fr, flen, outs, Loss())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dec.decode_device(fr, flen, outs, Loss())
    e1.record()
    torch.cuda.synchronize()
    dec_ms = e0.elapsed_time(e1) / 5
    st, nf, lo, want = oracle_decode_stream(orc, data, flen, stride, nfr)
    ok_loss = (loss.received, loss.dropped, loss.seq) == (lo.received, lo.dropped, lo.seq)
    ok_tr = all(np.array_equal(outs[t][:4096 * batches * 8].cpu().numpy().view(np.uint32),
                               np.concatenate(want[t][:4096]).view(np.uint32)) for t in range(4))
    samples = 4 * info.samples_per_trace
    return {"config": 3, "frames": nfr, "bytes": len(data), "trace_samples_total": int(samples),
            "gpu_MSps_decode_plus_4_cascades": samples / dt / 1e6, "frame_GBps": len(data) / dt / 1e9,
            "decode_only_ms": dec_ms, "decode_only_GBps_read_plus_written": (len(data) + samples * 4) / dec_ms / 1e6,
            "decode_only_GSps": samples / dec_ms / 1e6, "traces_bit_exact_first_4096_frames": bool(ok_tr),
            "loss_bit_exact": bool(ok_loss), "received": int(loss.received), "dropped": int(loss.dropped)}


def config2(dev):
    from oracle import binding as orc
    from stabilizer_stream_b200 import AvgOpts, Detrend, MergeOpts, PsdCascade
    n, per, steps = 4096, 200_000_000, 10
    x = torch.cat([noise_block(b, dev) for b in range(4)])[:per]
    out = {"config": 2, "samples_per_call": per, "calls": steps}
    for name, det, avg in (("default", Detrend.NONE, None), ("binary_preset", Detrend.MEAN, AvgOpts(limit=999, count=2 ** 32 - 2))):
        c = PsdCascade(n, device=dev.index)
        c.set_detrend(det)
        if avg is not None:
            c.set_avg(avg)
        for _ in range(3):
            c.process(x)
        c.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            c.process(x)
        p, br = c.psd(MergeOpts())
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        # (the same 200e6-sample block is resubmitted every call, as in bench.py, so the averaged spectrum does
        # not converge like 1/sqrt(count): no flatness check here; tests/ and config 5 cover that)
        out[name] = {"GSps": per * steps / ms / 1e6, "ms_per_call": ms / steps, "counts": [k.count for k in reversed(br)],
                     "avg": [k.avg for k in reversed(br)]}
    # parity of the preset against the CPU restatement on a bounded prefix (the full stream takes minutes on a core)
    m = 30_000_000
    xs = x[:m].cpu().numpy()
    c = PsdCascade(n, device=dev.index)
    c.set_detrend(Detrend.MEAN)
    c.set_avg(AvgOpts(limit=999, count=2 ** 32 - 2))
    c.process(x[:m])
    o = orc.Cascade(n, 1)
    o.set_detrend(3)
    o.set_avg(999, 2 ** 32 - 2)
    o.process(xs)
    p, br = c.psd(MergeOpts())
    po, bo = o.psd()
    out["preset_max_rel_vs_cpu_30e6"] = float(np.max(np.abs(p - po)[2:] / np.maximum(po[2:], 1e-30)))
    out["preset_counts_equal"] = [k.count for k in br] == [k.count for k in bo]
    return out


def config5(dev, total, n_local, check, preset=False):
    import torch.distributed as dist
    from stabilizer_stream_b200 import MergeOpts, PsdCascade, multi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    d = dist if world > 1 else None
    n = 4096
    from stabilizer_stream_b200 import AvgOpts, Detrend
    avg = AvgOpts(limit=999, count=2 ** 32 - 2) if preset else None
    c = PsdCascade(n, device=dev.index)
    if preset:
        c.set_detrend(Detrend.MEAN)
    # the rank's share of the capture is generated into HBM first (untimed): feed range of the plan
    plan = multi.plan_time_chunks(total, world, n, 1, n_local)[rank]
    lo, hi = plan["feed_lo"], plan["feed_hi"]
    parts = []
    feed_stream(lo, hi, lambda t: parts.append(t.clone()), dev)
    xs = torch.cat(parts)
    del parts
    # warm-up: one untimed pass of the whole pipeline on the same handle and stream (NCCL connections, the
    # handle's stage / tail buffers at their final size, kernels loaded), then reset() -- which keeps buffers
    multi.time_chunked_psd(c, lambda a, b, sink: sink(xs), total, n, d, 1, n_local, str(dev), None, avg)
    c.sync()
    c.reset()
    if preset:
        c.set_detrend(Detrend.MEAN)
    if d is not None:
        d.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gen_s = [0.0]

    def feed(a, b, sink):
        assert a == lo and b == hi
        sink(xs)

    tim = {}
    root = multi.time_chunked_psd(c, feed, total, n, d, 1, n_local, str(dev), tim, avg)
    torch.cuda.synchronize()
    if d is not None:
        d.barrier()
    dt = time.perf_counter() - t0
    if rank != 0:
        return None
    p, b = root.psd(MergeOpts())
    counts = [k.count for k in reversed(b)]
    want = [min(s[1], multi.stage_avg(avg, i) + 1) for i, s in enumerate(multi.stream_state(total, n, n // 2, multi.DRAIN[1]))]
    flat = all(bool(np.all(np.abs(p[k.start:k.start + len(k.bins)] * 0.5 - 1.0) < 10.0 / np.sqrt(k.count)))
               for k in b if k.include and k.count >= 20)
    out = {"config": 5, "preset": bool(preset), "samples": total, "world": world, "n_local": n_local, "stage_counts": counts,
           "counts_match_closed_form": counts == want, "flat_10sigma": flat, "wall_s": dt,
           "MSps": total / dt / 1e6, "phases_s_rank0": tim, "halo_overhead": (hi - lo) * world / total - 1 if world > 1 else 0.0,
           "bins": int(p.size)}
    if check and world > 1:
        seq = PsdCascade(n, device=dev.index)
        if preset:
            seq.set_detrend(Detrend.MEAN)
            seq.set_avg(avg)
        feed_stream(0, total, seq.process, dev)
        ps, bs = seq.psd(MergeOpts())
        out["max_rel_diff_vs_sequential"] = float(np.max(np.abs(p - ps) / np.maximum(ps, 1e-30)))
        o = MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True)
        pk, bk = root.psd(o)
        psk, bsk = seq.psd(o)
        out["per_stage_max_rel"] = {str(k.decimation): float(np.max(np.abs(pk[k.start:k.start + len(k.bins)] - psk[k.start:k.start + len(k.bins)])
                                                                      / np.maximum(psk[k.start:k.start + len(k.bins)], 1e-30)))
                                    for k in bk if k.count}
        out["breaks_equal"] = [(k.count, k.pending, k.processed) for k in b] == [(k.count, k.pending, k.processed) for k in bs]
    return out


def config5_group(dev, total, n_local, check, preset=False):
    """config 5 through the library's own multi-GPU plumbing (sspsd_group_time_*): planner, seek / window, the
    ONE NCCL reduction and the deep stages on rank 0 all run inside libsspsd.so; Python only feeds the samples."""
    import torch.distributed as dist
    from stabilizer_stream_b200 import AvgOpts, Detrend, Group, MergeOpts, PsdCascade, ShardMode, _lib, multi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    n = 4096
    avg = AvgOpts(limit=999, count=2 ** 32 - 2) if preset else None
    ids = [Group.unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(ids, src=0)
    g = Group(n, rank=rank, n_ranks=world, unique_id=ids[0], device=dev.index, mode=ShardMode.TIME)
    if preset:
        g.set_detrend(Detrend.MEAN)
        g.set_avg(avg)
    g.time_plan(total, n_local)
    ch = g.time_chunk(rank)
    lo, hi = ch.feed_lo, ch.feed_hi
    parts = []
    feed_stream(lo, hi, lambda t: parts.append(t.clone()), dev)   # the rank's share, generated into HBM first (untimed)
    xs = torch.cat(parts)
    del parts
    torch.cuda.synchronize()

    def one_pass():
        g.time_plan(total, n_local)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L = _lib
        L.check(L.lib().sspsd_group_time_process_f32(g._h, rank, xs.data_ptr(), xs.numel(), L.MEM_DEVICE))
        t1 = time.perf_counter()
        g.time_finish()
        p, b = g.psd(0)
        t2 = time.perf_counter()
        if world > 1:
            dist.barrier()
        return p, b, time.perf_counter() - t0, {"enqueue": t1 - t0, "finish+psd": t2 - t1}

    one_pass()                      # warm-up: NCCL connections, buffers at their final size
    p, b, dt, tim = one_pass()
    if rank != 0:
        return None
    counts = [k.count for k in reversed(b)]
    want = [min(s[1], multi.stage_avg(avg, i) + 1) for i, s in enumerate(multi.stream_state(total, n, n // 2, multi.DRAIN[1]))]
    flat = all(bool(np.all(np.abs(p[k.start:k.start + len(k.bins)] * 0.5 - 1.0) < 10.0 / np.sqrt(k.count)))
               for k in b if k.include and k.count >= 20)
    out = {"config": 5, "orchestration": "libsspsd sspsd_group_time_*", "preset": bool(preset), "samples": total, "world": world,
           "n_local": ch.n_local, "reduce": g.info()["reduce"], "stage_counts": counts, "counts_match_closed_form": counts == want,
           "flat_10sigma": flat, "wall_s": dt, "MSps": total / dt / 1e6, "phases_s_rank0": tim,
           "halo_overhead": (hi - lo) * world / total - 1 if world > 1 else 0.0, "bins": int(p.size)}
    if check and world > 1:
        seq = PsdCascade(n, device=dev.index)
        if preset:
            seq.set_detrend(Detrend.MEAN)
            seq.set_avg(avg)
        feed_stream(0, total, seq.process, dev)
        ps, bs = seq.psd(MergeOpts())
        out["max_rel_diff_vs_sequential"] = float(np.max(np.abs(p - ps) / np.maximum(ps, 1e-30)))
        out["breaks_equal"] = [(k.count, k.pending, k.processed) for k in b] == [(k.count, k.pending, k.processed) for k in bs]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--total", type=float, default=4.8e9)
    ap.add_argument("--n-local", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--preset", action="store_true")
    ap.add_argument("--python-orchestration", action="store_true",
                    help="config 5 through stabilizer_stream_b200/multi.py (round-1 path) instead of sspsd_group_*")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    r = {1: lambda: config1(dev), 2: lambda: config2(dev), 3: lambda: config3(dev),
         5: lambda: (config5 if a.python_orchestration else config5_group)(dev, int(a.total), a.n_local, a.check, a.preset)}[a.config]()
    if r is not None:
        print(json.dumps(r))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

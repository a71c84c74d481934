"""sspsd_group_*: multi-GPU partitioning inside the C ABI (north_star item 5; VERDICT r01 "missing" #1).

On a one-GPU box several ranks share device 0 (the group then reduces with its direct peer-load kernel instead
of NCCL, everything else -- planner, seek / window, exchange layout, deep stages on rank 0 -- is the same code);
with >= 2 GPUs the same tests also run over ncclCommInitAll, and tools/group_rank.py exercises the
one-process-per-GPU path (ncclCommInitRank) under torchrun."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import uniform_noise

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def breaks_tuple(b):
    return (b.start, b.include, b.count, b.avg, b.bins.start, b.bins.stop, b.fft_size, b.decimation, b.pending, b.processed)


def device_lists(world):
    import torch
    out = [[0] * world]                               # ranks share GPU 0: direct reduction
    if torch.cuda.device_count() >= world and world > 1:
        out.append(list(range(world)))                # one GPU per rank: ncclCommInitAll
    return out


def check_against_sequential(sp, oracle, g, x, n, hbf, avg, det):
    import torch
    p, b = g.psd(0)
    seq = sp.PsdCascade(n, hbf=sp.Hbf(hbf))
    if avg is not None:
        seq.set_avg(avg)
    seq.set_detrend(sp.Detrend(det))
    seq.process(torch.from_numpy(x).cuda())
    ps, bs = seq.psd()
    assert [breaks_tuple(v) for v in b] == [breaks_tuple(v) for v in bs]
    head = 2 if det >= 2 else 0
    np.testing.assert_allclose(p[head:], ps[head:], rtol=3e-5, atol=1e-6 * float(np.median(ps)))
    # and against the CPU oracle: all stages, all bins
    o = oracle.Cascade(n, hbf)
    if avg is not None:
        o.set_avg(avg.limit, avg.count)
    o.set_detrend(det)
    o.process(x)
    pk, bk = g.psd(0, sp.MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
    pok, bok = o.psd(True, 0, True)
    assert [breaks_tuple(v) for v in bk] == [v.as_tuple() for v in bok]
    for bi in bk:
        if bi.count:
            sl = slice(bi.start + head, bi.start + len(bi.bins))
            floor = 1e-5 * np.median(pok[sl])
            assert np.max((np.abs(pk[sl] - pok[sl]) - floor) / pok[sl]) < 1e-4, "stage dec=%d" % bi.decimation


@pytest.mark.parametrize("n,total,world,k,hbf,avg,det", [
    (512, 3_000_017, 4, 3, 1, None, 0),
    (4096, 30_000_000, 3, 2, 1, None, 3),
    (512, 2_500_000, 2, 0, 0, None, 1),            # n_local = 0: planner's choice
    (512, 3_000_017, 4, 3, 1, (999, 2 ** 32 - 2), 3),   # the psd binary's preset: EWMA over the GLOBAL segment order
    (256, 1_500_000, 5, 3, 1, (0, 64), 0),
    (64, 700_001, 1, 2, 1, None, 0),                # a group of one rank degenerates to the plain cascade
])
def test_time_chunked_group_equals_sequential(oracle, n, total, world, k, hbf, avg, det):
    import stabilizer_stream_b200 as sp
    x = uniform_noise(total, 77 + n) + np.float32(0.1)
    a = sp.AvgOpts(*avg) if avg else None
    for devs in device_lists(world):
        g = sp.Group(n, devices=devs, mode=sp.ShardMode.TIME, hbf=sp.Hbf(hbf))
        if a is not None:
            g.set_avg(a)
        g.set_detrend(sp.Detrend(det))
        g.time_plan(total, k)
        chunks = [g.time_chunk(r) for r in range(world)]
        assert chunks[0].own_lo == 0 and chunks[-1].own_hi is None
        if devs[-1] == 0 and world > 1:
            assert g.info()["reduce"] == "p2p"
            # rank by rank, ragged host calls
            for r, c in enumerate(chunks):
                mid = (c.feed_lo + c.feed_hi) // 2 + 3
                g.time_process(r, x[c.feed_lo:mid])
                g.time_process(r, x[mid:c.feed_hi])
        else:
            g.time_process_all(x)
        g.time_finish()
        check_against_sequential(sp, oracle, g, x, n, hbf, a, det)


@pytest.mark.parametrize("world", [3, 2])
def test_time_chunked_group_on_the_device_generated_stream(oracle, world):
    """config 5 in small: every rank generates its own range of the counter-based stream on its device (with one GPU
    per rank where the box has them: the sources then live on several devices of one process)"""
    import stabilizer_stream_b200 as sp
    import torch
    n = 512
    for devs in device_lists(world):
        g = sp.Group(n, devices=devs, mode=sp.ShardMode.TIME)
        for total in (6_000_000, 4_100_003):     # a second capture reuses the group's handles and sources
            x = oracle.Source(oracle.SOURCE_NOISE, 0, 0x7654321).get(total)
            g.time_plan(total)
            g.time_process_noise(0, 0x7654321)
            g.time_finish()
            assert torch.cuda.current_device() == 0, "the library left the caller's current device changed"
            check_against_sequential(sp, oracle, g, x, n, 1, None, 0)


def test_entry_points_restore_the_current_device():
    """a host application (or torch) keeps its own notion of the current device: handles that live on another GPU
    must not leave it changed (found on an 8-GPU box: Source::generate did)"""
    import stabilizer_stream_b200 as sp
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(0)
    src = sp.Source.noise(0, device=1)
    x = src.get(100_000)
    assert x.device.index == 1 and torch.cuda.current_device() == 0
    c = sp.PsdCascade(512, device=1)
    c.process_source(src, 300_000)
    c.process(x)
    p, b = c.psd()
    assert torch.cuda.current_device() == 0 and np.all(np.isfinite(p))
    dec = sp.FrameDecoder(device=1)
    del src, c, dec
    assert torch.cuda.current_device() == 0


def test_time_chunked_group_errors():
    import stabilizer_stream_b200 as sp
    from stabilizer_stream_b200 import _lib as L
    g = sp.Group(512, devices=[0, 0], mode=sp.ShardMode.TIME)
    with pytest.raises(L.SspsdError):
        g.time_finish()                      # no plan
    g.time_plan(1_000_000, 2)
    with pytest.raises(L.SspsdError):
        g.time_finish()                      # not fed
    with pytest.raises(L.SspsdError):
        g.psd(0)                             # not finished
    c = g.time_chunk(0)
    with pytest.raises(L.SspsdError):
        g.time_process(0, np.zeros(c.feed_hi - c.feed_lo + 1, np.float32))   # more than the rank's range
    with pytest.raises(L.SspsdError):
        sp.Group(512, devices=[0], mode=sp.ShardMode.CHANNELS).time_plan(1000)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_channel_group_equals_independent_cascades(oracle, world):
    import stabilizer_stream_b200 as sp
    n, n_ch = 512, 5
    xs = [uniform_noise(200 * n + 17 * c, 500 + c) * np.float32(1 + c) for c in range(n_ch)]
    for devs in device_lists(world):
        g = sp.Group(n, devices=devs, mode=sp.ShardMode.CHANNELS)
        g.set_detrend(sp.Detrend.MIDPOINT)
        for c in range(n_ch):                        # `for (trace, dec) in traces: dec.process(&trace)` (bin/psd.rs:174-182)
            if c == 3:
                continue                             # a trace that never shows up
            half = xs[c].size // 2
            g.process(c, xs[c][:half])
            g.process(c, xs[c][half:])
            assert g.channel_device(c) == (devs[c % world], c % world)
        allp = g.psd_all(n_ch)
        for c in range(n_ch):
            p, b = g.psd(c)
            assert np.array_equal(p, allp[c][0]) and [breaks_tuple(v) for v in b] == [breaks_tuple(v) for v in allp[c][1]]
            if c == 3:
                assert p.size == 0 and b == []
                continue
            o = oracle.Cascade(n, 1)
            o.set_detrend(1)
            o.process(xs[c])
            po, bo = o.psd()
            assert [breaks_tuple(v) for v in b] == [v.as_tuple() for v in bo]
            floor = 1e-5 * np.median(po)
            assert np.max((np.abs(p - po) - floor) / po) < 1e-4


@pytest.mark.parametrize("mode", ["time", "time_preset", "channels"])
def test_one_process_per_gpu_over_nccl(mode):
    """ncclCommInitRank path: torchrun, 2 ranks, 2 GPUs (tools/group_rank.py); skipped on a single-GPU box"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29581", os.path.join(ROOT, "tools", "group_rank.py"), "--mode", mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["world"] == 2, out

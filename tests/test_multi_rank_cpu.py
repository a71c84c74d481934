"""world_size = 2 gloo test of the multi-GPU host logic (CPU only): channel sharding + readout gather,
and time-chunked partial accumulators + sum reduction.  The per-rank spectra come from the CPU oracle
(no GPU here); what is under test is stabilizer_stream_b200.multi."""
import os
import socket

import numpy as np
import pytest

from conftest import uniform_noise


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from oracle import binding as orc
    from stabilizer_stream_b200 import multi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 512
        # ---- channels: 5 channels over 2 ranks, gather at readout ----
        mine = multi.shard_channels(5, world, rank)
        assert mine == ([0, 2, 4] if rank == 0 else [1, 3])
        spectra = {}
        for c in mine:
            o = orc.Cascade(n, orc.HBF_140)
            o.process(uniform_noise(40 * n + 100 * c, 100 + c))
            spectra[c] = o.psd()[0]
        for slot in range(3):
            p = spectra[mine[slot]] if slot < len(mine) else np.zeros(0, np.float32)
            got = multi.gather_spectra(p, dist, torch.device("cpu"))
            if rank == 0:
                for r, g in enumerate(got):
                    c = slot * world + r
                    if c < 5:
                        o = orc.Cascade(n, orc.HBF_140)
                        o.process(uniform_noise(40 * n + 100 * c, 100 + c))
                        assert np.array_equal(g, o.psd()[0]), "channel %d" % c
                    else:
                        assert g.size == 0
        # ---- time chunks of one stage: disjoint segment ranges, one sum reduction ----
        x = uniform_noise(301 * (n // 2) + 17, 7)
        hop = n // 2
        nseg = 1 + (x.size - n) // hop
        k0, k1 = multi.split_segments(nseg, world)[rank]
        a, b = multi.segment_sample_range(k0, k1, n, hop)
        st = orc.Stage(n)
        st.process(x[a:b])
        assert st.count() == k1 - k0
        acc = torch.zeros(16, 320)
        acc[0, :n // 2 + 1] = torch.from_numpy(st.spectrum())
        acc, counts = multi.reduce_partials(acc, [st.count()], dist)
        full = orc.Stage(n)
        full.process(x)
        assert counts == [full.count()] == [nseg]
        np.testing.assert_allclose(acc[0, :n // 2 + 1].numpy(), full.spectrum(), rtol=2e-5)
        # ---- the single readout collective of the time-chunked mode: rows + counts summed, tail slices and
        # their positions carried in disjoint slots (pack_exchange / unpack_exchange over a gloo reduce) ----
        total, n_local, stride = 3_000_017, 3, 260
        lay = multi.exchange_layout(total, world, n, 1, n_local, stride)
        stream = uniform_noise(2 * lay["slot"] - 100, 5)          # the "stage-3 stream": two consecutive slices
        cut = lay["slot"] - 37
        first, piece = (0, stream[:cut]) if rank == 0 else (cut, stream[cut:])
        rows = torch.from_numpy(uniform_noise(16 * stride, 20 + rank).reshape(16, stride) ** 2)
        factors = [1.0, 0.5, 1.0] if rank == 0 else [1.0, 1.0, 1.0]
        counts = [1000 + rank, 120 + rank, 2 ** 30 + 15 + rank]    # beyond f32 integer range on purpose
        buf = multi.pack_exchange(lay, rank, world, n_local, rows, factors, counts, first,
                                  torch.from_numpy(piece.copy()), piece.size, torch.device("cpu"))
        dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            got_rows, got_counts, tails = multi.unpack_exchange(lay, world, n_local, buf)
            r0 = uniform_noise(16 * stride, 20).reshape(16, stride) ** 2
            r1 = uniform_noise(16 * stride, 21).reshape(16, stride) ** 2
            want = r0[:3].astype(np.float64) * np.array([1.0, 0.5, 1.0])[:, None] + r1[:3]
            np.testing.assert_allclose(got_rows.numpy(), want, rtol=1e-6)
            assert got_counts == [2001, 241, 2 ** 31 + 31]
            assert [t[0] for t in tails] == [0, cut]
            assert np.array_equal(np.concatenate([t[1].numpy() for t in tails]), stream)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_split_segments_partitions_exactly():
    from stabilizer_stream_b200 import multi
    for nseg in (0, 1, 7, 97655, 2343749):
        for world in (1, 2, 3, 8):
            r = multi.split_segments(nseg, world)
            assert r[0][0] == 0 and r[-1][1] == nseg
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1

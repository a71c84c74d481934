"""Multi-GPU host logic in Python (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Since round 2 the product path is the library's own sspsd_group_* API (csrc/sspsd_group.cu; `Group` in psd.py):
planner, seek / window, the single NCCL reduction and the deep stages on rank 0 run inside libsspsd.so, and
bench.py / tools/run_configs.py call that.  This module stays as the executable specification the C++ planner is
tested against (tests/test_group_plan_cpu.py) and as the gloo-testable model of the exchange
(tests/test_multi_rank_cpu.py); `time_chunked_psd` below is the round-1 orchestration of the same steps.

The reference is single threaded and loops over its traces sequentially (src/bin/psd.rs:174-182);
two partitionings fall out of the cascade's structure:

* channels: every trace owns an independent PsdCascade -> channel c lives on rank c % world, nothing
  is exchanged while processing, and the merged spectra are gathered to rank 0 at readout;
* time chunks of one stream: the per-stage |X|^2 accumulators are plain sums over segments (boxcar
  averaging), so ranks that processed disjoint segment ranges combine them with ONE sum-reduction of
  the accumulator array (sspsd_cascade_partials) at readout.

Everything here takes torch tensors, so the same code runs over NCCL (CUDA tensors) and gloo (CPU).
"""
import numpy as np


def shard_channels(n_channels, world, rank):
    """Channels owned by `rank`: c % world == rank (config 4)."""
    return [c for c in range(n_channels) if c % world == rank]


def gather_spectra(p, dist, device, dst=0, max_len=None):
    """Gather variable-length merged spectra (1-D float32 numpy arrays, one per rank) to rank `dst`
    with a single collective.  Returns the list of arrays on dst, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    if max_len is None:
        n = torch.tensor([p.size], dtype=torch.int64, device=device)
        dist.all_reduce(n, op=dist.ReduceOp.MAX)
        max_len = int(n.item())
    buf = torch.zeros(max_len + 1, dtype=torch.float32, device=device)
    buf[0] = float(p.size)
    if p.size:
        buf[1:1 + p.size] = torch.from_numpy(np.ascontiguousarray(p, np.float32)).to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    res = []
    for t in out:
        t = t.cpu().numpy()
        res.append(t[1:1 + int(t[0])].copy())
    return res


def split_segments(n_segments, world):
    """Contiguous, balanced segment ranges [k0, k1) per rank for time-chunking one stage."""
    base, rem = divmod(n_segments, world)
    out, k = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((k, k + n))
        k += n
    return out


def segment_sample_range(k0, k1, n_fft, hop):
    """Samples a rank must see to own segments [k0, k1) of a stage: [k0*hop, (k1-1)*hop + N)."""
    if k1 <= k0:
        return (k0 * hop, k0 * hop)
    return (k0 * hop, (k1 - 1) * hop + n_fft)


def reduce_partials(acc, counts, dist, dst=None):
    """Sum-reduce partial accumulators ([stages, stride] float32 tensor, modified in place) and the
    per-stage segment counts (list of ints) over all ranks: one collective for the spectra, one tiny
    one for the counts.  With dst=None every rank gets the result (all_reduce)."""
    import torch
    c = torch.tensor(list(counts), dtype=torch.int64, device=acc.device)
    if dst is None:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)
        dist.reduce(c, dst=dst, op=dist.ReduceOp.SUM)
    return acc, [int(v) for v in c.tolist()]


class DeviceArray:
    """Zero-copy view of device memory (pointer from the C ABI) for torch.as_tensor(..., device='cuda')."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def cascade_partials_tensor(cascade):
    """The cascade's accumulator array as a CUDA tensor [MAX_STAGES, stride] aliasing library memory,
    plus the per-stage segment counts (sspsd_cascade_partials)."""
    import torch
    pc = cascade.partials()
    cascade.sync()
    t = torch.as_tensor(DeviceArray(pc.acc, (16, int(pc.acc_stride))), device="cuda")
    return t, [int(pc.count_raw[i]) for i in range(pc.n_stages)]


# =============================================================================================
# Time-chunked processing of ONE long stream (BASELINE config 5): planner + orchestration.
# Mirrors the integer bookkeeping of the library (csrc/sspsd_cascade.cu: seek, next_valid_from,
# own_offset) so that every rank's owned segments are provably valid and complete.
# =============================================================================================
class _HbfTable(dict):
    """drain / decimator halo per tap family, asked from the library (sspsd_hbf_info) on first use so the
    planner can never disagree with the kernels' geometry."""

    def __init__(self, which):
        super().__init__()
        self.which = which

    def __missing__(self, hbf):
        import ctypes as C

        from . import _lib as L
        d, h = C.c_uint32(), C.c_uint32()
        L.check(L.lib().sspsd_hbf_info(int(hbf), C.byref(d), C.byref(h)))
        self[hbf] = (d.value, h.value)[self.which]
        return self[hbf]


DRAIN = _HbfTable(0)     # hbf_dec_response_length(3) per tap family
DEC_HALO = _HbfTable(1)  # history the decimator kernel reads before a block (DecGeom::HALO)


def own_offset(i, drain):
    """Stage-0 position of sample 0 of stage i: c_0 = 0, c_{i+1} = 8 (c_i + R)."""
    c = 0
    for _ in range(i):
        c = 8 * (c + drain)
    return c


def stream_state(total, n, hop, drain, max_stages=16):
    """Closed-form state of a sequential cascade after `total` samples:
    list of (samples received L_i, segments craw_i, samples emitted to the next stage)."""
    out, L = [], total
    while L > 0 and len(out) < max_stages:
        craw = 0 if L < n else 1 + (L - n) // hop
        D = n + (craw - 1) * hop if craw else 0
        em = max(D // 8 - drain, 0)
        out.append((L, craw, em))
        L = em
    return out


def needed_input(stage, j_end, n, hop, drain):
    """Stage-0 samples needed so that stage `stage` has received at least j_end samples."""
    for _ in range(stage):
        d = 8 * (j_end + drain)                       # the previous stage must have decimated D >= d
        craw = 1 + max(0, -(-(d - n) // hop))
        j_end = n + (craw - 1) * hop
    return j_end


def valid_from(pos, stages, halo, drain):
    """Per-stage first stream index free of warm-up contamination for a cascade seek()ed to `pos`."""
    v, out = pos, []
    for _ in range(stages + 1):
        out.append(v)
        v = 0 if v == 0 else max((v + halo + 1) // 8 - drain, 0)
    return out


def first_index_at_or_after(pos, i, unit, drain):
    """Smallest k with own(i, k*unit) >= pos (unit = hop for segments, 1 for samples)."""
    c = own_offset(i, drain)
    step = unit * 8 ** i
    return 0 if pos <= c else -(-(pos - c) // step)


def plan_time_chunks(total, world, n, hbf=1, n_local=3, window_overlap=None):
    """Cut a stream of `total` samples into `world` chunks.  Returns one dict per rank:
    own_lo/own_hi (ownership interval in stage-0 samples, own_hi None for the last rank),
    feed_lo/feed_hi (samples the rank must process: chunk + FIR warm-up halo + completion halo),
    tail_lo/tail_hi (owned index range of the stage-n_local input stream)."""
    hop = n - (n // 2 if window_overlap is None else window_overlap)
    drain, halo = DRAIN[hbf], DEC_HALO[hbf]
    bounds = [(g * total // world) // hop * hop for g in range(world)] + [total]
    plans = []
    for g in range(world):
        own_lo, own_hi = bounds[g], bounds[g + 1]
        last = g == world - 1
        # ---- how far forward: every owned segment complete, every owned tail sample produced ----
        feed_hi = total
        if not last:
            need = own_hi
            for i in range(n_local):
                k_hi = first_index_at_or_after(own_hi, i, hop, drain)
                if k_hi > 0:
                    need = max(need, needed_input(i, (k_hi - 1) * hop + n, n, hop, drain))
            j_hi = first_index_at_or_after(own_hi, n_local, 1, drain)
            need = max(need, needed_input(n_local, j_hi, n, hop, drain))
            feed_hi = min(total, need)
        # ---- how far back: first owned segment / tail sample of every stage must be valid ----
        if own_lo == 0:
            feed_lo = 0
        else:
            back = 8 ** n_local * 64
            while True:
                feed_lo = max(0, (own_lo - back) // 8 * 8)
                v = valid_from(feed_lo, n_local, halo, drain)
                ok = all(first_index_at_or_after(own_lo, i, hop, drain) * hop >= v[i] for i in range(n_local))
                ok = ok and first_index_at_or_after(own_lo, n_local, 1, drain) >= v[n_local]
                if ok or feed_lo == 0:
                    break
                back *= 2
        plans.append(dict(rank=g, own_lo=own_lo, own_hi=None if last else own_hi, feed_lo=feed_lo, feed_hi=feed_hi,
                          tail_lo=first_index_at_or_after(own_lo, n_local, 1, drain),
                          tail_hi=None if last else first_index_at_or_after(own_hi, n_local, 1, drain)))
    return plans


def stage_avg(avg, i):
    """PsdCascade::get_or_add / set_avg: avg_i = min(count >> 3 i, limit) (psd.rs:434, 449)"""
    if avg is None:
        return 0xFFFFFFFF
    return min(avg.count >> (3 * i), avg.limit)


def ewma_tail_factors(plan, total, n, hbf, n_local, avg):
    """Per local stage: the factor a rank's accumulator row is multiplied with before the reduction, and
    the averaging count (Psd::count) of the completed stage.  A rank's row is normalised as if the stream
    ended after its last owned segment b; every later segment j rescales the running sum by
    g = avg/(avg+1) iff the reference's count exceeds avg there, i.e. iff j >= avg + 1 (psd.rs:218-225)."""
    hop = n // 2
    drain = DRAIN[hbf]
    st = stream_state(total, n, hop, drain)
    factors, counts = [], []
    for i in range(n_local):
        a = stage_avg(avg, i)
        s_i = st[i][1] if i < len(st) else 0
        if a == 0xFFFFFFFF:
            factors.append(1.0)
            counts.append(s_i)
            continue
        b = s_i if plan["own_hi"] is None else min(s_i, first_index_at_or_after(plan["own_hi"], i, hop, drain))
        later = max(0, s_i - max(b, a + 1))
        g = float(np.float32(a) / np.float32(a + 1))   # the reference divides in f32 (psd.rs:219)
        factors.append(g ** later)
        counts.append(min(s_i, a + 1))
    return factors, counts


def run_chunk(cascade, plan, samples, n_local):
    """One rank's share: position the fresh cascade, restrict it, process samples[feed_lo:feed_hi]
    (`samples` is that slice, host or device), and return its slice of the stage-n_local stream."""
    cascade.seek(plan["feed_lo"])
    cascade.set_window(plan["own_lo"], plan["own_hi"], n_local)
    cascade.process(samples)
    first, tail = cascade.take_tail(plan["tail_lo"], plan["tail_hi"] if plan["tail_hi"] is not None else 2 ** 63)
    return first, tail


def finish_on_root(root, reduced_counts, tails, total, n, hbf, n_local, avg=None):
    """Rank 0 after the reduction of the accumulator rows: install the global bookkeeping of the local
    stages and run the deep stages on the gathered stage-n_local stream."""
    hop = n // 2
    st = stream_state(total, n, hop, DRAIN[hbf])
    for i in range(min(n_local, len(st))):
        assert reduced_counts[i] == st[i][1], "stage %d: %d segments reduced, %d expected" % (i, reduced_counts[i], st[i][1])
        a = stage_avg(avg, i)
        root.set_stream_state(i, st[i][0], st[i][1] if a == 0xFFFFFFFF else min(st[i][1], a + 1))
    pos = 0
    parts = []
    for first, tail in tails:
        size = int(tail.numel()) if hasattr(tail, "numel") else int(tail.size)
        assert first == pos or size == 0, "tail slices are not contiguous: %d != %d" % (first, pos)
        if size:
            parts.append(tail)
            pos = first + size
    if parts:
        # the slices are consecutive pieces of one stream: feed them in one call (one set of launches for
        # the deep stages instead of one per rank)
        if len(parts) > 1 and all(hasattr(t, "numel") for t in parts):
            import torch
            parts = [torch.cat(parts)]
            torch.cuda.synchronize()  # made on torch's stream, read on the library's deep stream
        for t in parts:
            root.process_stage(n_local, t)
    if len(st) > n_local:
        assert pos == st[n_local][0], "tail stream length %d != %d" % (pos, st[n_local][0])
    return root


def exchange_layout(total, world, n, hbf, n_local, stride):
    """Layout of the single reduction buffer of the time-chunked readout (float64 words):
    [n_local accumulator rows of `stride`] [world tail slots of `slot`] [n_local counts] [world x (first, length)]"""
    slot = (stream_state(total, n, n // 2, DRAIN[hbf]) + [(0, 0, 0)] * 16)[n_local][0] // world + 64
    nacc = n_local * stride
    meta = nacc + world * slot
    return dict(slot=slot, nacc=nacc, meta=meta, size=meta + n_local + 2 * world, stride=stride)


def pack_exchange(lay, rank, world, n_local, acc_rows, factors, counts, first, tail32, tlen, device):
    """This rank's contribution: its rows (times the EWMA factors) and counts are summed by the reduction,
    its tail slice and (first, length) sit in slots nobody else writes."""
    import torch
    buf = torch.zeros(lay["size"], dtype=torch.float64, device=device)
    stride = lay["stride"]
    buf[:lay["nacc"]] = acc_rows[:n_local].reshape(-1)
    for i, f in enumerate(factors):  # EWMA: weight of everything that follows this rank's chunk
        if f != 1.0:
            buf[i * stride:(i + 1) * stride] *= f
    base = lay["nacc"] + rank * lay["slot"]
    buf[base:base + tlen] = tail32[:tlen]
    book = [0.0] * (n_local + 2 * world)
    book[:n_local] = [float(c) for c in counts]
    book[n_local + 2 * rank] = float(first)
    book[n_local + 2 * rank + 1] = float(tlen)
    buf[lay["meta"]:] = torch.tensor(book, dtype=torch.float64).to(device, non_blocking=True)
    return buf


def unpack_exchange(lay, world, n_local, buf):
    """Root side, after the sum reduction: (rows [n_local, stride] float32, counts, [(first, tail float32)])"""
    book = [int(round(v)) for v in buf[lay["meta"]:].tolist()]  # the one device->host readback (synchronises)
    rows = buf[:lay["nacc"]].reshape(n_local, lay["stride"]).float()
    all32 = buf[lay["nacc"]:lay["meta"]].float()
    slot = lay["slot"]
    tails = [(book[n_local + 2 * r], all32[r * slot:r * slot + book[n_local + 2 * r + 1]]) for r in range(world)]
    return rows, book[:n_local], tails


def time_chunked_psd(cascade, feed, total, n, dist=None, hbf=1, n_local=3, device="cuda", timings=None, avg=None):
    """Distributed driver of the time-chunked mode: every rank calls this with a FRESH cascade and a
    callable feed(lo, hi, sink) that pushes stream samples [lo, hi) into sink(x) in order (any block
    size).  One NCCL sum-reduction of the local stages' accumulator rows + counts, one gather of the
    (tiny) stage-n_local stream slices; rank 0 returns its completed cascade (call .psd() on it), the
    other ranks return None.  `timings` (dict) receives wall-clock seconds per phase."""
    import time

    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    t = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            torch.cuda.synchronize()
            now = time.perf_counter()
            timings[name] = now - t[0]
            t[0] = now

    plan = plan_time_chunks(total, world, n, hbf, n_local)[rank]
    if avg is not None:
        cascade.set_avg(avg)
    cascade.seek(plan["feed_lo"])
    cascade.set_window(plan["own_lo"], plan["own_hi"], n_local)
    feed(plan["feed_lo"], plan["feed_hi"], cascade.process)
    factors, _ = ewma_tail_factors(plan, total, n, hbf, n_local, avg)
    tail_hi = plan["tail_hi"] if plan["tail_hi"] is not None else 2 ** 63
    if world == 1:
        first, tail = cascade.take_tail(plan["tail_lo"], tail_hi)
        lap("process+tail")
        acc, counts = cascade_partials_tensor(cascade)
        counts = (counts + [0] * 16)[:n_local]
        for i, f in enumerate(factors):
            if f != 1.0:
                acc[i] *= f
        torch.cuda.synchronize()
        tails = [(first, tail)]
    else:
        # ONE collective for the readout: accumulator rows of the local stages, their counts and the
        # rank's slice of the stage-n_local stream travel in one buffer; rows and counts are summed, the
        # slices land in disjoint slots (everybody else contributes zeros there).  Everything stays on
        # the device; the root reads back only the 2 * world + n_local bookkeeping numbers.
        acc, counts = cascade_partials_tensor(cascade)
        lay = exchange_layout(total, world, n, hbf, n_local, acc.shape[1])
        tail32 = torch.empty(lay["slot"], dtype=torch.float32, device=device)
        first, tlen = cascade.take_tail_device(plan["tail_lo"], tail_hi, tail32)
        lap("process+tail")
        counts = (counts + [0] * 16)[:n_local]
        buf = pack_exchange(lay, rank, world, n_local, acc, factors, counts, first, tail32, tlen, device)
        dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
        if rank != 0:
            torch.cuda.synchronize()
            lap("reduce")
            return None
        rows, counts, tails = unpack_exchange(lay, world, n_local, buf)
        lap("reduce")
        acc[:n_local] = rows
        torch.cuda.synchronize()  # torch's stream wrote library memory / made the tensors the library reads
    root = finish_on_root(cascade, counts, tails, total, n, hbf, n_local, avg)
    if world > 1:
        root.sync()  # the gathered slices are torch temporaries read asynchronously on the library's deep stream
    lap("finish")
    return root

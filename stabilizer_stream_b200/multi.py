"""Multi-GPU host logic (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

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

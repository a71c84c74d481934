"""Time-chunked processing of one stream (BASELINE config 5) on one GPU: `world` handles play the ranks one
after the other, their accumulator rows are summed like the NCCL reduction would, rank 0 finishes the deep
stages on the gathered stage-K stream.  The result must equal the sequential cascade."""
import numpy as np
import pytest

from conftest import uniform_noise

pytestmark = pytest.mark.gpu


def run_emulated(x, n, world, n_local, hbf=1, avg=None):
    import torch
    from stabilizer_stream_b200 import Hbf, MergeOpts, PsdCascade, multi
    plans = multi.plan_time_chunks(x.size, world, n, hbf, n_local)
    xd = torch.from_numpy(x).cuda()
    cas, tails, accs, counts = [], [], [], []
    for p in plans:
        c = PsdCascade(n, hbf=Hbf(hbf))
        if avg is not None:
            c.set_avg(avg)
        tails.append(multi.run_chunk(c, p, xd[p["feed_lo"]:p["feed_hi"]], n_local))
        t, cnt = multi.cascade_partials_tensor(c)
        factors, _ = multi.ewma_tail_factors(p, x.size, n, hbf, n_local, avg)
        for i, f in enumerate(factors):   # what time_chunked_psd does before its reduction
            t[i] *= f
        cas.append(c)
        accs.append(t)
        counts.append((cnt + [0] * 16)[:16])
    total_acc = accs[0][:n_local].clone()
    for t in accs[1:]:
        total_acc += t[:n_local]
    accs[0][:n_local].copy_(total_acc)
    torch.cuda.synchronize()   # torch's stream wrote into library memory; the library uses its own stream
    red = [sum(c[i] for c in counts) for i in range(n_local)]
    root = multi.finish_on_root(cas[0], red, tails, x.size, n, hbf, n_local, avg)
    return root, plans


def breaks_tuple(b):
    return (b.start, b.include, b.count, b.bins.start, b.bins.stop, b.fft_size, b.decimation, b.pending, b.processed)


@pytest.mark.parametrize("n,total,world,k,hbf", [(512, 3_000_017, 4, 3, 1), (4096, 30_000_000, 3, 2, 1), (512, 2_500_000, 2, 2, 0),
                                                  (64, 700_001, 5, 3, 1)])
def test_time_chunked_equals_sequential(oracle, n, total, world, k, hbf):
    import torch
    from stabilizer_stream_b200 import Hbf, MergeOpts, PsdCascade
    x = uniform_noise(total, 77 + n) + np.float32(0.1)
    root, plans = run_emulated(x, n, world, k, hbf)
    p, b = root.psd(MergeOpts())
    seq = PsdCascade(n, hbf=Hbf(hbf))
    seq.process(torch.from_numpy(x).cuda())
    ps, bs = seq.psd(MergeOpts())
    assert [breaks_tuple(v) for v in b] == [breaks_tuple(v) for v in bs]
    assert p.size == ps.size
    np.testing.assert_allclose(p, ps, rtol=2e-5, atol=1e-6 * float(np.median(ps)))
    # and against the CPU oracle, all stages, all bins
    o = oracle.Cascade(n, hbf)
    o.process(x)
    pk, bk = root.psd(MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
    pok, bok = o.psd(True, 0, True)
    assert [v.count for v in bk] == [v.count for v in bok]
    for bi in bk:
        if bi.count:   # a trailing stage with count 0 is normal (gain 0 -> NaN bins in the reference too)
            sl = slice(bi.start, bi.start + len(bi.bins))
            floor = 1e-5 * np.median(pok[sl])
            assert np.max((np.abs(pk[sl] - pok[sl]) - floor) / pok[sl]) < 1e-4, "stage dec=%d" % bi.decimation
    # the halo overhead of the plan is what the planner promised
    assert all(q["feed_lo"] <= q["own_lo"] for q in plans)


@pytest.mark.parametrize("n,total,world,k,limit,count", [(512, 3_000_017, 4, 3, 999, 2 ** 32 - 2), (512, 2_000_000, 3, 2, 17, 2 ** 32 - 1),
                                                          (4096, 30_000_000, 4, 2, 200, 5000), (256, 1_500_000, 5, 3, 0, 64),
                                                          (512, 1_200_000, 8, 2, 3, 2 ** 32 - 1)])
def test_time_chunked_ewma_equals_sequential(oracle, n, total, world, k, limit, count):
    """the `psd` binary's preset (avg limit 999, bin/psd.rs:73-78) and other finite averages: the EWMA is a
    recurrence over the GLOBAL segment order, so every rank weights its rows for what follows its chunk"""
    import torch
    from stabilizer_stream_b200 import AvgOpts, MergeOpts, PsdCascade
    avg = AvgOpts(limit=limit, count=count)
    x = uniform_noise(total, 91 + n) + np.float32(0.1)
    root, plans = run_emulated(x, n, world, k, 1, avg)
    seq = PsdCascade(n)
    seq.set_avg(avg)
    seq.process(torch.from_numpy(x).cuda())
    o = MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True)
    p, b = root.psd(o)
    ps, bs = seq.psd(o)
    assert [breaks_tuple(v) + (v.avg,) for v in b] == [breaks_tuple(v) + (v.avg,) for v in bs]
    for bi in b:
        if bi.count:
            sl = slice(bi.start, bi.start + len(bi.bins))
            np.testing.assert_allclose(p[sl], ps[sl], rtol=5e-5, atol=1e-5 * float(np.median(ps[sl])))
    # and against the CPU restatement
    oc = oracle.Cascade(n, 1)
    oc.set_avg(limit, count)
    oc.process(x)
    po, bo = oc.psd(True, 0, True)
    assert [v.count for v in b] == [v.count for v in bo]
    for bi in b:
        if bi.count:
            sl = slice(bi.start, bi.start + len(bi.bins))
            floor = 1e-5 * np.median(po[sl])
            assert np.max((np.abs(p[sl] - po[sl]) - floor) / po[sl]) < 1e-4, "stage dec=%d" % bi.decimation


def _fuzz_cases():
    import os
    return list(range(4)) + [2000 + i for i in range(int(os.environ.get("SSPSD_FUZZ_EXTRA", "0")))]


@pytest.mark.parametrize("seed", _fuzz_cases())
def test_random_time_chunk_plans(seed):
    """random stream length / rank count / FFT size / tap family / number of locally processed stages:
    the chunked run must reproduce the sequential run's bookkeeping exactly and its bins to rounding"""
    import torch
    from stabilizer_stream_b200 import Hbf, MergeOpts, PsdCascade
    rng = np.random.default_rng(seed)
    n = int(rng.choice([64, 128, 256, 512, 1024, 2048, 4096]))
    world = int(rng.integers(1, 9))
    hbf = int(rng.integers(0, 2))
    k = int(rng.integers(1, 4))
    # long enough that every rank owns at least a few segments of the deepest local stage
    total = int(rng.integers(world * n * 8 ** k * 2, world * n * 8 ** k * 6)) + int(rng.integers(0, 8 ** k))
    total = min(total, 60_000_000)
    if rng.random() < 0.3:  # streams that are short for this many ranks: some ranks own nothing in the deep stages
        total = int(rng.integers(1, world * n * 8 ** k))
    x = uniform_noise(total, 300 + seed) + np.float32(0.1)
    from stabilizer_stream_b200 import AvgOpts
    avg = None
    if rng.random() < 0.5:
        avg = AvgOpts(limit=int(rng.choice([0, 1, 5, 60, 999])), count=int(rng.choice([2 ** 32 - 1, 2 ** 32 - 2, 700, 9])))
    root, plans = run_emulated(x, n, world, k, hbf, avg)
    p, b = root.psd(MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
    seq = PsdCascade(n, hbf=Hbf(hbf))
    if avg is not None:
        seq.set_avg(avg)
    seq.process(torch.from_numpy(x).cuda())
    ps, bs = seq.psd(MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
    assert [breaks_tuple(v) for v in b] == [breaks_tuple(v) for v in bs]
    for bi in b:
        if bi.count:
            sl = slice(bi.start, bi.start + len(bi.bins))
            np.testing.assert_allclose(p[sl], ps[sl], rtol=5e-5, atol=1e-5 * float(np.median(ps[sl])))

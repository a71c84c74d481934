"""Parity of the CUDA cascade (through the C ABI) against the CPU oracle.  Needs a B200."""
import numpy as np
import pytest

from conftest import uniform_noise

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # north_star: "within a stated relative tolerance per PSD bin (1e-4 in f32)"
AFLOOR = 1e-5  # absolute floor, as a fraction of the stage's median bin power: only matters for
#                bins that are accidentally ~0 in a stage with a single segment


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available()
    import stabilizer_stream_b200 as m
    return m


DC_RTOL = 1e-2  # bins 0 and 1 under Mean/Span detrend when no f64 truth is at hand (mid-stream option changes):
#                 what is left after subtracting the offset is set by the rounding of that offset; the oracle
#                 sums sequentially in f32 like the reference (src/psd.rs:100,104), the GPU sums as a tree, so
#                 the two agree only to ~N*eps*DC/sigma.  Wherever the float64 model can follow the run the
#                 statement is instead: the device is at least as close to the TRUTH on those bins as the
#                 reference's own f32 arithmetic, |gpu - f64| <= |oracle - f64| + 1e-4 * median(row).


def assert_bins_close(got, want, what="", loose_head=0, truth=None):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, what
    if want.size == 0:
        return
    floor = AFLOOR * np.median(np.abs(want))
    err = np.abs(got - want) - floor
    rel = err / np.maximum(np.abs(want), 1e-300)
    if loose_head:
        if truth is not None:
            truth = np.asarray(truth, np.float64)
            assert truth.shape == want.shape, what
            dg = np.abs(got[:loose_head] - truth[:loose_head])
            do = np.abs(want[:loose_head] - truth[:loose_head])
            eps = RTOL * np.median(np.abs(truth))
            assert np.all(dg <= do + eps), "%s: DC bins: |gpu - f64| = %s > |oracle - f64| = %s + %.3g" % (
                what, dg, do, eps)
        else:
            assert np.max(rel[:loose_head]) <= DC_RTOL, "%s: DC bins rel err %.3g" % (what, np.max(rel[:loose_head]))
        rel = rel[loose_head:]
    assert rel.size == 0 or np.max(rel) <= RTOL, "%s: max rel err %.3g" % (what, np.max(rel))


def merged_truth(c64, breaks, n):
    """PsdCascade::psd (src/psd.rs:479-543) evaluated on the float64 model's rows with the given breaks."""
    out = []
    for b in breaks:
        if not b.include:
            continue
        stage = 0
        while 8 ** stage < b.decimation:
            stage += 1
        row, cnt, _, _ = c64.stage(stage)
        gain = (n // 2) * cnt * 1.5 * 0.25
        out.append(row[b.bins.start:b.bins.stop] / (gain * b.decimation))
    return np.concatenate(out) if out else np.zeros(0)


def breaks_tuple(b):
    return (b.start, b.include, b.count, b.avg, b.bins.start, b.bins.stop, b.fft_size, b.decimation, b.pending,
            b.processed)


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("window", [0, 1])
def test_single_stage_all_sizes(sp, oracle, n, window):
    x = uniform_noise(40 * n + 13, n + window) + np.float32(0.1)
    g = sp.Psd(n, sp.Window(window))
    o = oracle.Stage(n, window)
    yg = g.process(x)
    yo = o.process(x)
    assert g.count() == o.count() and g.gain() == o.gain()
    assert_bins_close(g.spectrum(), o.spectrum(), "stage n=%d" % n)
    assert yg.shape == yo.shape
    np.testing.assert_allclose(yg, yo, atol=3e-6)
    np.testing.assert_array_equal(g.buf(), o.buf())


@pytest.mark.parametrize("n", [512, 4096])
@pytest.mark.parametrize("det", [0, 1, 2, 3])
def test_detrend_modes(sp, oracle, n, det):
    x = uniform_noise(64 * n, 17 * det + n) + np.float32(0.5)
    g = sp.Psd(n)
    g.set_detrend(sp.Detrend(det))
    o = oracle.Stage(n)
    o.set_detrend(det)
    g.process(x)
    o.process(x)
    t = oracle.CascadeF64(n, detrend=det, max_stages=1)
    t.process(x)
    assert_bins_close(g.spectrum(), o.spectrum(), "detrend %d" % det, loose_head=2 if det >= 2 else 0,
                      truth=t.stage(0)[0])
    # and the row against the truth itself (bins 0-1 under Mean/Span are cancellation residue for any f32
    # evaluation: covered by the statement above)
    head = 2 if det >= 2 else 0
    assert_bins_close(g.spectrum()[head:], t.stage(0)[0][head:], "detrend %d vs f64" % det)


@pytest.mark.parametrize("n", [2048, 8192])
def test_span_detrend_device_is_the_accurate_side(sp, oracle, n):
    """DESIGN.md section 4 claims that under Detrend::Span the reference's sequential `offset += slope`
    (src/psd.rs:98-101) is the noisier side of a GPU-vs-restatement difference.  Shown here against the float64
    model on single segments (the worst case: no averaging): the device stays within 1e-4 of the truth on every
    bin from 2 on, and is nowhere further from it than the f32 restatement is (plus 1e-4 of the median bin)."""
    worst_o = 0.0
    for seed in range(4):
        x = uniform_noise(n, 900 + seed) + np.float32(3.0) + np.linspace(0, 7, n, dtype=np.float32)
        g = sp.Psd(n)
        g.set_detrend(sp.Detrend.SPAN)
        o = oracle.Stage(n)
        o.set_detrend(2)
        t = oracle.CascadeF64(n, detrend=2, max_stages=1)
        g.process(x); o.process(x); t.process(x)
        truth = t.stage(0)[0]
        med = np.median(truth)
        dg = np.abs(g.spectrum().astype(np.float64) - truth)
        do = np.abs(o.spectrum().astype(np.float64) - truth)
        assert np.all(dg[2:] <= RTOL * np.maximum(truth[2:], AFLOOR * med) + AFLOOR * med), np.max(dg[2:] / truth[2:])
        assert np.all(dg <= do + RTOL * med)
        worst_o = max(worst_o, float(np.max(do[2:] / np.maximum(truth[2:], AFLOOR * med))))
    print("f32 restatement vs f64 under Span, N=%d: %.3g" % (n, worst_o))


def test_detrend_linear_is_unimplemented(sp):
    # reference: Detrend::Linear => unimplemented!() (src/psd.rs:110)
    from stabilizer_stream_b200 import _lib
    g = sp.PsdCascade(512)
    with pytest.raises(_lib.SspsdError) as e:
        g.set_detrend(sp.Detrend.LINEAR)
    assert e.value.status == _lib.EUNIMPLEMENTED


@pytest.mark.parametrize("n", [512, 4096])
def test_known_answer_hann_ones(sp, n):
    # reference src/psd.rs:562-597 generalised: x = ones -> [4N/3, N/3, 0, ...]
    g = sp.Psd(n)
    g.process(np.ones(n, np.float32))
    p = g.spectrum() / g.gain()
    assert g.count() == 1
    assert abs(p[0] - 4 * n / 3) < 1e-5 * n and abs(p[1] - n / 3) < 1e-5 * n
    assert np.all(np.abs(p[2:]) < 1e-6 * n)


def test_reference_statistical_test(sp):
    """The reference's own live test (src/psd.rs:599-644), seeded, on the device path."""
    x = uniform_noise(1 << 16, 0x7654321)
    n = 512
    s = sp.Psd(n)
    y = s.process(x)
    assert y.size == (x.size >> 3) - 115
    p = s.spectrum() / s.gain()
    assert s.count() == 255
    assert np.all(np.abs(p * 0.5 - 1.0) < 10.0 / np.sqrt(s.count()))
    d = sp.PsdCascade(n)
    d.process(x)
    p, b = d.psd()
    for bi in b:
        seg = p[bi.start:bi.start + len(bi.bins)]
        if bi.include and bi.count:
            assert np.all(np.abs(seg * 0.5 - 1.0) < 10.0 / np.sqrt(bi.count))
            # continuity across stage breaks: the decimator's gain of 8 cancels 1/decimation (psd.rs:516)
            assert abs(np.mean(seg) * 0.5 - 1.0) < 5.0 / np.sqrt(bi.count * seg.size)
    f = sp.Break.frequencies(b)
    assert f.size == p.size and f[0] == 0.0 and f[-1] == 0.5


@pytest.mark.parametrize("n,det,hbf", [(512, 0, 1), (512, 3, 0), (4096, 1, 1), (4096, 0, 0), (64, 2, 1), (1024, 3, 1)])
def test_cascade_matches_oracle_ragged_host_and_device(sp, oracle, n, det, hbf):
    import torch
    x = uniform_noise(700 * n + 77, 3 + n) + np.float32(0.25)
    g = sp.PsdCascade(n, hbf=sp.Hbf(hbf), host_stage=1 << 14)
    g.set_detrend(sp.Detrend(det))
    o = oracle.Cascade(n, hbf)
    o.set_detrend(det)
    o.process(x)
    rng = np.random.default_rng(1)
    xd = torch.from_numpy(x).cuda()
    pos = 0
    while pos < x.size:
        k = int(rng.integers(0, 40 * n))
        if rng.random() < 0.5:
            g.process(x[pos:pos + k])          # host pointer (small calls are staged)
        else:
            g.process(xd[pos:pos + k])         # device pointer, arbitrary alignment
        pos += k
    p, b = g.psd()
    po, bo = o.psd()
    assert [breaks_tuple(k) for k in b] == [k.as_tuple() for k in bo]
    loose = 2 if det >= 2 else 0
    t = oracle.CascadeF64(n, hbf=hbf, detrend=det)
    t.process(x)
    assert_bins_close(p, po, "cascade n=%d" % n, loose_head=loose, truth=merged_truth(t, b, n))
    np.testing.assert_array_equal(sp.Break.frequencies(b), oracle.break_frequencies(bo))
    # every stage on its own, all bins, all merge options
    pk, bk = g.psd(sp.MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
    pok, bok = o.psd(True, 0, True)
    assert [breaks_tuple(k) for k in bk] == [k.as_tuple() for k in bok]
    for bi in bk:
        if bi.count:
            sl = slice(bi.start, bi.start + len(bi.bins))
            assert_bins_close(pk[sl], pok[sl], "stage dec=%d" % bi.decimation, loose_head=loose,
                              truth=merged_truth(t, [bi], n))


def test_cascade_ewma_and_option_changes(sp, oracle):
    n = 512
    x = uniform_noise(3000 * n, 9)
    g = sp.PsdCascade(n)
    o = oracle.Cascade(n, 1)
    # the psd binary's defaults: limit = avg_max - 1 = 999, count = avg - 1 (src/bin/psd.rs:73-78)
    g.set_avg(sp.AvgOpts(limit=999, count=2 ** 32 - 2))
    o.set_avg(999, 2 ** 32 - 2)
    third = x.size // 3
    g.process(x[:third]); o.process(x[:third])
    g.set_detrend(sp.Detrend.MEAN); o.set_detrend(3)
    g.process(x[third:2 * third]); o.process(x[third:2 * third])
    g.set_avg(sp.AvgOpts(limit=50, count=400)); o.set_avg(50, 400)   # lowering avg mid-stream
    g.process(x[2 * third:]); o.process(x[2 * third:])
    p, b = g.psd()
    po, bo = o.psd()
    assert [breaks_tuple(k) for k in b] == [k.as_tuple() for k in bo]
    assert_bins_close(p, po, "ewma", loose_head=2)


def test_clone_reset_and_empty(sp, oracle):
    n = 512
    x = uniform_noise(100 * n, 21)
    g = sp.PsdCascade(n)
    assert g.psd()[0].size == 0 and g.num_stages() == 0
    g.process(np.zeros(0, np.float32))
    assert g.num_stages() == 0
    g.process(x[:100])
    p, b = g.psd()
    assert p.size == 0 and len(b) == 1 and b[0].pending == 100 and not b[0].include
    g.process(x[100:50 * n])
    c = g.clone()
    g.process(x[50 * n:])
    c.process(x[50 * n:])
    p1, b1 = g.psd()
    p2, b2 = c.psd()
    assert [breaks_tuple(k) for k in b1] == [breaks_tuple(k) for k in b2]
    np.testing.assert_allclose(p1, p2, rtol=1e-5)
    o = oracle.Cascade(n, 1)
    o.process(x)
    assert_bins_close(p2, o.psd()[0], "clone")
    g.reset()
    assert g.num_stages() == 0 and g.psd()[0].size == 0
    g.process(x)
    assert_bins_close(g.psd()[0], o.psd()[0], "after reset")


def test_large_batch_properties(sp):
    """BASELINE config 2 size (200e6 samples, N=4096): closed-form stage counts and flat PSD."""
    import torch
    n = 4096
    total = 200_000_000
    g = sp.PsdCascade(n)
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.rand(total, device="cuda", generator=gen) - 0.5) * (12 ** 0.5)
    g.process(x)
    p, b = g.psd()
    counts = [k.count for k in reversed(b)]
    assert counts == [97655, 12205, 1524, 189, 22, 1, 0]   # SURVEY.md 8(a9), drain-independent
    for k in b:
        if k.count >= 20:
            seg = p[k.start:k.start + len(k.bins)]
            assert np.all(np.abs(seg * 0.5 - 1.0) < 10.0 / np.sqrt(k.count))
    # linearity: scaling the input by 2 scales every bin by 4 (same segmentation)
    g2 = sp.PsdCascade(n)
    g2.process(x * 2.0)
    p2, _ = g2.psd()
    np.testing.assert_allclose(p2, 4.0 * p, rtol=1e-5)


def test_ewma_n4096_ring_kernel(sp, oracle):
    """EWMA weights inside the persistent N=4096 kernel, including the boxcar -> EWMA transition."""
    n = 4096
    x = uniform_noise(600 * n, 12)
    g = sp.PsdCascade(n)
    o = oracle.Cascade(n, 1)
    g.set_avg(sp.AvgOpts(limit=99, count=2 ** 32 - 2))
    o.set_avg(99, 2 ** 32 - 2)
    for part in np.array_split(x, 5):
        g.process(part)
        o.process(part)
    p, b = g.psd()
    po, bo = o.psd()
    assert [breaks_tuple(k) for k in b] == [k.as_tuple() for k in bo]
    assert b[-1].count == 100 and b[-1].avg == 99
    assert_bins_close(p, po, "ewma 4096")


def test_small_max_batch_chunks_long_inputs(sp, oracle):
    """process() cuts inputs longer than cfg.max_batch into chunks; results must not change."""
    import torch
    n = 512
    x = uniform_noise(1_000_003, 13)
    g = sp.PsdCascade(n, max_batch=70_000)
    g.process(torch.from_numpy(x).cuda())
    g2 = sp.PsdCascade(n, max_batch=50_000, host_stage=1 << 12)
    g2.process(x)
    o = oracle.Cascade(n, 1)
    o.process(x)
    po, bo = o.psd()
    for c in (g, g2):
        p, b = c.psd()
        assert [breaks_tuple(k) for k in b] == [k.as_tuple() for k in bo]
        assert_bins_close(p, po, "chunked")


@pytest.mark.parametrize("n", [512, 2048])
def test_first_segment_inside_a_full_chunk(sp, oracle, n):
    """short first call, then a full 8N chunk: the corner in which the reference's N-sized ping-pong
    arrays overflow (psd.rs:253 panics); the library and the restatement just continue the stream"""
    x = uniform_noise(n - 1 + 8 * n + 5 * n, 22)
    g = sp.PsdCascade(n, host_stage=1 << 10)
    o = oracle.Cascade(n, 1)
    for c in (g, o):
        c.process(x[:n - 1])
        c.process(x[n - 1:])
    p, b = g.psd(sp.MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
    po, bo = o.psd(True, 0, True)
    assert [breaks_tuple(k) for k in b] == [k.as_tuple() for k in bo]
    for k in b:
        if k.count:
            sl = slice(k.start, k.start + len(k.bins))
            assert_bins_close(p[sl], po[sl], "first segment in chunk, dec %d" % k.decimation)


def test_rect_window_stage_with_decimation(sp, oracle):
    """Window::rectangular (overlap 0, hop = N) through the single-stage API incl. its decimated output."""
    for n in (512, 4096):
        x = uniform_noise(50 * n + 5, 14)
        g = sp.Psd(n, sp.Window.RECTANGULAR)
        o = oracle.Stage(n, oracle.WINDOW_RECT)
        yg = np.concatenate([g.process(x[:7 * n + 3]), g.process(x[7 * n + 3:])])
        yo = np.concatenate([o.process(x[:7 * n + 3]), o.process(x[7 * n + 3:])])
        assert g.count() == o.count() == 50
        np.testing.assert_allclose(yg, yo, atol=2e-5)
        assert_bins_close(g.spectrum(), o.spectrum(), "rect %d" % n)
        np.testing.assert_array_equal(g.buf(), o.buf())


def test_device_tensors_on_other_torch_streams(sp, oracle):
    """a handle created BEFORE torch touched CUDA used to get a private stream, and
    process() read caller tensors on it with no ordering against the torch stream that produced them; a handle
    is also legitimately used under `with torch.cuda.stream(s)` later.  The mirror now makes the handle's stream
    wait for torch's current stream before the call and torch's current stream wait for the handle's after it, so
    temporaries may be dropped right after the call.  Exercised with temporaries produced on side streams and an allocator under churn."""
    import torch
    n = 512
    x = uniform_noise(400 * n, 31)
    g = sp.PsdCascade(n)
    o = oracle.Cascade(n, 1)
    o.process(x)
    xd = torch.from_numpy(x).cuda()
    side = [torch.cuda.Stream(), torch.cuda.Stream()]
    block = 37 * n + 5
    pos = 0
    i = 0
    while pos < x.size:
        s = side[i % 2]
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            # a temporary made on the side stream (non-contiguous view -> .contiguous() copy inside process)
            tmp = torch.stack([xd[pos:pos + block], xd[pos:pos + block]], dim=1) * 1.0
            g.process(tmp[:, 0])
            del tmp
            # allocator churn on the same stream: would recycle the block if the side stream did not wait for the handle's
            junk = torch.full((block * 2,), float("nan"), device="cuda")
            del junk
        pos += block
        i += 1
    p, b = g.psd()
    po, bo = o.psd()
    assert [breaks_tuple(k) for k in b] == [k.as_tuple() for k in bo]
    assert_bins_close(p, po, "side streams")


@pytest.mark.parametrize("n", [512, 4096])
def test_deterministic_accumulation_is_bit_reproducible(sp, oracle, n):
    """SSPSD_FLAG_DETERMINISTIC (VERDICT r01 missing #8): per-CTA partial rows summed in fixed
    order instead of float atomics -- two runs over the same calls give bit-identical readouts."""
    import torch
    x = uniform_noise(3000 * n + 11, 5 + n)
    xd = torch.from_numpy(x).cuda()
    outs = []
    for _ in range(3):
        c = sp.PsdCascade(n, deterministic=True)
        c.set_detrend(sp.Detrend.MEAN)
        c.process(xd[:1000 * n])
        c.process(xd[1000 * n:])
        p, b = c.psd(sp.MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
        outs.append((p, [breaks_tuple(k) for k in b]))
    for p, b in outs[1:]:
        assert b == outs[0][1]
        assert np.array_equal(p.view(np.uint32), outs[0][0].view(np.uint32))
    o = oracle.Cascade(n, 1)
    o.set_detrend(3)
    o.process(x)
    po, bo = o.psd(True, 0, True)
    for bi in [k for k in c.psd(sp.MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))[1] if k.count]:
        sl = slice(bi.start + 2, bi.start + len(bi.bins))
        assert_bins_close(outs[0][0][sl], po[sl], "deterministic stage dec=%d" % bi.decimation)


@pytest.mark.parametrize("n,defer", [(512, 1 << 16), (4096, 1 << 20), (64, 1 << 12)])
def test_deferred_deep_stages_are_invisible(sp, oracle, n, defer):
    """sspsd_config::deep_defer lets the input of the stages >= 1 pile up; every observing call runs what is
    pending first, so results, breaks, option changes, clone and reset behave exactly as without it."""
    import torch
    rng = np.random.default_rng(n)
    x = uniform_noise(1500 * n, 3 * n) + np.float32(0.2)
    xd = torch.from_numpy(x).cuda()
    g = sp.PsdCascade(n, deep_defer=defer, max_batch=64 * n)
    o = oracle.Cascade(n, 1)
    pos = 0
    step = 0
    clone = None
    while pos < x.size:
        k = int(rng.integers(1, 90 * n))
        (g.process(xd[pos:pos + k]) if step % 2 else g.process(x[pos:pos + k]))
        o.process(x[pos:pos + k])
        pos += k
        step += 1
        if step == 5:
            g.set_detrend(sp.Detrend.MIDPOINT); o.set_detrend(1)      # applies to segments completed after the call
        if step == 9:
            g.set_avg(sp.AvgOpts(limit=40, count=3000)); o.set_avg(40, 3000)
        if step == 7:
            clone = (g.clone(), o.clone(), pos)
        if step % 6 == 0:
            p, b = g.psd()
            po, bo = o.psd()
            assert [breaks_tuple(v) for v in b] == [v.as_tuple() for v in bo]
            assert_bins_close(p, po, "deferred step %d" % step)
    p, b = g.psd()
    po, bo = o.psd()
    assert [breaks_tuple(v) for v in b] == [v.as_tuple() for v in bo]
    assert_bins_close(p, po, "deferred final")
    gc, oc, cpos = clone
    gc.process(xd[cpos:cpos + 300 * n]); oc.process(x[cpos:cpos + 300 * n])
    p, b = gc.psd(); po, bo = oc.psd()
    assert [breaks_tuple(v) for v in b] == [v.as_tuple() for v in bo]
    assert_bins_close(p, po, "deferred clone")
    g.reset()
    assert g.psd()[1] == []

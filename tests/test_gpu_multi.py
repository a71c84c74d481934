"""Partial-accumulator path of the C ABI (time-chunked stage 0) on one GPU: two handles play
two ranks, their accumulator arrays are summed like an NCCL reduction would."""
import numpy as np
import pytest

from conftest import uniform_noise

pytestmark = pytest.mark.gpu


def test_stage0_partials_sum_to_full_spectrum(oracle):
    import torch
    from stabilizer_stream_b200 import PsdCascade, multi
    n, hop = 4096, 2048
    x = uniform_noise(1001 * hop + 123, 31)
    nseg = 1 + (x.size - n) // hop
    ranges = multi.split_segments(nseg, 2)
    parts = []
    for k0, k1 in ranges:
        a, b = multi.segment_sample_range(k0, k1, n, hop)
        c = PsdCascade(n)
        c.process(torch.from_numpy(x[a:b]).cuda())
        t, counts = multi.cascade_partials_tensor(c)
        assert counts[0] == k1 - k0
        parts.append((c, t, counts))
    total = parts[0][1][0] + parts[1][1][0]          # what ncclReduce(sum) does to row 0
    full = oracle.Stage(n)
    full.process(x)
    got = total[:n // 2 + 1].cpu().numpy()
    assert np.max(np.abs(got - full.spectrum()) / full.spectrum()) < 1e-4
    # install the reduced row + count into rank 0's handle and read it back through psd()
    parts[0][1][0].copy_(total)
    torch.cuda.synchronize()
    parts[0][0].set_counts([nseg])
    from stabilizer_stream_b200 import MergeOpts
    p, b = parts[0][0].psd(MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
    top = b[-1]
    assert top.decimation == 1 and top.count == nseg
    want = full.spectrum() / full.gain()
    np.testing.assert_allclose(p[top.start:top.start + len(top.bins)], want, rtol=1e-4)


@pytest.mark.parametrize("preset", [False, True])
def test_time_chunked_over_nccl_two_ranks(preset):
    """the real thing: two processes, two GPUs, NCCL reduction (tools/run_configs.py --config 5 --check);
    skipped on a single-GPU box"""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "tools", "run_configs.py"), "--config", "5", "--total", "6e8",
           "--n-local", "4", "--check"] + (["--preset"] if preset else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["world"] == 2 and out["counts_match_closed_form"] and out["breaks_equal"]
    assert out["max_rel_diff_vs_sequential"] < 5e-5

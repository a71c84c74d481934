"""The C++ planner of the time-chunked mode (sspsd_time_plan, csrc/sspsd_group.cu -- what the group API uses)
against the Python planner it was ported from (multi.plan_time_chunks, itself checked against the cascade's
bookkeeping in tests/test_timechunk_plan_cpu.py).  Pure host code: runs without a GPU."""
import numpy as np
import pytest

from stabilizer_stream_b200 import Hbf, multi, time_plan


@pytest.mark.parametrize("n,hbf", [(4096, 1), (512, 1), (512, 0), (64, 1), (8192, 0)])
def test_cpp_plan_equals_python_plan(n, hbf):
    rng = np.random.default_rng(n + hbf)
    for _ in range(60):
        world = int(rng.integers(1, 9))
        k = int(rng.integers(1, 5))
        total = int(rng.integers(n, 40_000_000)) if rng.random() < 0.8 else int(rng.integers(4_000_000_000, 6_000_000_000))
        want = multi.plan_time_chunks(total, world, n, hbf, k)
        for r in range(world):
            got = time_plan(n, total, world, r, n_local=k, hbf=Hbf(hbf))
            w = want[r]
            assert (got.own_lo, got.own_hi, got.feed_lo, got.feed_hi, got.tail_lo, got.tail_hi, got.n_local) == \
                (w["own_lo"], w["own_hi"], w["feed_lo"], w["feed_hi"], w["tail_lo"], w["tail_hi"], k), (total, world, k, r)


def test_auto_n_local_for_config5():
    # BASELINE config 5: 4.8e9 samples over 8 GPUs -> 5 local stages (halo 0.35 % of a chunk), DESIGN.md section 5
    assert time_plan(4096, 4_800_000_000, 8, 3).n_local == 5
    assert time_plan(4096, 4_800_000_000, 1, 0).n_local >= 5
    assert time_plan(512, 100_000, 4, 0).n_local == 1
    c = [time_plan(4096, 4_800_000_000, 8, r) for r in range(8)]
    assert c[0].own_lo == 0 and c[-1].own_hi is None and all(c[r].own_hi == c[r + 1].own_lo for r in range(7))
    overhead = sum(x.feed_hi - x.feed_lo for x in c) / 4_800_000_000 - 1
    assert 0 <= overhead < 0.04

"""CPU checks of the time-chunk planner (stabilizer_stream_b200.multi): ownership partitions every stage's
segments exactly, and every rank's feed range makes its owned segments valid and complete."""
import pytest

from stabilizer_stream_b200 import multi


@pytest.mark.parametrize("total,world,n,hbf,k", [(4_000_000, 4, 512, 1, 3), (40_000_000, 8, 4096, 1, 3),
                                                  (4_800_000_000, 8, 4096, 1, 5), (3_000_000, 3, 512, 0, 2),
                                                  (1_000_000, 2, 64, 1, 3)])
def test_plan_partitions_and_halos(total, world, n, hbf, k):
    hop = n // 2
    drain, halo = multi.DRAIN[hbf], multi.DEC_HALO[hbf]
    plans = multi.plan_time_chunks(total, world, n, hbf, k)
    glob = multi.stream_state(total, n, hop, drain)
    assert plans[0]["own_lo"] == 0 and plans[-1]["own_hi"] is None
    for i in range(min(k, len(glob))):
        covered = 0
        for p in plans:
            lo = multi.first_index_at_or_after(p["own_lo"], i, hop, drain)
            hi = glob[i][1] if p["own_hi"] is None else min(glob[i][1], multi.first_index_at_or_after(p["own_hi"], i, hop, drain))
            assert lo == covered or hi <= lo, "stage %d rank %d" % (i, p["rank"])
            covered = max(covered, hi)
            if hi > lo:
                # valid: the first owned segment starts after the contaminated prefix
                v = multi.valid_from(p["feed_lo"], k, halo, drain)
                assert p["feed_lo"] == 0 or lo * hop >= v[i]
                # complete: the rank's stage i receives the last owned segment's samples
                loc = multi.stream_state(p["feed_hi"], n, hop, drain)
                assert loc[i][0] >= (hi - 1) * hop + n
        assert covered == glob[i][1], "stage %d: %d of %d segments owned" % (i, covered, glob[i][1])
    # the exported slices tile the stage-k stream
    if len(glob) > k:
        pos = 0
        for p in plans:
            hi = glob[k][0] if p["tail_hi"] is None else min(glob[k][0], p["tail_hi"])
            assert p["tail_lo"] == pos or hi <= p["tail_lo"]
            pos = max(pos, hi)
            loc = multi.stream_state(p["feed_hi"], n, hop, drain)
            assert len(loc) > k and loc[k][0] >= hi
        assert pos == glob[k][0]
    # overhead stays small for the BASELINE config
    if total == 4_800_000_000:
        extra = sum(p["feed_hi"] - p["feed_lo"] for p in plans) / total - 1
        assert extra < 0.05


def test_ewma_tail_factors_reproduce_the_sequential_recurrence():
    """Psd::process averaging (psd.rs:215-233) over S segments split among ranks: every rank applies the
    reference's rule from the count the reference would have at its first segment, then multiplies by the
    factor multi.ewma_tail_factors gives; the sum must equal the sequential recurrence."""
    import numpy as np
    from stabilizer_stream_b200 import AvgOpts
    rng = np.random.default_rng(3)
    n, hbf, hop = 64, 1, 32
    drain = multi.DRAIN[hbf]
    for total, world, limit, count in [(40_000, 3, 17, 2 ** 32 - 1), (40_000, 4, 0, 64), (90_000, 5, 999, 2 ** 32 - 2),
                                       (40_000, 2, 200, 700), (9_000, 8, 5, 2 ** 32 - 1)]:
        avg = AvgOpts(limit=limit, count=count)
        plans = multi.plan_time_chunks(total, world, n, hbf, 2)
        st = multi.stream_state(total, n, hop, drain)
        for i in range(min(2, len(st))):
            a = multi.stage_avg(avg, i)
            s_i = st[i][1]
            v = rng.random(s_i)
            # sequential
            p, c = 0.0, 0
            g32 = float(np.float32(a) / np.float32(a + 1))
            for j in range(s_i):
                if c > a:
                    p *= float(np.float32(a) / np.float32(c))
                    c = a
                c += 1
                p += v[j]
            # chunked
            tot, cnt = 0.0, None
            for pl in plans:
                lo = multi.first_index_at_or_after(pl["own_lo"], i, hop, drain)
                hi = s_i if pl["own_hi"] is None else min(s_i, multi.first_index_at_or_after(pl["own_hi"], i, hop, drain))
                q, c2 = 0.0, min(lo, a + 1)
                for j in range(lo, hi):
                    if c2 > a:
                        q *= float(np.float32(a) / np.float32(c2))
                        c2 = a
                    c2 += 1
                    q += v[j]
                f, cnts = multi.ewma_tail_factors(pl, total, n, hbf, 2, avg)
                tot += q * f[i]
                cnt = cnts[i]
            assert abs(tot - p) <= 1e-9 * max(p, 1e-30), (total, world, limit, i)
            assert cnt == c == min(s_i, a + 1)
            assert g32 <= 1.0

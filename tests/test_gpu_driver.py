"""The headless stream_test-style driver (tools/stream_psd.py) over raw and frame files."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import uniform_noise
from frames_util import make_frames, oracle_decode_stream

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stream_psd.py"), "--json"] + args,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_raw_file_source(tmp_path, oracle):
    x = uniform_noise(3_000_001, 5)
    p = tmp_path / "x.raw"
    x.tofile(p)
    out = run(["--raw", str(p), "--fft", "512", "--detrend", "midpoint", "--block", "700001"])
    o = oracle.Cascade(512, 1)
    o.set_detrend(1)
    o.process(x)
    po, bo = o.psd()
    assert out["items_per_trace"] == x.size
    assert [b["count"] for b in out["breaks"]] == [b.count for b in bo]
    assert [b["pending"] for b in out["breaks"]] == [b.pending for b in bo]
    assert out["psd_bins"] == po.size and abs(out["psd_median"] / np.median(po) - 1) < 1e-4
    assert len(out["fdev"]) > 3


def test_frame_file_source(tmp_path, oracle):
    data, flen, stride, hdrs = make_frames(1, 22, 2000, seed=4, drop_every=97)
    p = tmp_path / "frames.bin"
    p.write_bytes(data)
    out = run(["--file", str(p), "--frame-size", str(flen), "--fft", "512", "--trace", "2", "--block", str(1 << 18)])
    st, nf, lo, want = oracle_decode_stream(oracle, data, flen, stride, 2000)
    assert out["trace"] == "DAC0" and out["traces"] == 4
    assert out["loss"] == {"received": lo.received, "dropped": lo.dropped}
    o = oracle.Cascade(512, 1)
    o.set_detrend(1)
    o.process(np.concatenate(want[2]))
    assert [b["count"] for b in out["breaks"]] == [b.count for b in o.psd()[1]]


def test_noise_source_runs():
    out = run(["--noise", "-1", "--samples", "5e6", "--fft", "512"])
    assert out["items_per_trace"] == 5_000_000 and out["breaks"][-1]["count"] > 1000


def test_dsm_source_runs():
    out = run(["--dsm", str(0x1234567), "--samples", "3e6", "--fft", "512"])
    assert out["trace"] == "dsm" and out["items_per_trace"] == 3_000_000


def test_udp_source(tmp_path):
    """the reference's default source: a live UDP stream (here: loopback, resent until the driver has enough)"""
    import socket
    import time
    data, flen, stride, _ = make_frames(1, 22, 64, seed=8)
    port = 19293
    p = subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "stream_psd.py"), "--json", "--udp",
                          "127.0.0.1:%d" % port, "--samples", "200000", "--fft", "512"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    tx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    t0 = time.time()
    seq = 0
    while p.poll() is None and time.time() - t0 < 300:
        for f in range(64):
            fr = bytearray(data[f * stride:f * stride + flen])
            fr[4:8] = seq.to_bytes(4, "little")  # keep the sequence continuous across resends
            seq = (seq + 22) & 0xFFFFFFFF
            tx.sendto(bytes(fr), ("127.0.0.1", port))
        time.sleep(0.02)
    out, err = p.communicate(timeout=60)
    assert p.returncode == 0, err
    res = json.loads(out.strip().splitlines()[-1])
    assert res["traces"] == 4 and res["items_per_trace"] >= 200_000
    assert res["loss"]["received"] * 8 >= 200_000

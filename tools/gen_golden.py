#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

The reference cannot be run here (Rust, no toolchain) and ships no fixtures for src/de or src/loss
(SURVEY.md 4), so these vectors are hand-derived from the reference's formulas in exact integer /
power-of-two arithmetic, independently of the C oracle:
  * frames_*.bin + frames_*.json: frames of every format whose words are chosen so that the expected
    f32 trace values are exactly representable and can be written down by hand;
  * loss_cases.json: header sequences with the received/dropped/seq triple computed with Python ints.
Both the oracle (tests/test_oracle_decode.py, CPU) and the CUDA decoder (tests/test_gpu_decode.py)
must reproduce them bit for bit.
"""
import json
import os
import struct

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def header(fmt, batches, seq):
    return bytes([0x7B, 0x05, fmt, batches]) + struct.pack("<I", seq & 0xFFFFFFFF)


def f32bits(x):
    return int(np.float32(x).view(np.uint32))


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {}
    # ---- AdcDac (src/de/data.rs:12-82): i16 * f32(4.096*2.5/32768); DAC words are offset binary ----
    k = np.float32(4.096) * np.float32(2.5) / np.float32(32768.0)
    assert f32bits(k) == 0x39A3D70B
    adc = [0, 1, -1, 32767, -32768, 1000, -1000, 12345]
    dac_words = [0x8000, 0x8001, 0x7FFF, 0xFFFF, 0x0000, 0x83E8, 0x7C18, 0xB039]  # same values, offset binary
    payload = b""
    for b in range(2):
        payload += struct.pack("<8h", *[(v if b == 0 else -v if v != -32768 else v) for v in adc])  # ADC0
        payload += struct.pack("<8h", *adc[::-1])                                                     # ADC1
        payload += struct.pack("<8H", *dac_words)                                                     # DAC0
        payload += struct.pack("<8H", *dac_words[::-1])                                               # DAC1
    frame = header(1, 2, 0x10) + payload
    adc_b1 = [(-v if v != -32768 else v) for v in adc]
    expect = [
        [f32bits(np.float32(v) * k) for v in adc + adc_b1],
        [f32bits(np.float32(v) * k) for v in adc[::-1] + adc[::-1]],
        [f32bits(np.float32(v) * k) for v in adc + adc],
        [f32bits(np.float32(v) * k) for v in adc[::-1] + adc[::-1]],
    ]
    cases["adcdac"] = {"frame_len": len(frame), "format": 1, "batches": 2, "seq": 0x10,
                       "names": ["ADC0", "ADC1", "DAC0", "DAC1"], "expect_bits": expect}
    open(os.path.join(OUT, "frames_adcdac.bin"), "wb").write(frame)

    # ---- Fls (src/de/data.rs:85-140): [[i32;7];2] ----
    def fls_batch(re, im, ph, bi, bq):
        row0 = struct.pack("<ii", re, im) + struct.pack("<q", ph) + struct.pack("<iii", 7, 8, 9)
        row1 = struct.pack("<ii", bi, bq) + struct.pack("<5i", 1, 2, 3, 4, 5)
        return row0 + row1
    fb = [(3 << 20, 4 << 20, 1 << 16, 1 << 30, -(1 << 30)),      # AR = 5*2^20/2^31, AP = tau, BI = .5, BQ = -.5
          (0, -(1 << 31), -(1 << 40), (1 << 31) - 1, -(1 << 31)),  # AR = 1, AP = -tau*2^24, BI = 1 (rounded), BQ = -1
          (5, 12, 3, 1, -1)]
    frame = header(2, len(fb), 0xFFFFFFFE) + b"".join(fls_batch(*v) for v in fb)
    tau = np.float32(6.283185307179586)
    inv = np.float32(1.0) / np.float32(2147483648.0)
    e_ar = [np.sqrt(np.float32(re) * np.float32(re) + np.float32(im) * np.float32(im)) * inv for re, im, *_ in fb]
    e_ap = [np.float32(ph) * (tau / np.float32(65536.0)) for _, _, ph, _, _ in fb]
    e_bi = [np.float32(bi) / np.float32(2147483648.0) for *_, bi, _ in fb]
    e_bq = [np.float32(bq) / np.float32(2147483648.0) for *_, bq in fb]
    assert float(e_ar[0]) == 5 * 2.0 ** 20 / 2.0 ** 31 and float(e_ar[1]) == 1.0 and float(e_ar[2]) == 13 / 2.0 ** 31
    assert float(e_bi[0]) == 0.5 and float(e_bq[1]) == -1.0
    cases["fls"] = {"frame_len": len(frame), "format": 2, "batches": len(fb), "seq": 0xFFFFFFFE,
                    "names": ["AR", "AP", "BI", "BQ"],
                    "expect_bits": [[f32bits(v) for v in t] for t in (e_ar, e_ap, e_bi, e_bq)]}
    open(os.path.join(OUT, "frames_fls.bin"), "wb").write(frame)

    # ---- ThermostatEem (src/de/data.rs:143-164): f32 words 0, 8, 13, 16 of 20 ----
    tb = []
    for b in range(3):
        words = [np.float32(100 * b + i + 0.25) for i in range(20)]
        tb.append(words)
    frame = header(3, 3, 7) + b"".join(struct.pack("<20f", *w) for w in tb)
    cases["thermostat_eem"] = {"frame_len": len(frame), "format": 3, "batches": 3, "seq": 7,
                               "names": ["T00", "T20", "I0", "I1"],
                               "expect_bits": [[f32bits(w[i]) for w in tb] for i in (0, 8, 13, 16)]}
    open(os.path.join(OUT, "frames_thermostat_eem.bin"), "wb").write(frame)

    # ---- Mpll (src/de/data.rs:167-212): [i32;6]; phase=[4], freq=[5], amp from [0],[1] ----
    mb = [(3 << 10, 4 << 10, 0, 0, 1 << 31 - 1, 1 << 20), (0, 0, 0, 0, -(1 << 31), -(1 << 31)), (8, 15, 1, 2, 12345, -54321)]
    frame = header(4, len(mb), 0) + b"".join(struct.pack("<6i", *v) for v in mb)
    kph = tau / np.float32(4294967296.0)
    kfr = np.float32(1.0) / np.float32(1.28e-3) / np.float32(4294967296.0)
    kam = np.float32(10.24) / np.float32(10.0) * np.float32(2.0) * np.float32(2.0) / np.float32(4294967296.0)
    assert (f32bits(kph), f32bits(kfr), f32bits(kam)) == (0x30C90FDB, 0x34435000, 0x3083126E)
    e_ph = [np.float32(v[4]) * kph for v in mb]
    e_fr = [np.float32(v[5]) * kfr for v in mb]
    e_am = [np.sqrt(np.float32(v[0]) * np.float32(v[0]) + np.float32(v[1]) * np.float32(v[1])) * kam for v in mb]
    cases["mpll"] = {"frame_len": len(frame), "format": 4, "batches": len(mb), "seq": 0,
                     "names": ["phase (rad)", "frequency (kHz)", "amplitude (V/G10)"],
                     "expect_bits": [[f32bits(v) for v in t] for t in (e_ph, e_fr, e_am)]}
    open(os.path.join(OUT, "frames_mpll.bin"), "wb").write(frame)
    json.dump(cases, open(os.path.join(OUT, "frames.json"), "w"), indent=1)

    # ---- loss (src/loss.rs:11-26), Python ints ----
    seqs = [
        [(0, 22), (22, 22), (44, 22)],                                   # no loss
        [(100, 22), (122 + 5, 22), (149 + 22, 10)],                      # 5 then 22 lost
        [(0xFFFFFFF0, 22), (0x00000006, 22), (0x0000001C + 3, 22)],      # u32 wrap, then 3 lost
        [(50, 22), (10, 22)],                                            # sequence goes backwards: huge wrapping gap
        [(5, 0), (5, 0), (9, 1)],                                        # zero-batch frames
    ]
    lc = []
    for hs in seqs:
        rec = drop = 0
        prev = None
        for seq, bat in hs:
            rec += bat
            if prev is not None:
                drop += (seq - prev) % (1 << 32)
            prev = (seq + bat) % (1 << 32)
        lc.append({"headers": hs, "received": rec, "dropped": drop, "seq": prev})
    json.dump(lc, open(os.path.join(OUT, "loss_cases.json"), "w"), indent=1)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()

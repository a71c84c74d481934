"""Synthetic stabilizer frame streams for the decode/loss tests."""
import struct

import numpy as np

BATCH_BYTES = {1: 64, 2: 56, 3: 80, 4: 24}


def make_frames(fmt, batches, n_frames, seed, drop_every=0, start_seq=0, stride=None):
    """-> (bytes, frame_len, stride, headers[(seq, batches)]).  Every `drop_every`-th frame is removed
    from the stream (its sequence numbers are skipped); start_seq near 2^32 exercises the u32 wrap."""
    rng = np.random.default_rng(seed)
    flen = 8 + BATCH_BYTES[fmt] * batches
    stride = stride or flen
    out = bytearray()
    hdrs = []
    seq = start_seq
    f = 0
    while len(hdrs) < n_frames:
        f += 1
        if drop_every and f % drop_every == 0:
            seq = (seq + batches) & 0xFFFFFFFF
            continue
        if fmt == 1:
            adc = np.clip(np.rint(rng.standard_normal((batches, 2, 8)) * 3000), -32768, 32767).astype(np.int16)
            dac = (np.clip(np.rint(rng.standard_normal((batches, 2, 8)) * 3000), -32768, 32767).astype(np.int16)
                   .view(np.uint16) ^ np.uint16(0x8000))
            payload = np.concatenate([adc.view(np.uint16), dac], axis=1).astype("<u2").tobytes()
        elif fmt == 3:
            payload = rng.standard_normal((batches, 20)).astype("<f4").tobytes()
        else:
            payload = rng.integers(-2 ** 31, 2 ** 31, BATCH_BYTES[fmt] // 4 * batches, dtype=np.int64).astype("<i4").tobytes()
        fr = bytes([0x7B, 0x05, fmt, batches]) + struct.pack("<I", seq) + payload
        out += fr + bytes(stride - flen)
        hdrs.append((seq, batches))
        seq = (seq + batches) & 0xFFFFFFFF
    return bytes(out), flen, stride, hdrs


def oracle_decode_stream(orc, data, flen, stride, n_frames):
    """Loop Frame::from_bytes + Loss::update + traces() over the stream with the CPU oracle."""
    loss = orc.Loss()
    traces = None
    for f in range(n_frames):
        st, hdr, tr = orc.frame_decode(data[f * stride:f * stride + flen])
        if st != 0:
            return st, f, loss, traces
        loss.update(hdr.seq, hdr.batches)
        if traces is None:
            traces = [[] for _ in tr]
        for t, v in enumerate(tr):
            traces[t].append(v)
    return 0, n_frames, loss, traces

"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol include/sspsd.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sspsd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sspsd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_prototyped():
    from stabilizer_stream_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 30
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), "libsspsd.so does not export %s" % n
    assert sorted(_lib.PROTOTYPES) == names


def test_struct_layouts_match_header():
    from stabilizer_stream_b200 import _lib
    assert C.sizeof(_lib.BreakC) == 72 and C.sizeof(_lib.Config) == 56
    assert C.sizeof(_lib.LossC) == 24 and C.sizeof(_lib.DecodeInfoC) == 24
    assert C.sizeof(_lib.PartialsC) == 8 + 8 + 8 + 8 * 16
    assert C.sizeof(_lib.TimeChunkC) == 56 and C.sizeof(_lib.PlotOptsC) == 16


def test_host_helpers_without_gpu():
    from stabilizer_stream_b200 import Loss, Var
    v = Var().eval([1000.0, 100.0, 1.2, 3.4, 5.6], [0.0, 1.0, 3.0, 6.0, 9.0], 2.7)
    assert abs(0.13478442 - v) < 1e-6          # reference src/var.rs:52-60
    l = Loss()
    l.update(0xFFFFFFF0, 22)
    l.update(0x00000006 + 10, 22)              # wrapped: expected 0x6, 10 lost
    assert (l.received, l.dropped, l.seq) == (44, 10, 0x26)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stabilizer_stream_b200 import PsdCascade, _lib
    with pytest.raises(_lib.SspsdError) as e:
        PsdCascade(512)
    assert e.value.status == _lib.ECUDA


def test_bounds_build_carries_the_device_asserts_and_the_product_build_does_not():
    """libsspsd_bounds.so (-DSSPSD_BOUNDS) is the memory-safety build the GPU tests run (tests/test_gpu_bounds.py); make
    sure its device-side asserts are really compiled in, and that the product library is free of them."""
    import os
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stabilizer_stream_b200")
    b = os.path.join(here, "libsspsd_bounds.so")
    if not os.path.exists(b):
        pytest.skip("libsspsd_bounds.so not built")
    needles = [b"g + 4 <= s.end", b"gidx0 + r < gcap", b"<= out.cap"]
    blob = open(b, "rb").read()
    for n in needles:
        assert n in blob, n
    prod = open(os.path.join(here, "libsspsd.so"), "rb").read()
    for n in needles:
        assert n not in prod, n

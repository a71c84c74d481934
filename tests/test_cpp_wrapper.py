"""The header-only C++ mirror (include/sspsd.hpp) compiles against the C ABI and the C++ twin of the
reference's unit tests passes on a GPU; without a GPU the binary must refuse to run (exit code 3)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_psd")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib = os.path.join(ROOT, "stabilizer_stream_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_psd.cpp"), "-o", EXE, "-L", lib, "-lsspsd", "-Wl,-rpath," + lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_cpp_wrapper_compiles_and_refuses_without_gpu():
    import torch
    exe = build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_cpp_twin_of_reference_tests():
    r = subprocess.run([build()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout

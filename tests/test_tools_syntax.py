"""Every script the GPU box runs must at least compile here (no GPU needed): bench.py, __graft_entry__.py, tools/."""
import glob
import os
import py_compile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scripts_compile(tmp_path):
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    files += sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")))
    files += sorted(glob.glob(os.path.join(ROOT, "stabilizer_stream_b200", "*.py")))
    assert len(files) > 8
    for f in files:
        py_compile.compile(f, cfile=str(tmp_path / (os.path.basename(f) + "c")), doraise=True)


def test_bench_reference_arm_contract_keys():
    """bench.py --impl reference and the default arm print the keys the driver reads (static check of the source)"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"roofline"', '"cpu_baseline"', '"clocks"',
                '"gpu_launches"', '"e2e"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"impl"'):
        assert key in src, key

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


def test_reference_arm_runs_and_matches_the_gpu_arms_config():
    """the CPU arm on a tiny step count: same `config` as the GPU arm builds (same_config was false),
    one JSON line on stdout, sane scaling keys"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, SSPSD_BENCH_CPU_SAMPLES="4000000"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert out["impl"] == "reference" and out["config"] == bench.config_dict(1)
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] == 1 and out["value"] > 1
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and "parity_note" in out
    for key in ("sustained_2s", "timechunk", "e2e_small_calls", "e2e_frames", "cpu_baseline_n512", "value_readout_every_step"):
        assert '"%s"' % key in open(os.path.join(ROOT, "bench.py")).read()

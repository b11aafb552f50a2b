"""bench.py contract (CPU side): the reference arm prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")        # what torchrun exports; the arm must override it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-channels", "8"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "sample" in cb
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback anywhere on the product path: on a box without CUDA `python bench.py` exits non-zero with a clear message
    instead of timing something else."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_ops_fail_loudly_when_the_library_is_missing():
    """MVWARP_LIB pointing nowhere: loading raises MvwarpError naming the build command; nothing falls back to torch / the oracle."""
    code = ("import torch, mvdet_b200 as mv\n"
            "from mvdet_b200 import _lib\n"
            "try:\n"
            "    _lib.load()\n"
            "except _lib.MvwarpError as e:\n"
            "    print('RAISED', 'mvdet_b200.build' in str(e), 'fallback' in str(e))\n"
            "try:\n"
            "    mv.warp_views(torch.zeros(1, 1, 4, 8, 8), torch.eye(3)[None].double(), [4, 4])\n"
            "except mv.MvwarpError as e:\n"
            "    print('CPU TENSOR REFUSED')\n")
    env = dict(os.environ, MVWARP_LIB="/nonexistent/libmvwarp.so", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "RAISED True True" in out.stdout and "CPU TENSOR REFUSED" in out.stdout


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mvdet_b200/ may import it (bench.py's cpu_baseline / --impl reference legs and
    __graft_entry__.smoke() are the only other users)."""
    import re
    pkg = os.path.join(ROOT, "mvdet_b200")
    for name in sorted(os.listdir(pkg)):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), name
    code = "import sys, mvdet_b200, mvdet_b200.ops, mvdet_b200.integration, mvdet_b200.conv1, mvdet_b200.map_head, mvdet_b200.sharding\n" \
           "print([m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')])\n"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=ROOT), timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "[]", out.stdout + out.stderr

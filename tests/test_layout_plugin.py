"""Constants outside the built-in layout list: the reference tells its user to edit ReplicaCount / Values /
StartViewOnTimerLimit in VSR.cfg (README.md:11-18); the loader compiles the packed layout for such constants on first
use (csrc/vsr_layout_plugin.cu -> vsr-tlaplus_b200/layouts/) and it must meet the same parity bar as a built-in one."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plugin_layout_matches_oracle_on_the_host(pkg):
    """|Values| = 4 is not in VSR_FOR_EACH_CONFIG: loaded through the plug-in path, then the complete state space of
    (R=2, 4 values, limit 2; 7135 states) and random walks are compared with the oracle successor by successor."""
    mc = pkg.ModelChecker.from_constants(2, 4, 2)
    assert mc.state_bytes % 16 == 0
    assert os.path.exists(os.path.join(ROOT, "vsr-tlaplus_b200", "layouts", "libvsr_layout_2_4_3.so"))
    exe = os.path.join(ROOT, "build", "diff_host")
    for sym in (1, 0):
        out = json.loads(subprocess.run([exe, "2", "4", "2", str(sym), "8000", "1", "100000", "3"], check=True, capture_output=True, text=True).stdout)
        assert out["mismatches"] == 0 and out["checked"] == 8000, out
        assert out["assumption_violations"] == 0


def test_start_view_on_timer_limit_zero(pkg):
    """StartViewOnTimerLimit = 0 (no timer-triggered view change: the view-change and state-transfer candidate groups are
    EMPTY in the packed layout) is a legal constant; complete space (3 replicas, 2 values: 129 states without SYMMETRY)
    and walks against the oracle"""
    pkg.ModelChecker.from_constants(3, 2, 0)
    exe = os.path.join(ROOT, "build", "diff_host")
    for sym in (1, 0):
        out = json.loads(subprocess.run([exe, "3", "2", "0", str(sym), "100000"], check=True, capture_output=True, text=True).stdout)
        assert out["complete"] == 1 and out["mismatches"] == 0 and out["assumption_violations"] == 0, out
    assert out["checked"] == 129


def _load_in_subprocess(R, V, L, env):
    code = ("import sys; sys.path.insert(0, %r); import _pkg; pkg = _pkg.load()\n"
            "try:\n    pkg.ModelChecker.from_constants(%d, %d, %d); print('LOADED')\n"
            "except Exception as e:\n    print('ERROR', e)\n" % (ROOT, R, V, L))
    e = dict(os.environ)
    e.update(env)
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e).stdout


def test_unbuilt_layout_with_jit_forbidden_says_so(pkg):
    out = _load_in_subprocess(2, 5, 1, {"VSR_B200_JIT": "0"})
    assert "ERROR" in out and "VSR_B200_JIT=0" in out and "ReplicaCount=2 |Values|=5" in out, out


def test_constants_beyond_the_encoding_are_refused(pkg):
    out = _load_in_subprocess(8, 1, 1, {})
    assert "ERROR" in out and "outside the packed encoding's range" in out, out


def test_compile_failure_is_reported_not_hidden(pkg):
    """A broken compiler path must surface as a config error naming the compiler, never as a silent fallback."""
    out = _load_in_subprocess(2, 6, 1, {"VSR_B200_NVCC": "/nonexistent/nvcc"})
    assert "ERROR" in out and "compiling it failed" in out and "/nonexistent/nvcc" in out, out


@pytest.mark.gpu
def test_plugin_layout_full_state_space_on_the_gpu(pkg):
    """Complete BFS of (R=2, 4 values, limit 2) on the GPU through a plug-in layout, per-depth state sets against the
    oracle.  Runs in a child process: the plug-in's kernels have not been on a GPU yet, and a crash there must not
    take the rest of the GPU suite with it."""
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "import _pkg; pkg = _pkg.load()\n"
            "from test_gpu_parity import assert_same_exploration, run_pair\n"
            "mc, res, q, o = run_pair(pkg, 2, 4, 2)\n"
            "assert res.rc == 0 and res.distinct == 7135, res\n"
            "assert_same_exploration(pkg, mc, res, q, o, complete=True)\n"
            "print('PLUGIN-GPU-OK')\n" % (ROOT, os.path.join(ROOT, "tests")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "PLUGIN-GPU-OK" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])

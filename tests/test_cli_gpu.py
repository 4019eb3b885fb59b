"""`vsrmc` — TLC's command line — on a real GPU: exit statuses, summary lines, -deadlock, -depth, -dumpTrace."""
import os
import subprocess

import pytest

import orc
from conftest import ROOT

pytestmark = pytest.mark.gpu
VSRMC = os.path.join(ROOT, "vsr-tlaplus_b200", "vsrmc")


def run(args, tmp_path, cfg, env=None):
    p = tmp_path / "m.cfg"
    p.write_text(cfg)
    r = subprocess.run([VSRMC, "-config", str(p), "-table", "1048576", "-frontier", "200000"] + args, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, **(env or {})))
    return r.returncode, r.stdout + r.stderr


def test_cli_complete_run_and_tlc_summary_lines(pkg, tmp_path):
    rc, out = run(["-deadlock"], tmp_path, pkg.cfg_text(2, ["v1"], 1))
    assert rc == 0
    assert "Model checking completed. No error has been found." in out
    assert "100 states generated, 76 distinct states found, 0 states left on queue." in out
    assert "The depth of the complete state graph search is 14." in out


def test_cli_checks_deadlock_by_default_like_tlc(pkg, tmp_path):
    """TLC checks deadlock unless -deadlock is given; VSR.tla has reachable terminal states"""
    rc, out = run([], tmp_path, pkg.cfg_text(2, ["v1"], 1))
    assert rc == 11 and "Error: Deadlock reached." in out
    o = orc.bfs(orc.params(2, 1, 1, symmetry=False), workers=1, check_deadlock=True, keep_trace=False)
    assert o.rc == 11 and f"The depth of the state graph search so far is {o.depth}." in out


def test_cli_violation_exit_12_and_dumptrace(pkg, tmp_path):
    cfg = pkg.cfg_text(3, ["v1", "v2"], 1, invariants=["AcknowledgedWritesExistOnMajority"])
    dump = tmp_path / "trace.txt"
    rc, out = run(["-deadlock", "-dumpTrace", "tlc", str(dump)], tmp_path, cfg)
    assert rc == 12 and "Error: Invariant AcknowledgedWritesExistOnMajority is violated." in out
    text = dump.read_text()
    assert text.startswith("<<\n[\n _TEAction |-> [\n   position |-> 1,\n   name |-> \"Initial predicate\"") and text.endswith(">>")
    # the dumped file parses with the oracle's TLC-text parser and replays through its Next
    r = subprocess.run([os.path.join(ROOT, "oracle", "_build", "vsr_oracle"), "replay", str(dump)], capture_output=True, text=True)
    assert "NOT A STEP" not in r.stdout and r.stdout.count(" ok (") == 18


def test_cli_depth_bound(pkg, tmp_path):
    rc, out = run(["-deadlock", "-depth", "5"], tmp_path, pkg.cfg_text(3, ["v1", "v2"], 2))
    assert rc == 0 and "The depth of the state graph search so far is 5." in out and "173 distinct states found" in out


def test_cli_gpus_flag_shards_the_search_and_names_the_violated_invariant(pkg, tmp_path):
    """`-gpus N` (TLC's -workers, here: fingerprint-sharded GPUs).  On a one-GPU box the test hook puts the ranks on device 0.
    Two invariants configured, the SECOND one fails: the message must name that one."""
    cfg = pkg.cfg_text(3, ["v1", "v2"], 1, invariants=["AcknowledgedWriteNotLost", "AcknowledgedWritesExistOnMajority"])
    one = {"VSR_B200_MULTI_ONE_DEVICE": "1"}
    rc1, out1 = run(["-deadlock"], tmp_path, cfg)
    rc4, out4 = run(["-deadlock", "-gpus", "4"], tmp_path, cfg, env=one)
    assert rc1 == rc4 == 12
    for out in (out1, out4):
        assert "Error: Invariant AcknowledgedWritesExistOnMajority is violated." in out
        assert "Invariant AcknowledgedWriteNotLost is violated" not in out
    assert "on GPUs 0..3" in out4
    line = [ln for ln in out1.splitlines() if "distinct states found" in ln]
    assert line and line == [ln for ln in out4.splitlines() if "distinct states found" in ln]


def test_cli_honours_check_deadlock_false_in_the_cfg(pkg, tmp_path):
    rc, out = run([], tmp_path, pkg.cfg_text(2, ["v1"], 1) + "CHECK_DEADLOCK FALSE\n")
    assert rc == 0 and "Model checking completed. No error has been found." in out

"""vsrmc's command line where no GPU is needed: flag handling, config errors, and — on a machine without CUDA — the loud
status 153 (there is no CPU fallback behind the CLI either)."""
import os
import subprocess

import pytest
import torch

from conftest import ROOT

EXE = os.path.join(ROOT, "vsr-tlaplus_b200", "vsrmc")


def run(args, tmp_path, cfg_text):
    cfg = tmp_path / "VSR.cfg"
    cfg.write_text(cfg_text)
    r = subprocess.run([EXE] + args + ["-config", str(cfg)], capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr


def test_tlc_housekeeping_flags_are_accepted(pkg, tmp_path):
    """a TLC command line with -workers / -metadir / -cleanup / -coverage keeps working: the flags have no counterpart here"""
    rc, out = run(["-deadlock", "-workers", "auto", "-metadir", str(tmp_path), "-cleanup", "-nowarning", "-coverage", "1"], tmp_path,
                  pkg.cfg_text(2, ["v1"], 1))
    assert "unrecognized option" not in out
    assert rc == (0 if torch.cuda.is_available() else 153), out


def test_modes_that_do_not_exist_are_refused_by_name(pkg, tmp_path):
    for flag in (["-dfid", "10"], ["-generateSpecTE"]):
        rc, out = run(["-deadlock"] + flag, tmp_path, pkg.cfg_text(2, ["v1"], 1))
        assert rc == 255 and flag[0] + " is not available" in out
    rc, out = run(["-bogus"], tmp_path, pkg.cfg_text(2, ["v1"], 1))
    assert rc == 255 and "unrecognized option -bogus" in out


def test_config_error_is_exit_151(pkg, tmp_path):
    rc, out = run(["-deadlock"], tmp_path, pkg.cfg_text(2, ["v1"], 1) + "PROPERTY ViewChangeCompletes\n")
    assert rc == 151 and "PROPERTY" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_without_cuda_the_cli_fails_loudly(pkg, tmp_path):
    rc, out = run(["-deadlock"], tmp_path, pkg.cfg_text(2, ["v1"], 1))
    assert rc == 153 and "153" in out and "states generated" in out  # the summary shows 0 states: nothing ran anywhere else

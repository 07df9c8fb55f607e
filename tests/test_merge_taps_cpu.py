"""The static tap tables of the scale-2 merge kernel (csrc/merge_taps.h), checked on the host with plain g++ against the
reference's per-tap index arithmetic (DeBayerKernels.cu:398-437) for every residue of (X, Y, sx, sy)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_merge_tap_tables(tmp_path):
    exe = tmp_path / "merge_taps_check"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "host" / "merge_taps_check.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad 0" in out.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_merge_slot_fold(tmp_path):
    """csrc/merge_slots.h (the predicate-free slot fold of merge_pf.cu) against the reference's per-tap index arithmetic."""
    exe = tmp_path / "merge_slots_check"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "host" / "merge_slots_check.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad 0" in out.stdout

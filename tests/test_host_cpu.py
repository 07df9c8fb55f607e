"""C++ host program over the C ABI (host/multi_frame_sr_b200.cpp): builds with g++ against include/mfsr.h alone, keeps the
reference program's command line and messages (multi_frame_sr.cpp:122-163), and fails loudly without a GPU."""
import subprocess

import numpy as np
import pytest

from multi_frame_super_resolution_b200 import build as B


def _write_pgm16(path, img):
    with open(path, "wb") as f:
        f.write(f"P5\n# raw frame\n{img.shape[1]} {img.shape[0]}\n1023\n".encode())
        f.write(img.astype(">u2").tobytes())


@pytest.fixture(scope="module")
def host_bin():
    return str(B.build_host())


def test_usage_and_wrong_input(host_bin, tmp_path):
    r = subprocess.run([host_bin, "farneback", "city"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 255 and "optFlowName inputName iterations" in r.stdout       # return -1 (:143)
    r = subprocess.run([host_bin, "farneback", "nosuchset", "3"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 255 and "wrong input" in r.stdout                              # :161-162
    r = subprocess.run([host_bin, "farneback", "city", "3"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 255 and "cannot read" in r.stdout


def test_reads_frames_and_fails_loudly_without_gpu(host_bin, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_host_gpu.py")
    rng = np.random.default_rng(0)
    for i in range(1, 6):
        _write_pgm16(tmp_path / f"img_{i:06d}.pgm", rng.integers(64, 1023, (96, 128)))
    r = subprocess.run([host_bin, "pyrlk", "city", "2"], capture_output=True, text=True, cwd=tmp_path)
    assert "img_000001.pgm, [128 x 96]" in r.stdout and "img_000005.pgm" in r.stdout      # the reference's per-frame echo (:175)
    assert r.returncode == 1 and "no compute-capability-10.x CUDA device" in r.stderr      # no CPU fallback
    assert not list(tmp_path.glob("*_result.ppm"))

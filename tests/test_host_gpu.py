"""The C++ host program (host/multi_frame_sr_b200.cpp) run as a process on the GPU box: its 8-bit result must equal the image the
Python mirror gets from the same library for the same frames, and its sharpened copy must follow sharpenImg2."""
import subprocess

import numpy as np
import pytest
import torch

from multi_frame_super_resolution_b200 import build as B
from multi_frame_super_resolution_b200.cli import _sharpen
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst

pytestmark = pytest.mark.gpu


def _read_ppm(path):
    data = open(path, "rb").read()
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P6" and parts[2] == b"255"
    w, h = map(int, parts[1].split())
    return np.frombuffer(parts[3], np.uint8).reshape(h, w, 3)


def test_host_program_matches_python_mirror(cuda_device, tmp_path):
    host = str(B.build_host())
    n, h, w = 5, 192, 256
    fr, _ = synth_burst(n, h, w, seed=21)
    raw = fr.numpy().view(np.uint16)
    for i in range(n):
        with open(tmp_path / f"img_{i + 1:06d}.pgm", "wb") as f:
            f.write(f"P5\n{w} {h}\n1023\n".encode())
            f.write(raw[i].astype(">u2").tobytes())
    r = subprocess.run([host, "farneback", "city", "3"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == f"img_000001.pgm, [{w} x {h}]" and lines[n].endswith(" sec") and lines[n + 1].endswith(" FPS")
    got = _read_ppm(tmp_path / "city_farneback_sr_result.ppm")
    got2 = _read_ppm(tmp_path / "city_farneback_sr2_result.ppm")
    p = default_params()
    p.lk_iterations = 3
    p.merge_flags = 1
    while p.levels > 1 and min(h, w) >> (p.levels - 1) < 2 * p.max_shift + p.tile_size:
        p.levels -= 1
    sr = BurstSuperResolution(p, 0, w, h, n)
    sr.set_input(raw)                                  # host frames, like the C++ program
    exp = sr.next_frame(host=True, dtype=torch.uint8).numpy()
    sr.close()
    assert got.shape == (2 * h, 2 * w, 3) and np.array_equal(got, exp)
    assert np.array_equal(got2, _sharpen(exp))


def test_temporal_area_mode_python(cuda_device):
    """multi_frame_sr.cpp:182-194: radius 1, one result per frame of the sequence, each merged from frames [i-1, i+1] onto frame i;
    None after the last frame; reset() rewinds."""
    n, h, w = 5, 128, 192
    fr, _ = synth_burst(n, h, w, seed=8)
    dev = fr.cuda()
    p = default_params()
    p.levels = 2
    sr = BurstSuperResolution(p, 0, w, h, 3)
    sr.set_temporal_area_radius(1)
    sr.set_input(dev)
    outs = []
    while True:
        o = sr.next_frame()
        if o is None:
            break
        outs.append(o.clone())
    assert len(outs) == n and sr.next_frame() is None
    sr.reset()
    again = sr.next_frame()
    assert torch.equal(again, outs[0])
    whole = BurstSuperResolution(p, 0, w, h, 3)
    for i in range(n):
        lo, hi = max(0, i - 1), min(n, i + 2)
        whole.set_input(dev[lo:hi], ref_idx=i - lo)
        assert torch.equal(whole.next_frame(), outs[i]), i
    with pytest.raises(ValueError):
        sr.set_temporal_area_radius(2)
        sr.set_input(dev)                                      # window of 5 > max_frames 3
    sr.close()
    whole.close()


def test_host_program_temporal_area(cuda_device, tmp_path):
    host = str(B.build_host())
    n, h, w = 5, 128, 192
    fr, _ = synth_burst(n, h, w, seed=9)
    raw = fr.numpy().view(np.uint16)
    for i in range(n):
        with open(tmp_path / f"img_{i + 1:06d}.pgm", "wb") as f:
            f.write(f"P5\n{w} {h}\n1023\n".encode())
            f.write(raw[i].astype(">u2").tobytes())
    r = subprocess.run([host, "--radius", "1", "farneback", "city", "2"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stderr
    assert '"temporal_radius": 1' in r.stdout.strip().splitlines()[-1]
    got = _read_ppm(tmp_path / "city_farneback_sr_result.ppm")
    # the last result of the 50-frame sequence: frame 49 = set frame 5, merged from sequence frames 48, 49 onto 49
    p = default_params()
    p.lk_iterations = 2
    p.merge_flags = 1
    while p.levels > 1 and min(h, w) >> (p.levels - 1) < 2 * p.max_shift + p.tile_size:
        p.levels -= 1
    sr = BurstSuperResolution(p, 0, w, h, 3)
    sr.set_input(np.ascontiguousarray(raw[3:5]), ref_idx=1)
    exp = sr.next_frame(host=True, dtype=torch.uint8).numpy()
    sr.close()
    assert np.array_equal(got, exp)

"""numpy-only synthetic burst for the CPU arm of bench.py (`--impl reference` and the `cpu_baseline` leg).

TEST / BASELINE INFRASTRUCTURE ONLY.  Same workload class as multi_frame_super_resolution_b200/synth.py (a textured scene, per-frame
sub-pixel translation of up to +-3 px, RGGB mosaic, signal-dependent noise, 10-bit quantisation with black level 64) without
importing the product package, so that the reference arm of the bench loads nothing but oracle/.  The pixel values differ from
the torch generator's (different random streams); the oracle's run time depends on the burst's dimensions, not its content."""
from __future__ import annotations

import numpy as np


def synth_burst_np(n_frames: int, height: int, width: int, seed: int = 1234, max_shift: float = 3.0,
                   black: int = 64, white: int = 1023, alpha: float = 1e-3, beta: float = 1e-5) -> tuple[np.ndarray, np.ndarray]:
    """Returns (frames uint16 [N, H, W], shifts float32 [N, 2])."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    pad = int(np.ceil(max_shift)) + 2
    hh, ww = height + 2 * pad, width + 2 * pad
    scene = np.zeros((hh, ww), np.float32)
    for sigma, amp in ((48.0, 0.30), (12.0, 0.22), (3.0, 0.16), (1.0, 0.10)):       # 1/f-like octaves of filtered noise
        layer = ndimage.gaussian_filter(rng.standard_normal((hh, ww)).astype(np.float32), sigma, mode="wrap")
        scene += amp * layer / (layer.std() + 1e-9)
    ys, xs = np.mgrid[0:hh, 0:ww].astype(np.float32)
    for _ in range(6):                                                              # a few straight edges
        th, off, c = rng.uniform(0, np.pi), rng.uniform(0.2, 0.8), rng.uniform(-0.25, 0.25)
        scene += c * ((np.cos(th) * xs / ww + np.sin(th) * ys / hh) > off)
    scene = 0.5 + 0.45 * np.tanh(scene)
    shifts = rng.uniform(-max_shift, max_shift, size=(n_frames, 2)).astype(np.float32)
    shifts[0] = 0
    frames = np.empty((n_frames, height, width), np.uint16)
    for f in range(n_frames):
        img = ndimage.shift(scene, (-shifts[f, 1], -shifts[f, 0]), order=1, mode="nearest")[pad:pad + height, pad:pad + width]
        img = img + rng.standard_normal(img.shape).astype(np.float32) * np.sqrt(alpha * img + beta)
        frames[f] = np.clip(np.round(img * (white - black) + black), 0, white).astype(np.uint16)
    return frames, shifts

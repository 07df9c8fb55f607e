"""Seeded synthetic bursts: an analytic scene rendered at sub-pixel offsets.

Scene = multi-octave value-noise texture (non-periodic, ~1/f spectrum like natural images:
block matching has one minimum) + random sinusoids (area-sampled exactly: amplitude x sinc
of the pixel aperture, so frequencies above the LR Nyquist alias the way a real sensor
aliases) + soft straight edges + a slow colour tint.  A scene made of global plane waves
alone is self-similar under translation and makes the tile matcher lock onto wrong periods
for some seeds (tools/flow_stats2.py), which no camera scene does; per-frame motion = global translation U(-a,a)
plus a long-wavelength warp; noise sigma^2 = alpha*I + beta; 10-bit quantisation in a u16
container (black 64, white 1023); RGGB mosaic or gray.  Runs on CPU (tests) or CUDA (bench)
with identical code; there is no network, so this replaces real captures.
"""
from __future__ import annotations

import math

import torch


def synth_burst(n_frames: int, height: int, width: int, seed: int = 1234, device="cpu", bayer: bool = True,
                max_shift: float = 3.0, warp_amp: float = 0.4, alpha: float = 1e-3, beta: float = 1e-5,
                black: int = 64, white: int = 1023, n_sines: int = 24, n_edges: int = 12, noise: bool = True):
    """Returns (frames u16 [N,H,W] on `device`, shifts float32 [N,2] = (dx,dy) of frame f relative to frame 0)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = torch.device(device)

    def rnd(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float64)

    # sinusoids: frequency in cycles per LR pixel, up to 0.9 (0.45 cycles / HR pixel at 2x)
    fr = rnd(n_sines) * 0.9
    th = rnd(n_sines) * 2 * math.pi
    fx, fy = fr * torch.cos(th), fr * torch.sin(th)
    ph = rnd(n_sines) * 2 * math.pi
    amp = (0.5 + rnd(n_sines)) / (1.0 + 6.0 * fr)
    amp = amp * torch.sinc(fx) * torch.sinc(fy)            # exact area sampling over a 1x1 LR pixel
    # edges: unit normal, offset, contrast
    en = rnd(n_edges) * 2 * math.pi
    eo = rnd(n_edges)
    ec = (rnd(n_edges) - 0.5) * 0.6
    shifts = (rnd(n_frames, 2) * 2 - 1) * max_shift
    shifts[0] = 0
    wph = rnd(n_frames, 2) * 2 * math.pi
    # value-noise octaves: cell size 192 .. 3 LR pixels, amplitude ~ sqrt(cell) (1/f-like), bilinear evaluation
    octaves = []
    cell = 192.0
    while cell >= 3.0:
        gw, gh = int(width / cell) + 3, int(height / cell) + 3
        grid = torch.rand((1, 1, gh, gw), generator=g, dtype=torch.float32) * 2 - 1
        octaves.append((cell, grid.to(device), math.sqrt(cell / 192.0)))
        cell /= 2.0
    onorm = math.sqrt(sum(a * a for _, _, a in octaves))

    ys = torch.arange(height, device=dev, dtype=torch.float32).view(-1, 1)
    xs = torch.arange(width, device=dev, dtype=torch.float32).view(1, -1)
    gn = torch.Generator(device=dev).manual_seed(seed + 7919)
    frames = torch.empty((n_frames, height, width), dtype=torch.int16, device=dev)
    diag = float(max(width, height))
    for f in range(n_frames):
        dx = float(shifts[f, 0]) + warp_amp * torch.sin(2 * math.pi * ys / 512.0 + float(wph[f, 0]))
        dy = float(shifts[f, 1]) + warp_amp * torch.sin(2 * math.pi * xs / 640.0 + float(wph[f, 1]))
        X = xs + dx
        Y = ys + dy
        img = torch.full((height, width), 0.5, device=dev, dtype=torch.float32)
        acc = torch.zeros_like(img)
        for k in range(n_sines):
            acc += float(amp[k]) * torch.sin(2 * math.pi * (float(fx[k]) * X + float(fy[k]) * Y) + float(ph[k]))
        img += 0.10 * acc / math.sqrt(n_sines / 8.0)
        tex = torch.zeros_like(img)
        for cell, grid, a in octaves:
            gh, gw = grid.shape[-2:]
            # grid node i sits at LR coordinate (i - 1) * cell; align_corners=True maps [-1, 1] onto nodes 0 .. n-1
            u = ((X / cell + 1.0) / (gw - 1)) * 2 - 1
            v = ((Y / cell + 1.0) / (gh - 1)) * 2 - 1
            tex += a * torch.nn.functional.grid_sample(grid, torch.stack([u, v], dim=-1).unsqueeze(0), mode="bilinear",
                                                       padding_mode="border", align_corners=True)[0, 0]
        img += 0.22 * tex / onorm
        for k in range(n_edges):
            d = math.cos(float(en[k])) * X + math.sin(float(en[k])) * Y - float(eo[k]) * diag
            img += float(ec[k]) * torch.sigmoid(d / 0.35)
        img = img.clamp(0.05, 0.95)
        if bayer:
            tint = [0.85 + 0.15 * torch.sin(2 * math.pi * (X / 900.0 + Y / 1300.0) + 2.1 * c) for c in range(3)]
            ypar = (torch.arange(height, device=dev) % 2).view(-1, 1)
            xpar = (torch.arange(width, device=dev) % 2).view(1, -1)
            col = ypar + xpar                                   # RGGB: 0 -> R, 1 -> G, 2 -> B
            img = torch.where(col == 0, img * tint[0], torch.where(col == 1, img * tint[1], img * tint[2]))
        if noise:
            sigma = torch.sqrt(alpha * img + beta)
            img = img + sigma * torch.randn(img.shape, generator=gn, device=dev, dtype=torch.float32)
        q = torch.round(black + img * (white - black)).clamp(0, white)
        frames[f] = q.to(torch.int16)
    return frames, shifts.to(torch.float32)


def synth_merge_inputs(n_frames: int, height: int, width: int, seed: int = 1234, device="cpu", flow_amp: float = 2.5):
    """Seeded inputs for the merge stage alone: raw u16 [N,H,W], mask [N,H/2,W/2,4], flow [N,H,W,2], kernel4 [H,W,4].

    Flow is smooth (low-frequency sinusoids + a per-frame offset), masks are smooth fields in [0,1]
    with hard zero regions, kernel params are positive-definite inverse covariances."""
    raw, _ = synth_burst(n_frames, height, width, seed=seed, device=device)
    dev = torch.device(device)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    ys = torch.arange(height, device=dev, dtype=torch.float32).view(-1, 1)
    xs = torch.arange(width, device=dev, dtype=torch.float32).view(1, -1)
    flow = torch.empty((n_frames, height, width, 2), dtype=torch.float32, device=dev)
    mask = torch.empty((n_frames, height // 2, width // 2, 4), dtype=torch.float32, device=dev)
    yh, xh = ys[: height // 2], xs[:, : width // 2]
    for f in range(n_frames):
        r = torch.rand(8, generator=g)
        off = (r[:2] * 2 - 1) * flow_amp
        flow[f, ..., 0] = float(off[0]) + 0.8 * torch.sin(2 * math.pi * (xs / 300.0 + ys / 517.0) + float(r[2]) * 6.28)
        flow[f, ..., 1] = float(off[1]) + 0.8 * torch.cos(2 * math.pi * (xs / 411.0 - ys / 289.0) + float(r[3]) * 6.28)
        base = 0.5 + 0.5 * torch.sin(2 * math.pi * (xh / 97.0 + yh / 131.0) + float(r[4]) * 6.28)
        hole = (torch.sin(2 * math.pi * (xh / 61.0) + float(r[5]) * 6.28) * torch.sin(2 * math.pi * (yh / 73.0)) > 0.7)
        for c in range(3):
            m = (base * (0.8 + 0.2 * math.cos(c + float(r[6])))).clamp(0, 1)
            mask[f, ..., c] = torch.where(hole, torch.zeros_like(m), m)
        mask[f, ..., 3] = 0.1 * base
    ang = 2 * math.pi * (xs / 700.0 + ys / 900.0)
    k1 = 0.35 + 0.3 * torch.sin(2 * math.pi * xs / 173.0) ** 2 + 0 * ys
    k2 = 0.15 + 0.1 * torch.cos(2 * math.pi * ys / 211.0) ** 2 + 0 * xs
    c, s = torch.cos(ang), torch.sin(ang)
    b11 = k1 * c * c + k2 * s * s
    b22 = k1 * s * s + k2 * c * c
    b12 = (k1 - k2) * c * s
    det = b11 * b22 - b12 * b12
    kernel4 = torch.stack([b22 / det, b11 / det, -b12 / det, torch.zeros_like(det)], dim=-1).contiguous()
    return raw, mask.contiguous(), flow.contiguous(), kernel4

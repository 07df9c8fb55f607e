"""How many output pixel-frames fall outside the merge kernel's staged raw window (tile mean shift +- 13 raw px / +- 6 raw rows), per
burst seed, and how certain the robustness model is about them."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
import ctypes as C

n, h, w = 8, 3024, 4032
dev = torch.device("cuda", 0)
sr = BurstSuperResolution(default_params(), device=0, max_width=w, max_height=h, max_frames=n)
for seed in [int(a) for a in sys.argv[1:]] or [1234, 1235]:
    fr, sh = synth_burst(n, h, w, seed=seed, device=dev)
    sr.set_input(fr); sr.next_frame(); sr.synchronize()
    tot_out, tot = 0, 0
    for f in range(1, n):
        flow = torch.from_numpy(sr.buffer("flow", h, w * 8, f).view(np.float32).reshape(h, w, 2)).to(dev)
        mask = torch.from_numpy(sr.buffer("mask", h // 2, (w // 2) * 16, f).view(np.float32).reshape(h // 2, w // 2, 4)).to(dev)
        s2 = torch.round(2 * flow)                                   # raw-resolution proxy of the per-pixel HR shift
        # tile = 64 x 8 raw pixels (128 x 16 HR)
        th, tw = h // 8, w // 64
        t = s2[: th * 8, : tw * 64].reshape(th, 8, tw, 64, 2)
        mean = t.mean(dim=(1, 3), keepdim=True).round()
        dx = (t[..., 0] - mean[..., 0]).abs(); dy = (t[..., 1] - mean[..., 1]).abs()
        out = (dx > 24) | (dy > 10)
        m = mask[..., :3].max(-1).values
        mfull = m.repeat_interleave(2, 0).repeat_interleave(2, 1)[: th * 8, : tw * 64].reshape(th, 8, tw, 64)
        tot_out += int(out.sum()); tot += out.numel()
        if f in (1, 4, 7):
            print(f"  seed {seed} frame {f}: outside window {out.float().mean().item():.4f}; of those certainty>0: {(mfull[out] > 0).float().mean().item() if out.any() else 0:.3f}; "
                  f"|flow| p50 {flow.norm(dim=-1).median().item():.2f} p99.9 {torch.quantile(flow.norm(dim=-1).flatten()[::97], 0.999).item():.1f}; tiles with any outlier {(out.any(dim=3).any(dim=1)).float().mean().item():.3f}")
    print(f"seed {seed}: pixel-frames outside the window {tot_out / tot:.4f}, merge {sr.stage_ms()['merge']:.2f} ms, consolidate {sr.stage_ms()['consolidate']:.2f}")

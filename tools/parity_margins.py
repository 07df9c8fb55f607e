import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
from oracle import pyoracle as O
from util import psnr, u16
for seed in (77, 78, 79):
  for ff in (1, 0):
    p = default_params(); p.full_frame = ff; p.levels = 3
    fr, sh = synth_burst(5, 256, 320, seed=seed)
    sr = BurstSuperResolution(p, 0, 320, 256, 5); sr.set_input(fr.cuda(), ref_idx=2); out = sr.next_frame().cpu().numpy(); sr.synchronize()
    exp, it = O.run_pipeline(u16(fr), p, ref_idx=2, keep=True)
    h, w = 256, 320
    fl99 = max(float(np.percentile(np.abs(sr.buffer("flow", h, w * 8, f).view(np.float32).reshape(h, w, 2) - it["flow"][f]), 99.9)) for f in range(5))
    flmax = max(float(np.abs(sr.buffer("flow", h, w * 8, f).view(np.float32).reshape(h, w, 2) - it["flow"][f]).max()) for f in range(5))
    mk = max(float((np.abs(sr.buffer("mask", h // 2, (w // 2) * 16, f).view(np.float32).reshape(h // 2, w // 2, 4) - it["mask"][f]) > 1e-3).mean()) for f in range(5))
    both = np.isfinite(out) & np.isfinite(exp)
    d = np.abs(np.where(both, out, 0) - np.where(both, exp, 0))
    print(f"seed {seed} full_frame {ff}: flow 99.9pct {fl99:.2e} (limit 5e-3) max {flmax:.3g}; mask frac>1e-3 {mk:.2e} (limit 5e-3); image frac>1e-3 {(d > 1e-3).mean():.2e} (limit 2e-3) max {d.max():.3g} psnr {psnr(np.where(both, out, 0), np.where(both, exp, 0)):.1f} dB (limit 50)")
    sr.close()

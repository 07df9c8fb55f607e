"""Where does the CUDA chain leave the oracle's chain on the bundled burst (config 1)?  Per-frame flow / mask differences."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from oracle import pyoracle as O

fr_np = np.load(ROOT / "tests" / "golden" / "bundled_burst_rggb.npz")["frames"]
n, h, w = fr_np.shape
p = default_params()
p.prealign = 1
sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
sr.set_input(torch.from_numpy(fr_np.view(np.int16)).cuda(), ref_idx=0)
out = sr.next_frame().cpu().numpy()
exp, it = O.run_pipeline(fr_np, p, ref_idx=0, keep=True)
bad = np.abs(out - exp) > 1e-3
print("image: frac beyond 1e-3", bad.mean(), "max", np.abs(out - exp).max(), "psnr", 10 * np.log10(1.0 / np.mean((out.astype(np.float64) - exp) ** 2)))
for f in range(n):
    flow = sr.buffer("flow", h, w * 8, f).view(np.float32).reshape(h, w, 2)
    d = np.abs(flow - it["flow"][f])
    mask = sr.buffer("mask", h // 2, (w // 2) * 16, f).view(np.float32).reshape(h // 2, w // 2, 4)
    dm = np.abs(mask - it["mask"][f])
    r2 = (np.round(2 * flow) != np.round(2 * it["flow"][f])).any(-1)
    mk = it["mask"][f][..., :3].max(-1) > 0
    mk_full = np.repeat(np.repeat(mk, 2, 0), 2, 1)
    print(f"frame {f}: flow max {d.max():.3g} p99.9 {np.percentile(d, 99.9):.3g} frac>1e-2 {(d > 1e-2).mean():.3g} | round(2f) flips {r2.mean():.3g} (where mask>0: {(r2 & mk_full).mean():.3g}) | mask max {dm.max():.3g} frac>1e-3 {(dm > 1e-3).mean():.3g} | mask>0 frac {mk.mean():.3g}")

"""Debug: per merge tile (64x8 raw px), how far do integer HR shifts deviate from the tile mean?"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
n, h, w = 8, 3024, 4032
seeds = [int(a) for a in sys.argv[1:]] or [1235]
sr = BurstSuperResolution(default_params(), device=0, max_width=w, max_height=h, max_frames=n)
for seed in seeds:
    print('seed', seed)
    fr, sh = synth_burst(n, h, w, seed=seed, device='cuda')
    sr.set_input(fr); out = sr.next_frame(); sr.synchronize()
    for f in (1, 4, 7):
        fl = torch.from_numpy(sr.buffer('flow', h, w * 8, f).view(np.float32).reshape(h, w, 2)).cuda()
        s = torch.round(fl * 2)
        t = s[: h // 8 * 8, : w // 64 * 64].reshape(h // 8, 8, w // 64, 64, 2)
        mean = t.mean((1, 3), keepdim=True).round()
        d = (t - mean).abs()
        msg = []
        for thx, thy in ((8, 4), (12, 8), (20, 12), (28, 14), (60, 30)):
            bad = ((d[..., 0] > thx) | (d[..., 1] > thy)).float().mean().item()
            msg.append(f'>({thx},{thy}):{bad:.4f}')
        ts = sr.tile_shifts(f).reshape(-1, 2)
        med = np.median(ts, axis=0)
        print(f'frame {f}: true {(-(sh[f]-sh[0])).tolist()} tile-shift median {med.tolist()} tiles off by >1px: {(np.abs(ts-med).max(1)>1).mean():.3f}  pixel dev from merge-tile mean ' + ' '.join(msg))

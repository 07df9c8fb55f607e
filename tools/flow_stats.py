"""Debug: statistics of the pipeline's final flow on the bench burst (how noisy is round(2*flow)?)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
n, h, w = 8, 3024, 4032
p = default_params()
sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
fr, sh = synth_burst(n, h, w, seed=1234, device='cuda')
sr.set_input(fr); out = sr.next_frame(); sr.synchronize()
print('true shifts', sh)
for f in range(n):
    fl = torch.from_numpy(sr.buffer('flow', h, w * 8, f).view(np.float32).reshape(h, w, 2)).cuda()
    s = torch.round(fl * 2)
    med = s.reshape(-1, 2).median(0).values
    d = (s - med).abs()
    dx = (s[:, 1:] != s[:, :-1]).any(-1).float().mean().item()
    # tile-wise range (64x8 raw = 128x16 HR)
    t = s[: h // 8 * 8, : w // 64 * 64].reshape(h // 8, 8, w // 64, 64, 2)
    rng = (t.amax((1, 3)) - t.amin((1, 3)))
    print(f'frame {f}: median shift {med.tolist()}, |dev|>2: {(d > 2).any(-1).float().mean().item():.4f}, >8: {(d > 8).any(-1).float().mean().item():.5f}, max {d.max().item():.0f}, '
          f'adjacent-differ {dx:.3f}, tile range>8: {(rng > 8).any(-1).float().mean().item():.4f}, finite {torch.isfinite(fl).all().item()}')
    if f == 1:
        st = sr.tile_shifts(f)
        print('  tile shift std', st.reshape(-1, 2).std(0), 'flow std', fl.reshape(-1, 2).std(0).tolist())
        big = (d > 8).any(-1)
        ys, xs = torch.nonzero(big, as_tuple=True)
        if len(ys):
            print('  outliers rows range', ys.min().item(), ys.max().item(), 'cols', xs.min().item(), xs.max().item(), 'count', len(ys))
            print('  sample outlier coords', [(ys[i].item(), xs[i].item()) for i in range(0, len(ys), max(1, len(ys) // 10))][:10])

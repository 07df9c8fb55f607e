"""Debug: where do the garbage flows of some synthetic seeds come from?  Per frame: share of pixels whose final flow is more than
2 px from the truth, the same for the consolidated tile shifts (before flow-from-tiles / LK), and a coarse map of where they sit."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1235
dev = torch.device('cuda', 0)
n, h, w = 8, 3024, 4032
p = default_params()
sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
fr, sh = synth_burst(n, h, w, seed=seed, device=dev)
sr.set_input(fr); sr.next_frame(); torch.cuda.synchronize()
tx, ty, m = sr.tile_grid()
print('seed', seed, 'true shifts', [tuple(round(float(v), 2) for v in s) for s in sh])
for f in range(1, n):
    fl = sr.buffer('flow', h, w * 8, f).view(np.float32).reshape(h, w, 2)[::8, ::8]
    e = np.abs(fl + sh[f].numpy()[None, None, :]).max(-1)
    ts = sr.tile_shifts(f)
    et = np.abs(ts + sh[f].numpy()[None, None, :]).max(-1)
    bad = e > 2
    print(f'frame {f}: flow >2px {bad.mean():.3%}  tile shift >2px {(et > 2).mean():.3%}  |tile err| median {np.median(et):.2f}  max flow err {e.max():.1f}')
    if bad.mean() > 0.01:
        g = bad.reshape(6, bad.shape[0] // 6, 8, bad.shape[1] // 8).mean(axis=(1, 3))
        print('   bad share on a 6x8 grid:\n   ' + '\n   '.join(' '.join(f'{v:4.2f}' for v in row) for row in g))
        gt = (et > 2)
        gy, gx = gt.shape[0] // 6 * 6, gt.shape[1] // 8 * 8
        g2 = gt[:gy, :gx].reshape(6, gy // 6, 8, gx // 8).mean(axis=(1, 3))
        print('   bad TILE share on a 6x8 grid:\n   ' + '\n   '.join(' '.join(f'{v:4.2f}' for v in row) for row in g2))
for k in range(m):
    a = sr.tile_argmin(k)
    print('pair', k, 'argmin spread: |.|==max_shift share', float((np.abs(a) >= p.max_shift).any(-1).mean().round(3)), end='; ')
print()
st = sr.stage_ms(); print({k: round(v, 2) for k, v in st.items()})

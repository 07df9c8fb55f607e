"""Stage times of a row-band handle against the plain full-frame handle on the same burst (one GPU): what does band mode cost?"""
import sys
import torch
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200 import rowband
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst

dev = torch.device('cuda', 0)
n, h, w = 8, 3024, 4032
fr, _ = synth_burst(n, h, w, seed=1234, device=dev)
p = default_params()
for world in (0, 1, 2):
    if world == 0:
        sr = BurstSuperResolution(p, 0, w, h, n); banded = fr; name = 'full frame'
    else:
        bands = rowband.plan_bands(h, world, p.tile_size << (p.levels - 1), rowband.DEFAULT_HALO)
        b = bands[0]
        bp = rowband.band_params(p, b, h, rowband.default_margin(p))
        sr = BurstSuperResolution(bp, 0, w, b.bottom - b.top, n); banded = fr[:, b.top:b.bottom].contiguous(); name = f'band 0 of {world}'
    ow, oh = sr.output_size(w, banded.shape[1])
    out = torch.empty((oh, ow, 3), dtype=torch.float32, device=dev)
    for _ in range(3):
        sr.set_input(banded); sr.next_frame(out=out)
    torch.cuda.synchronize()
    st = sr.stage_ms()
    print(name, 'rows', banded.shape[1], 'out', (oh, ow), {k: round(v, 3) for k, v in st.items()}, 'sum', round(sum(st.values()), 2), flush=True)
    sr.close()

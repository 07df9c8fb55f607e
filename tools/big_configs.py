"""Smoke + timing of the larger BASELINE configs on one GPU (not bench lines): config 4 size (48 MP x 15, 2x) and a config-5-like
3x run (4K x 30 frames, generic merge kernel)."""
import sys, time, torch
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
dev = torch.device('cuda', 0)
for (name, n, h, w, scale, ms) in (('config4 48MP x15 2x', 15, 6048, 8064, 2, 3.0), ('config5 4K x30 3x', 30, 2160, 3840, 3, 48.0)):
    p = default_params(); p.scale = scale
    sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
    print(name, 'workspace GB', sr.workspace_bytes / 1e9, flush=True)
    fr, sh = synth_burst(n, h, w, seed=7, device=dev, max_shift=ms)
    ow, oh = sr.output_size(w, h)
    out = torch.empty((oh, ow, 3), dtype=torch.float32, device=dev)
    for _ in range(2):
        sr.set_input(fr); sr.next_frame(out=out)
    torch.cuda.synchronize()
    st = sr.stage_ms()
    med = [sr.tile_shifts(f).reshape(-1, 2).mean(0).tolist() for f in (1, n - 1)]
    print(name, {k: round(v, 2) for k, v in st.items()}, 'sum', round(sum(st.values()), 1), 'MP/s', round(ow * oh / 1e6 / (sum(st.values()) / 1e3)),
          'finite', bool(torch.isfinite(out).all()), 'mean', float(out.mean()), 'shift f1,fN', med, 'true', (-(sh[1]-sh[0])).tolist(), (-(sh[n-1]-sh[0])).tolist(), flush=True)
    sr.close(); del fr, out
    torch.cuda.empty_cache()

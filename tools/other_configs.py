import sys, torch
import numpy as np
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
dev = torch.device('cuda', 0)
for (name, n, h, w, scale, ms, fmt) in (('config5 4K x30 3x', 30, 2160, 3840, 3, 48.0, 0), ('config3 1080p gray x8 2x', 8, 1080, 1920, 2, 3.0, 1)):
    p = default_params(); p.scale = scale
    if ms > p.max_shift * (2 ** p.levels - 1) * 0.75:
        # reach of the matcher = max_shift * (2^levels - 1) px: 60 px at the default 4 levels, 124 px at 5; beyond that (or for
        # rotations) the global pre-alignment finds the frame pose first (+-160 px at this size)
        if '--levels5' in sys.argv: p.levels = 5
        else: p.prealign = 1; p.pair_span = 1 if n > 20 else p.pair_span        # pairs: neighbours + every frame against the reference (<= 64)
    sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
    fr, sh = synth_burst(n, h, w, seed=7, device=dev, max_shift=ms)
    ow, oh = sr.output_size(w, h)
    out = torch.empty((oh, ow, 3), dtype=torch.float32, device=dev)
    for _ in range(3):
        sr.set_input(fr, fmt=fmt); sr.next_frame(out=out)
    torch.cuda.synchronize()
    st = sr.stage_ms()
    # back-to-back bursts on the handle's stream (what BASELINE configs[2] asks for: 256 independent bursts), device events around the lot
    reps = 64 if n <= 8 else 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ext = torch.cuda.ExternalStream(sr.stream, device=dev)
    e0.record(ext)
    for _ in range(reps):
        sr.set_input(fr, fmt=fmt); sr.next_frame(out=out)
    e1.record(ext)
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / reps
    # final per-pixel flow (pose + tile shifts + LK) against the burst's true shifts, central rows, a few frames:
    # recovered motion = -truth (frame(x) = scene(x + d))
    errs = []
    for f in range(1, n, max(1, n // 6)):
        fl = sr.buffer('flow', h, w * 8, f).view(np.float32).reshape(h, w, 2)[h // 4: 3 * h // 4: 8, w // 4: 3 * w // 4: 8]
        errs.append(np.abs(fl + sh[f].cpu().numpy()[None, None, :]).max(axis=-1).ravel())
    err = np.concatenate(errs)
    print(name, 'levels', p.levels, 'prealign', p.prealign, 'flow error px: median', round(float(np.median(err)), 3), '90th pct', round(float(np.percentile(err, 90)), 3),
          '99th pct', round(float(np.percentile(err, 99)), 3), flush=True)
    print(name, {k: round(v, 2) for k, v in st.items()}, 'sum', round(sum(st.values()), 2), 'MP/s', round(ow * oh / 1e6 / (sum(st.values()) / 1e3)),
          'back-to-back ms/burst', round(per, 3), 'launches', sr.launch_count(), flush=True)
    if n <= 8:
        # the same bursts on several handles (one stream each) issued round-robin: kernels of different bursts overlap and fill the
        # tails of the small grids
        import time
        for nh in (2, 3):
            hs = [BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n) for _ in range(nh)]
            outs = [torch.empty((oh, ow, 3), dtype=torch.float32, device=dev) for _ in range(nh)]
            for k in range(2 * nh):
                hs[k % nh].set_input(fr, fmt=fmt); hs[k % nh].next_frame(out=outs[k % nh])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(reps):
                hs[k % nh].set_input(fr, fmt=fmt); hs[k % nh].next_frame(out=outs[k % nh])
            torch.cuda.synchronize()
            print(name, f'{nh} handles round-robin: ms/burst', round((time.perf_counter() - t0) * 1e3 / reps, 3), flush=True)
            for x in hs: x.close()
    sr.close(); del fr, out
if '--all' not in sys.argv:
    sys.exit(0)
# config 4 size on one GPU (48 MP x 15 frames): the frame-chunked merge path
for (name, n, h, w) in (('config4 48MP x15 2x', 15, 6048, 8064),):
    p = default_params()
    sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
    fr, sh = synth_burst(n, h, w, seed=4321, device=dev)
    ow, oh = sr.output_size(w, h)
    out = torch.empty((oh, ow, 3), dtype=torch.float32, device=dev)
    for _ in range(3):
        sr.set_input(fr); sr.next_frame(out=out)
    torch.cuda.synchronize()
    st = sr.stage_ms()
    print(name, {k: round(v, 2) for k, v in st.items()}, 'sum', round(sum(st.values()), 2), 'MP/s', round(ow * oh / 1e6 / (sum(st.values()) / 1e3)),
          'workspace GB', round(sr.workspace_bytes / 1e9, 2), flush=True)
    sr.close(); del fr, out

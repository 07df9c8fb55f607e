"""Sweep of the tile matcher's flat-surface threshold (findMinimum `threshold`, kernel.cu:519: a tile whose SSD surface spans less than
the threshold gets a zero shift) on synthetic seeds: step / merge time, pixels outside the merge window, flow error on textured areas."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst

dev = torch.device('cuda', 0)
n, h, w = 8, 3024, 4032
seeds = [int(a) for a in sys.argv[1:]] or [1234, 1235, 1237]
for seed in seeds:
    fr, sh = synth_burst(n, h, w, seed=seed, device=dev)
    for thr in (0.0, 64.0, 256.0, 1024.0, 4096.0):
        p = default_params(); p.min_threshold = thr
        sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
        for _ in range(2):
            sr.set_input(fr); sr.next_frame(); torch.cuda.synchronize()
        st = sr.stage_ms()
        errs, outs = [], []
        for f in (1, 4, 7):
            fl = sr.buffer('flow', h, w * 8, f).view(np.float32).reshape(h, w, 2)
            e = np.abs(fl[::8, ::8] + sh[f].numpy()[None, None, :]).max(-1)
            errs.append(e.ravel())
            s2 = np.round(2 * fl[: h // 8 * 8, : w // 64 * 64]).reshape(h // 8, 8, w // 64, 64, 2)
            mean = np.round(s2.mean(axis=(1, 3), keepdims=True))
            outs.append(((np.abs(s2[..., 0] - mean[..., 0]) > 24) | (np.abs(s2[..., 1] - mean[..., 1]) > 10)).mean())
        e = np.concatenate(errs)
        print(f'seed {seed} thr {thr:6.0f}: step {sum(st.values()):6.2f} ms merge {st["merge"]:5.2f} consolidate {st["consolidate"]:4.2f}  flow err median {np.median(e):.3f} '
              f'>1px {np.mean(e > 1):.3%} >8px {np.mean(e > 8):.3%}  outside merge window {np.mean(outs):.3%}', flush=True)
        sr.close()

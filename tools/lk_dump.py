"""Runs the config-2 pipeline, prints the flow stage time and writes a hash of the final flow field (for same-box A/B builds)."""
import hashlib, sys
import numpy as np, torch
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
dev = torch.device('cuda', 0)
n, h, w = 8, 3024, 4032
sr = BurstSuperResolution(default_params(), 0, w, h, n)
for seed in (1234, 1237):
    fr, _ = synth_burst(n, h, w, seed=seed, device=dev)
    ms = []
    for _ in range(4):
        sr.set_input(fr); out = sr.next_frame(); torch.cuda.synchronize(); ms.append(sr.stage_ms()['flow'])
    hs = hashlib.sha256()
    for f in range(n):
        hs.update(sr.buffer('flow', h, w * 8, f).tobytes())
    print('seed', seed, 'flow ms', [round(m, 3) for m in ms], 'flow sha', hs.hexdigest()[:16], 'image sha', hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16], flush=True)

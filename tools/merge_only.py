"""Runs the merge stage alone at config-2 size on the flows / masks / kernels of a real pipeline run (for ncu captures and timing).
    python tools/merge_only.py [reps]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
n, h, w = 8, 3024, 4032
p = default_params()
sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
fr, _ = synth_burst(n, h, w, seed=1234, device=dev)
ms = []
for _ in range(reps):
    sr.set_input(fr)
    sr.next_frame()
    torch.cuda.synchronize()
    ms.append(sr.stage_ms()["merge"])
print("merge ms", ms)
sr.close()

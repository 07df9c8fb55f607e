"""Cycles per phase of merge_s2_dyn_kernel (thread 0 of every CTA, i.e. warp 0's view), from a library built with
MFSR_NVCC_EXTRA=-DMFSR_MERGE_TIMING:   MFSR_NVCC_EXTRA=-DMFSR_MERGE_TIMING python -m multi_frame_super_resolution_b200.build --force
ticks: 0 prefetch + shift phase (warp 0's share), 1 certainty/kernel staging, 2 wait at barrier 1, 3 raw staging, 4 wait at barrier 2,
5 phase 2 of warp 0 (four passes), 6 wait for the slowest warp of the CTA."""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
from multi_frame_super_resolution_b200 import _lib
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
lib = C.CDLL(str(_lib.LIB_PATH)) if hasattr(_lib, "LIB_PATH") else _lib.load()
fn = lib.mfsr_debug_merge_phase_cycles
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
n, h, w = 8, 3024, 4032
p = default_params()
sr = BurstSuperResolution(p, 0, w, h, n)
fr, _ = synth_burst(n, h, w, seed=1234, device='cuda')
for _ in range(2):
    sr.set_input(fr); sr.next_frame()
torch.cuda.synchronize()
fn(None, 1)
sr.set_input(fr); sr.next_frame(); torch.cuda.synchronize()
buf = (C.c_ulonglong * 8)()
fn(buf, 0)
v = list(buf)[:7]; tot = sum(v)
names = ["prefetch+shifts", "stage mask/kernel", "barrier 1 wait", "stage raw", "barrier 2 wait", "phase 2 (warp 0)", "wait slowest warp"]
for nme, x in zip(names, v):
    print(f"{nme:20s} {100 * x / tot:5.1f} %")
print("merge ms", sr.stage_ms()["merge"])

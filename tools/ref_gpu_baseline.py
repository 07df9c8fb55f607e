#!/usr/bin/env python
"""The reference's OWN CUDA kernels (test_opencv/*.cu compiled unmodified into oracle/_ref/libmfsr_ref.so, launched by the
restated host oracle/ref_driver.cu) timed on the same B200 as the product, stage by stage, at BASELINE configs[1] size
(4032 x 3024 RGGB x 8 frames, 2x).

    gpurun -- 'python tools/ref_gpu_baseline.py > gpurun_out/ref_gpu_baseline.json'

Per stage one invocation is timed (CUDA events around the kernel launches only: textures, plans and buffers are set up outside)
and multiplied by the number of invocations the product's schedule makes per burst (pairs x levels, frames x LK sweeps ...),
which is the schedule the absent upstream host would have to run too.  The merge is timed as the whole reference chain
(8 x accumulateImagesSuperRes read-modify-write passes + ApplyWeighting + GammasRGB) on the reference's geometry (central crop,
output dims == raw dims) and compared with mfsr_stage_merge on the same buffers and geometry.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from multi_frame_super_resolution_b200 import stages  # noqa: E402
from multi_frame_super_resolution_b200._lib import MergeGeom  # noqa: E402
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params, measured_pairs  # noqa: E402
from multi_frame_super_resolution_b200.synth import synth_burst, synth_merge_inputs  # noqa: E402
from oracle import pyref  # noqa: E402

WHITE, BLACK = [959.0, 959.0, 959.0], [64.0, 64.0, 64.0]
SCALE = [float(np.float32(1.0) / np.float32(959.0))] * 3
RGGB = [0, 1, 1, 2]


def best_of(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        fn()
        best = min(best, pyref.last_kernel_ms())
    return best


def ours_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = torch.device("cuda:0")
    n, h, w = 8, 3024, 4032
    p = default_params()
    pairs = measured_pairs(n, p.pair_span)
    T, M, L = p.tile_size, p.max_shift, p.levels
    res = {"config": f"{w}x{h} RGGB x {n} frames, 2x", "gpu": torch.cuda.get_device_name(0),
           "schedule": {"pairs": len(pairs), "levels": L, "lk_sweeps": (n - 1) * p.lk_iterations}}

    # ---- the product's own chain on this box (stage events of mfsr_run)
    fr, _ = synth_burst(n, h, w, seed=1234, device=dev)
    sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
    for _ in range(3):
        sr.set_input(fr)
        sr.next_frame()
    torch.cuda.synchronize()
    res["ours_stage_ms"] = {k: round(v, 3) for k, v in sr.stage_ms().items()}
    sr.close()

    ref = {}
    d = fr[0]
    # ---- front end: subsample3 per frame, demosaic (two kernels) per frame for the tracking image
    ref["subsample3_per_frame"] = best_of(lambda: pyref.subsample3(d, 1023.0, RGGB))
    ref["debayer_per_frame"] = best_of(lambda: pyref.debayer(d, BLACK, SCALE, RGGB))
    # ---- tile matcher chain, cuFFT cross-correlation as upstream, one pair per level
    gray, gq = stages.tracking_image(d, BLACK, SCALE, p.track_sigma, p.track_bits, RGGB)
    gray1, gq1 = stages.tracking_image(fr[1], BLACK, SCALE, p.track_sigma, p.track_bits, RGGB)
    lv_a, lv_b = [gq], [gq1]
    for _ in range(1, L):
        lv_a.append(stages.pyramid_down(lv_a[-1]))
        lv_b.append(stages.pyramid_down(lv_b[-1]))
    per_level = []
    for l in range(L):
        a, b = lv_a[l].float().contiguous(), lv_b[l].float().contiguous()
        per_level.append(best_of(lambda: pyref.tile_align(a, b, None, T, M, use_fft=True), reps=2))
        del a, b
        torch.cuda.empty_cache()
    ref["tile_align_per_pair_by_level"] = per_level
    # ---- flow
    ty, tx = (h - 2 * M) // T, (w - 2 * M) // T
    tiles = (torch.rand((ty, tx, 2), device=dev) * 4 - 2).contiguous()
    ref["flow_from_tiles_per_frame"] = best_of(lambda: pyref.flow_from_tiles(tiles, T, w, h))
    flow = pyref.flow_from_tiles(tiles, T, w, h)
    t_warp = best_of(lambda: pyref.warp(flow, gray1))
    warped = pyref.warp(flow, gray1)
    t_der = best_of(lambda: pyref.derivatives(warped, gray))
    ix, iy, iz = pyref.derivatives(warped, gray)
    t_lk = best_of(lambda: pyref.lucas_kanade(flow, ix, iy, iz, p.lk_half_window, p.lk_min_det))
    ref["lk_sweep"] = {"warp": t_warp, "derivatives": t_der, "lucas_kanade": t_lk, "total": t_warp + t_der + t_lk}
    del warped, ix, iy, iz
    # ---- kernel parameters (the box smoothing between tensor and eigen-analysis belongs to NPP in upstream: not included)
    t_d2 = best_of(lambda: pyref.derivatives2(gray))
    ix2, iy2 = pyref.derivatives2(gray)
    t_st = best_of(lambda: pyref.structure_tensor(ix2, iy2))
    t3 = pyref.structure_tensor(ix2, iy2)
    t_kp = best_of(lambda: pyref.kernel_param(t3, p.Dth, p.Dtr, p.kDetail, p.kDenoise, p.kStretch, p.kShrink))
    ref["kernel_params"] = {"derivatives2": t_d2, "structure_tensor": t_st, "kernel_param": t_kp, "total": t_d2 + t_st + t_kp}
    del ix2, iy2, t3
    # ---- robustness
    a3, b3 = pyref.subsample3(d, 1023.0, RGGB), pyref.subsample3(fr[1], 1023.0, RGGB)
    ref["robustness_per_frame"] = best_of(lambda: pyref.robustness_mask(a3, b3, flow, p.alpha, p.beta, p.thresholdM))
    del a3, b3, flow, gray, gray1, fr
    torch.cuda.empty_cache()

    # ---- merge: reference chain vs mfsr_stage_merge, same buffers, reference geometry
    raw, mask, mflow, kern = synth_merge_inputs(n, h, w, seed=1234, device=dev)
    fb = torch.rand((h, w, 3), device=dev)
    ms_ref = min(pyref.merge_chain_ms(raw, mask, mflow, kern, fb, WHITE, BLACK, 0.1, RGGB, gamma=True)[0] for _ in range(3))
    geom = MergeGeom.reference(w, h)
    ms_ours = ours_ms(lambda: stages.merge(raw, mask, mflow, kern, fb, geom, WHITE, BLACK, 0.1, flags=1))
    gfull = MergeGeom.full_frame(w, h, 2)
    fbf = torch.rand((2 * h, 2 * w, 3), device=dev)
    ms_ours_full = ours_ms(lambda: stages.merge(raw, mask, mflow, kern, fbf, gfull, WHITE, BLACK, 0.1, flags=1))
    ref["merge_chain_reference_geometry"] = ms_ref
    res["merge"] = {"ref_ms_12mp_out": round(ms_ref, 3), "ours_ms_12mp_out": round(ms_ours, 3), "speedup_same_geometry": round(ms_ref / ms_ours, 2),
                    "ours_ms_full_frame_48mp_out": round(ms_ours_full, 3),
                    "ref_ms_per_out_mp": round(ms_ref / (w * h / 1e6), 4), "ours_ms_per_out_mp_full_frame": round(ms_ours_full / (4 * w * h / 1e6), 4),
                    "note": "ours includes the output allocation of stages.merge (torch.empty) inside the timed call"}
    res["ref_kernel_ms"] = ref

    # ---- whole-burst estimate for the reference kernels under the product's schedule (central-crop merge: 1/4 of the output pixels)
    align = sum(per_level) * len(pairs)
    est = {"frontend": n * (ref["subsample3_per_frame"] + ref["debayer_per_frame"]), "align": align,
           "flow": (n - 1) * ref["flow_from_tiles_per_frame"] + (n - 1) * p.lk_iterations * ref["lk_sweep"]["total"],
           "kernel_params": ref["kernel_params"]["total"], "robustness": n * ref["robustness_per_frame"], "merge_12mp_out": ms_ref}
    est["sum"] = sum(est.values())
    res["ref_burst_estimate_ms"] = {k: round(v, 3) for k, v in est.items()}
    res["ours_burst_ms"] = round(sum(res["ours_stage_ms"].values()), 3)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

"""Thin torch-tensor wrappers over the stage entry points of include/mfsr.h.

PyTorch is used only for device memory and streams; every function below makes
exactly one call into libmfsr_b200.so.  Layouts (all CUDA, contiguous):
  raw u16 -> torch.int16/uint16 [..., H, W];  float3 -> float32 [H, W, 3];
  float4 -> float32 [H, W, 4];  float2 -> float32 [H, W, 2].
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import MergeGeom, check, farr, iarr

RGGB = (0, 1, 1, 2)   # c_cfaPattern[2][2] row-major, BayerColor values (DeBayerKernels.cu:28-41)
GRAY = (1, 1, 1, 1)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError("stage wrappers need contiguous CUDA tensors")


def subsample3(raw, max_val, cfa=RGGB):
    _cuda(raw)
    h2, w2 = raw.shape[-2] // 2, raw.shape[-1] // 2
    out = torch.empty((h2, w2, 3), dtype=torch.float32, device=raw.device)
    check(_lib.load().mfsr_stage_subsample3(_p(raw), raw.shape[-1] * 2, _p(out), w2 * 12, float(max_val), w2, h2, iarr(cfa), _stream()),
          "mfsr_stage_subsample3")
    return out


def demosaic(raw, black, scale, cfa=RGGB):
    _cuda(raw)
    h, w = raw.shape[-2:]
    out = torch.empty((h, w, 3), dtype=torch.float32, device=raw.device)
    check(_lib.load().mfsr_stage_demosaic(_p(raw), w * 2, _p(out), w * 12, w, h, iarr(cfa), farr(black), farr(scale), _stream()),
          "mfsr_stage_demosaic")
    return out


def tracking_image(raw, black, scale, sigma=0.5, track_bits=7, cfa=RGGB):
    _cuda(raw)
    h, w = raw.shape[-2:]
    gray = torch.empty((h, w), dtype=torch.float32, device=raw.device)
    gq = torch.empty((h, w), dtype=torch.uint8, device=raw.device)
    check(_lib.load().mfsr_stage_tracking_image(_p(raw), w * 2, _p(gray), w * 4, _p(gq), w, w, h, iarr(cfa), farr(black), farr(scale),
                                                float(sigma), int(track_bits), _stream()), "mfsr_stage_tracking_image")
    return gray, gq


def pyramid_down(img):
    _cuda(img)
    h, w = img.shape
    out = torch.empty((h // 2, w // 2), dtype=torch.uint8, device=img.device)
    check(_lib.load().mfsr_stage_pyramid_down(_p(img), w, w, h, _p(out), w // 2, _stream()), "mfsr_stage_pyramid_down")
    return out


def tile_counts(w, h, tile_size, max_shift):
    return (w - 2 * max_shift) // tile_size, (h - 2 * max_shift) // tile_size


def tile_align(ref, mov, pre_shift=None, tile_size=16, max_shift=4, base_shift=(0.0, 0.0), base_rotation=0.0,
               threshold=0.0, want_ssd=False):
    _cuda(ref, mov, pre_shift)
    h, w = ref.shape
    tx, ty = tile_counts(w, h, tile_size, max_shift)
    S = 2 * max_shift + 1
    out = torch.empty((ty, tx, 2), dtype=torch.float32, device=ref.device)
    arg = torch.empty((ty, tx, 2), dtype=torch.int32, device=ref.device)
    ssd = torch.empty((ty * tx, S * S), dtype=torch.float32, device=ref.device) if want_ssd else None
    check(_lib.load().mfsr_stage_tile_align(_p(ref), _p(mov), w, w, h, _p(pre_shift), tx * 8, _p(out), tx * 8, _p(arg), _p(ssd),
                                            tile_size, max_shift, tx, ty, float(base_shift[0]), float(base_shift[1]),
                                            float(base_rotation), float(threshold), _stream()), "mfsr_stage_tile_align")
    return out, arg, ssd


def prealign_table():
    """(cos, sin) of the candidate angles of the global pre-alignment: 0.125 degree steps over +-21 degrees, computed in double and
    rounded once (pipeline.cu uploads the same table).  Returns (float32 [337, 2], index of angle 0)."""
    import numpy as np
    i = np.arange(2 * 21 * 8 + 1, dtype=np.float64) - 21 * 8
    th = i * (0.125 * 3.14159265358979323846 / 180.0)
    return np.stack([np.cos(th), np.sin(th)], axis=1).astype(np.float32), 21 * 8


def prealign_search(ref, mov, cs_table, idx0, step, n_ang, cx=0, cy=0, radius=8, sub=1):
    """One stage of the global pre-alignment search (csrc/prealign.cu) on a pair of 8-bit tracking images.  Returns int32 [3] =
    (angle candidate or -1, bx, by)."""
    _cuda(ref, mov, cs_table)
    h, w = ref.shape
    out = torch.zeros((3,), dtype=torch.int32, device=ref.device)
    check(_lib.load().mfsr_stage_prealign_search(_p(ref), _p(mov), ref.stride(0), w, h, _p(cs_table), cs_table.shape[0], idx0, step, n_ang,
                                                 cx, cy, radius, sub, _p(out), _stream()), "mfsr_stage_prealign_search")
    return out


def upsample_shifts(in_shift, old_level, new_level, new_cx, new_cy, old_t, new_t):
    _cuda(in_shift)
    ocy, ocx = in_shift.shape[:2]
    out = torch.empty((new_cy, new_cx, 2), dtype=torch.float32, device=in_shift.device)
    check(_lib.load().mfsr_stage_upsample_shifts(_p(in_shift), ocx * 8, _p(out), new_cx * 8, old_level, new_level, ocx, ocy,
                                                 new_cx, new_cy, old_t, new_t, _stream()), "mfsr_stage_upsample_shifts")
    return out


def consolidate_shifts(measured, pair_from, pair_to, image_count, tiles_x, tiles_y, reference_image):
    """measured: float32 [tiles, m, 2] (concatenateShifts layout)."""
    _cuda(measured)
    nt, m = measured.shape[:2]
    n1 = image_count - 1
    dev = measured.device
    one = torch.empty((nt, n1, 2), dtype=torch.float32, device=dev)
    fs = torch.empty((image_count, tiles_y, tiles_x, 2), dtype=torch.float32, device=dev)
    status = torch.empty((nt,), dtype=torch.int32, device=dev)
    check(_lib.load().mfsr_stage_consolidate_shifts(_p(measured), iarr(pair_from), iarr(pair_to), m, image_count, tiles_x, tiles_y,
                                                    reference_image, _p(one), _p(fs), _p(status), _stream()),
          "mfsr_stage_consolidate_shifts")
    return one, fs, status


def flow_from_tiles(tile_shift, tile_size, w, h, base_shift=(0.0, 0.0), base_rotation=0.0):
    _cuda(tile_shift)
    ty, tx = tile_shift.shape[:2]
    flow = torch.empty((h, w, 2), dtype=torch.float32, device=tile_shift.device)
    check(_lib.load().mfsr_stage_flow_from_tiles(_p(tile_shift), tx * 8, tx, ty, tile_size, _p(flow), w * 8, w, h,
                                                 float(base_shift[0]), float(base_shift[1]), float(base_rotation), _stream()),
          "mfsr_stage_flow_from_tiles")
    return flow


def lk_iteration(ref, mov, flow, half_window=3, min_det=1e-3, texture=False):
    """One Lucas-Kanade sweep.  texture=True: the warp samples `mov` through the texture unit (mfsr_stage_lk_iteration_tex, what
    mfsr_run does outside row-band mode); False: the ALU model of the texture filter (bit-identical to the oracle)."""
    _cuda(ref, mov, flow)
    h, w = ref.shape
    out = torch.empty_like(flow)
    if texture:
        wp = (w + 7) // 8 * 8                       # texture pitch alignment: 32 bytes
        refp = torch.zeros((h, wp), dtype=torch.float32, device=ref.device); refp[:, :w] = ref
        movp = torch.zeros((h, wp), dtype=torch.float32, device=ref.device); movp[:, :w] = mov
        check(_lib.load().mfsr_stage_lk_iteration_tex(_p(refp), _p(movp), wp * 4, _p(flow), _p(out), w * 8, w, h, half_window, float(min_det),
                                                      _stream()), "mfsr_stage_lk_iteration_tex")
        return out
    check(_lib.load().mfsr_stage_lk_iteration(_p(ref), _p(mov), w * 4, _p(flow), _p(out), w * 8, w, h, half_window, float(min_det),
                                              _stream()), "mfsr_stage_lk_iteration")
    return out


def kernel_params(gray, box_radius=2, Dth=0.005, Dtr=0.012, kDetail=0.3, kDenoise=4.0, kStretch=4.0, kShrink=2.0):
    _cuda(gray)
    h, w = gray.shape
    out = torch.empty((h, w, 4), dtype=torch.float32, device=gray.device)
    check(_lib.load().mfsr_stage_kernel_params(_p(gray), w * 4, _p(out), w * 16, w, h, box_radius, Dth, Dtr, kDetail, kDenoise,
                                               kStretch, kShrink, _stream()), "mfsr_stage_kernel_params")
    return out


def robustness(rgb_ref, rgb_mov, flow, alpha, beta, threshold_m, erode_radius=0, fused=True):
    """rgb_*: [h, w, 3] half-res; flow: [2h, 2w, 2].  fused=False: certainty kernel + min-filter kernel through a scratch image;
    fused=True (what mfsr_run does): one kernel, same mask."""
    _cuda(rgb_ref, rgb_mov, flow)
    h, w = rgb_ref.shape[:2]
    mask = torch.empty((h, w, 4), dtype=torch.float32, device=rgb_ref.device)
    scratch = torch.empty_like(mask) if (erode_radius > 0 and not fused) else None
    check(_lib.load().mfsr_stage_robustness(_p(rgb_ref), _p(rgb_mov), w * 12, _p(flow), flow.shape[1] * 8, _p(mask), w * 16,
                                            _p(scratch), w, h, float(alpha), float(beta), float(threshold_m), erode_radius, _stream()),
          "mfsr_stage_robustness")
    return mask


def fallback_upsample(rgb, geom: MergeGeom):
    _cuda(rgb)
    h, w = rgb.shape[:2]
    out = torch.empty((geom.out_h, geom.out_w, 3), dtype=torch.float32, device=rgb.device)
    check(_lib.load().mfsr_stage_fallback_upsample(_p(rgb), w * 12, w, h, _p(out), geom.out_w * 12, C.byref(geom), _stream()),
          "mfsr_stage_fallback_upsample")
    return out


def merge(raw, mask, flow, kernel4, fallback, geom: MergeGeom, white, black, threshold, cfa=RGGB, flags=0,
          want_accumulators=False, out=None):
    """raw [N,H,W] u16; mask [N,H/2,W/2,4]; flow [N,H,W,2]; kernel4 [H,W,4]; fallback [OH,OW,3] or None."""
    _cuda(raw, mask, flow, kernel4, fallback)
    n, h, w = raw.shape
    dev = raw.device
    if out is None:
        out = torch.empty((geom.out_h, geom.out_w, 3), dtype=torch.float32, device=dev)
    s = wt = None
    if want_accumulators:
        s = torch.empty_like(out)
        wt = torch.empty_like(out)
    if fallback is None:
        flags |= 2
    check(_lib.load().mfsr_stage_merge(_p(raw), w * 2, w * h * 2, _p(mask), (w // 2) * 16, (w // 2) * (h // 2) * 16,
                                       _p(flow), w * 8, w * h * 8, _p(kernel4), w * 16, _p(fallback), geom.out_w * 12,
                                       _p(out), geom.out_w * 12, _p(s), _p(wt), geom.out_w * 12, n, C.byref(geom), iarr(cfa),
                                       farr(white), farr(black), float(threshold), int(flags), _stream()), "mfsr_stage_merge")
    return (out, s, wt) if want_accumulators else out

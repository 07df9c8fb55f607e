"""torch bindings for oracle/_ref/libmfsr_ref.so — the reference's OWN kernels
(compiled unmodified from /root/reference/test_opencv/*.cu by oracle/Makefile) behind the
restated host driver oracle/ref_driver.cu.

TEST INFRASTRUCTURE ONLY.  Needs a GPU; `available()` is False when the prebuilt library
is absent (it is built in the authoring container, where /root/reference exists, and
travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "libmfsr_ref.so"
vp, c_i, c_f = C.c_void_p, C.c_int, C.c_float
_lib = None


def available() -> bool:
    return LIB.exists()


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(LIB))
    return _lib


def _p(t):
    return vp(t.data_ptr()) if t is not None else None


def _i(v):
    return (c_i * len(v))(*[int(x) for x in v])


def _f(v):
    return (c_f * len(v))(*[float(x) for x in v])


def _ck(rc, what):
    if rc != 0:
        raise RuntimeError(f"ref_driver {what} failed: {rc}")


def subsample3(raw, max_val, cfa):
    h2, w2 = raw.shape[0] // 2, raw.shape[1] // 2
    out = torch.zeros((h2, w2, 3), dtype=torch.float32, device=raw.device)
    _ck(lib().ref_subsample3(_p(raw), _p(out), c_f(max_val), w2, h2, _i(cfa)), "subsample3")
    return out


def debayer(raw_u16, black, scale, cfa):
    rawf = raw_u16.to(torch.int32).bitwise_and(0xFFFF).to(torch.float32).contiguous()
    h, w = rawf.shape
    out = torch.zeros((h, w, 3), dtype=torch.float32, device=rawf.device)
    _ck(lib().ref_debayer(_p(rawf), _p(out), w, h, _i(cfa), _f(black), _f(scale)), "debayer")
    return out


def tile_align(ref_f, mov_f, pre=None, T=16, M=4, base_shift=(0.0, 0.0), rot=0.0, threshold=0.0, use_fft=False):
    h, w = ref_f.shape
    tx, ty = (w - 2 * M) // T, (h - 2 * M) // T
    S = 2 * M + 1
    coord = torch.zeros((ty, tx, 2), dtype=torch.float32, device=ref_f.device)
    ssd = torch.zeros((ty * tx, S * S), dtype=torch.float32, device=ref_f.device)
    _ck(lib().ref_tile_align(_p(ref_f), _p(mov_f), w, h, _p(pre), _p(coord), _p(ssd), T, M, tx, ty,
                             c_f(base_shift[0]), c_f(base_shift[1]), c_f(rot), c_f(threshold), int(use_fft)), "tile_align")
    return coord, ssd


def upsample_shifts(in2, old_level, new_level, new_cx, new_cy, old_t, new_t):
    ocy, ocx = in2.shape[:2]
    out = torch.zeros((new_cy, new_cx, 2), dtype=torch.float32, device=in2.device)
    _ck(lib().ref_upsample_shifts(_p(in2), _p(out), old_level, new_level, ocx, ocy, new_cx, new_cy, old_t, new_t), "upsample")
    return out


def flow_from_tiles(tile2, T, w, h, base_shift=(0.0, 0.0), rot=0.0):
    ty, tx = tile2.shape[:2]
    flow = torch.zeros((h, w, 2), dtype=torch.float32, device=tile2.device)
    _ck(lib().ref_flow_from_tiles(_p(tile2), tx, ty, T, _p(flow), w, h, c_f(base_shift[0]), c_f(base_shift[1]), c_f(rot)), "flow_from_tiles")
    return flow


def warp(flow, img):
    h, w = img.shape
    out = torch.zeros_like(img)
    _ck(lib().ref_warp(_p(flow), _p(img), _p(out), w, h), "warp")
    return out


def derivatives(src, tgt):
    h, w = src.shape
    ix, iy, iz = torch.zeros_like(src), torch.zeros_like(src), torch.zeros_like(src)
    _ck(lib().ref_derivatives(_p(src), _p(tgt), _p(ix), _p(iy), _p(iz), w, h), "derivatives")
    return ix, iy, iz


def derivatives2(img):
    h, w = img.shape
    ix, iy = torch.zeros_like(img), torch.zeros_like(img)
    _ck(lib().ref_derivatives2(_p(img), _p(ix), _p(iy), w, h), "derivatives2")
    return ix, iy


def lucas_kanade(flow, ix, iy, it, half_window, min_det):
    out = flow.clone()
    h, w = ix.shape
    _ck(lib().ref_lucas_kanade(_p(out), _p(ix), _p(iy), _p(it), w, h, half_window, c_f(min_det)), "lucas_kanade")
    return out


def lk_iteration(ref, mov, flow, half_window=3, min_det=1e-3):
    warped = warp(flow, mov)
    ix, iy, iz = derivatives(warped, ref)   # texSource = warped, texTarget = reference (see mfsr_oracle.c)
    return lucas_kanade(flow, ix, iy, iz, half_window, min_det)


def structure_tensor(ix, iy):
    h, w = ix.shape
    t = torch.zeros((h, w, 3), dtype=torch.float32, device=ix.device)
    _ck(lib().ref_structure_tensor(_p(ix), _p(iy), _p(t), w, h), "structure_tensor")
    return t


def kernel_param(t3, Dth, Dtr, kDetail, kDenoise, kStretch, kShrink):
    k = t3.clone()
    h, w = k.shape[:2]
    _ck(lib().ref_kernel_param(_p(k), w, h, c_f(Dth), c_f(Dtr), c_f(kDetail), c_f(kDenoise), c_f(kStretch), c_f(kShrink)), "kernel_param")
    return k


def robustness_mask(ref3, mov3, flow, alpha, beta, threshold_m):
    h, w = ref3.shape[:2]
    fh, fw = flow.shape[:2]
    mask = torch.zeros((h, w, 4), dtype=torch.float32, device=ref3.device)
    _ck(lib().ref_robustness_mask(_p(ref3), _p(mov3), _p(mask), _p(flow), fw, fh, w, h, c_f(alpha), c_f(beta), c_f(threshold_m)), "robustness")
    return mask


def merge_superres(raw, mask, flow, kernel4, fallback, white, black, threshold, cfa, gamma=False, want_accumulators=False):
    """N x accumulateImagesSuperRes -> ApplyWeighting -> GammasRGB on the reference geometry (out dims == raw dims)."""
    n, h, w = raw.shape
    s = torch.zeros((h, w, 3), dtype=torch.float32, device=raw.device)
    wt = torch.zeros_like(s)
    for f in range(n):
        _ck(lib().ref_accumulate_superres(_p(raw[f]), _p(s), _p(wt), _p(mask[f]), _p(kernel4), _p(flow[f]), w, h, _i(cfa), _f(white), _f(black)),
            "accumulate_superres")
    out = fallback.clone() if fallback is not None else torch.zeros_like(s)
    _ck(lib().ref_apply_weighting(_p(out), _p(s), _p(wt), w, h, c_f(threshold)), "apply_weighting")
    if gamma:
        _ck(lib().ref_gamma(_p(out), w, h), "gamma")
    return (out, s, wt) if want_accumulators else out


def merge_1x(raw, mask, flow, kernel3, fallback, white, black, threshold, cfa):
    n, h, w = raw.shape
    s = torch.zeros((h, w, 3), dtype=torch.float32, device=raw.device)
    wt = torch.zeros_like(s)
    for f in range(n):
        _ck(lib().ref_accumulate_1x(_p(raw[f]), _p(s), _p(wt), _p(mask[f]), _p(kernel3), _p(flow[f]), w, h, _i(cfa), _f(white), _f(black)),
            "accumulate_1x")
    out = fallback.clone() if fallback is not None else torch.zeros_like(s)
    _ck(lib().ref_apply_weighting(_p(out), _p(s), _p(wt), w, h, c_f(threshold)), "apply_weighting")
    return out


def merge_chain_ms(raw, mask, flow, kernel4, fallback, white, black, threshold, cfa, gamma=False):
    """Device time (ms) of the reference's own merge chain: N RMW passes + ApplyWeighting (+ gamma)."""
    n, h, w = raw.shape
    s = torch.zeros((h, w, 3), dtype=torch.float32, device=raw.device)
    wt = torch.zeros_like(s)
    io = fallback.clone()
    ms = c_f(0)
    _ck(lib().ref_merge_chain_timed(_p(raw), _p(mask), _p(kernel4), _p(flow), _p(s), _p(wt), _p(io), n, w, h, _i(cfa), _f(white), _f(black),
                                    c_f(threshold), int(gamma), C.byref(ms)), "merge_chain_timed")
    return ms.value, io


def texture_probe(texels, xn):
    out = torch.zeros_like(xn)
    _ck(lib().ref_texture_probe(_p(texels), texels.numel(), _p(xn), _p(out), xn.numel()), "texture_probe")
    return out


def last_kernel_ms() -> float:
    """Device time of the kernels of the last ref_* call (events around the launches, set-up excluded)."""
    f = lib().ref_last_kernel_ms
    f.restype = c_f
    return float(f())


def consolidate(measured, pair_from, pair_to, image_count, tx, ty, reference_image):
    """Reference shift consolidation: copyShiftMatrix / setPointers / transposeShifts / checkForOutliers / getOptimalShifts
    (ShiftMinimizerKernels.cu) around cuBLAS batched normal equations (the restated host, oracle/ref_driver.cu).
    measured: float32 [tiles, m, 2] on the device.  Returns (one_to_one [tiles, n-1, 2], frame_shift [n, ty, tx, 2],
    status [tiles] (reference: -1 when done), removed [tiles] (measurements removed per tile))."""
    nt, m = measured.shape[:2]
    assert nt == tx * ty
    dev = measured.device
    o2o = torch.zeros((nt, image_count - 1, 2), dtype=torch.float32, device=dev)
    fs = torch.zeros((image_count, ty, tx, 2), dtype=torch.float32, device=dev)
    st = torch.zeros((nt,), dtype=torch.int32, device=dev)
    rm = torch.zeros((nt,), dtype=torch.int32, device=dev)
    _ck(lib().ref_consolidate(_p(measured.contiguous()), _i(pair_from), _i(pair_to), m, image_count, tx, ty, reference_image,
                              _p(o2o), _p(fs), _p(st), _p(rm)), "consolidate")
    return o2o, fs, st, rm

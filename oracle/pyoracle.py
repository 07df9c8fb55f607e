"""numpy bindings for oracle/libmfsr_oracle.so — the CPU restatement of the reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Never imported by multi_frame_super_resolution_b200.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libmfsr_oracle.so"

c_i, c_f, vp = C.c_int, C.c_float, C.c_void_p
ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)


class Geom(C.Structure):
    _fields_ = [(n, c_i) for n in ("raw_w", "raw_h", "scale", "out_w", "out_h", "org_x", "org_y",
                                   "clamp_x0", "clamp_x1", "clamp_y0", "clamp_y1")]

    @classmethod
    def reference(cls, w, h):
        return cls(w, h, 2, w, h, w // 2, h // 2, w // 4, w // 2 - 1 + w // 4, h // 4, h // 2 - 1 + h // 4)

    @classmethod
    def full_frame(cls, w, h, s):
        num, den = s & 0xffff, ((s >> 16) & 0xffff) or 1          # include/mfsr.h MFSR_SCALE_RATIONAL
        return cls(w, h, s, w * num // den, h * num // den, 0, 0, 0, w - 1, 0, h - 1)

    @classmethod
    def from_product(cls, g):
        return cls(*[getattr(g, n) for n, _ in cls._fields_])


def build(force: bool = False) -> Path:
    src = HERE / "mfsr_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
                               "-o", str(LIB), str(src), "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_set_threads.restype = c_i
        _lib.orc_gauss_taps.restype = c_i
    return _lib


def _a(x, dt):
    return np.ascontiguousarray(x, dtype=dt)


def _p(x):
    return x.ctypes.data_as(vp) if x is not None else None


def _i(v):
    return (c_i * len(v))(*[int(x) for x in v])


def _f(v):
    return (c_f * len(v))(*[float(x) for x in v])


def set_threads(n):
    return lib().orc_set_threads(int(n))


def subsample3(raw, max_val, cfa):
    raw = _a(raw, np.uint16)
    h2, w2 = raw.shape[0] // 2, raw.shape[1] // 2
    out = np.zeros((h2, w2, 3), np.float32)
    lib().orc_subsample3(_p(raw), _p(out), c_f(max_val), w2, h2, _i(cfa))
    return out


def demosaic(raw, black, scale, cfa):
    rawf = _a(raw, np.float32)
    h, w = rawf.shape
    out = np.zeros((h, w, 3), np.float32)
    lib().orc_debayer_green(_p(rawf), _p(out), w, h, _i(cfa), _f(black), _f(scale))
    lib().orc_debayer_redblue(_p(rawf), _p(out), w, h, _i(cfa), _f(black), _f(scale))
    return out


def gauss_taps(sigma):
    t = np.zeros(99, np.float32)
    n = lib().orc_gauss_taps(c_f(sigma), _p(t))
    return t[:n].copy()


def tracking_image(rgb, sigma, track_bits):
    rgb = _a(rgb, np.float32)
    h, w = rgb.shape[:2]
    gray = np.zeros((h, w), np.float32)
    gq = np.zeros((h, w), np.uint8)
    lib().orc_tracking_image(_p(rgb), _p(gray), _p(gq), w, h, c_f(sigma), int(track_bits))
    return gray, gq


def pyramid_down(img):
    img = _a(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h // 2, w // 2), np.uint8)
    lib().orc_pyramid_down(_p(img), w, h, _p(out))
    return out


def tile_align(ref, mov, pre=None, T=16, M=4, base_shift=(0.0, 0.0), rot=0.0, threshold=0.0):
    ref, mov = _a(ref, np.uint8), _a(mov, np.uint8)
    h, w = ref.shape
    tx, ty = (w - 2 * M) // T, (h - 2 * M) // T
    S = 2 * M + 1
    pre = _a(pre, np.float32) if pre is not None else None
    out = np.zeros((ty, tx, 2), np.float32)
    arg = np.zeros((ty, tx, 2), np.int32)
    ssd = np.zeros((ty * tx, S * S), np.float32)
    lib().orc_tile_align(_p(ref), _p(mov), w, h, _p(pre), _p(out), _p(arg), _p(ssd), T, M, tx, ty,
                         c_f(base_shift[0]), c_f(base_shift[1]), c_f(rot), c_f(threshold))
    return out, arg, ssd


def cross_correlation(a_tiles, b_tiles):
    a, b = _a(a_tiles, np.float32), _a(b_tiles, np.float32)
    n, P, _ = a.shape
    cc = np.zeros_like(a)
    lib().orc_cross_correlation(_p(a), _p(b), _p(cc), P, n)
    return cc


def find_minimum(ssd, M, threshold=0.0):
    ssd = _a(ssd, np.float32)
    n = ssd.shape[0]
    coord = np.zeros((n, 2), np.float32)
    arg = np.zeros((n, 2), np.int32)
    lib().orc_find_minimum(_p(ssd), _p(coord), _p(arg), M, n, c_f(threshold))
    return coord, arg


def upsample_shifts(in2, old_level, new_level, new_cx, new_cy, old_t, new_t):
    in2 = _a(in2, np.float32)
    ocy, ocx = in2.shape[:2]
    out = np.zeros((new_cy, new_cx, 2), np.float32)
    lib().orc_upsample_shifts(_p(in2), _p(out), old_level, new_level, ocx, ocy, new_cx, new_cy, old_t, new_t)
    return out


def consolidate_shifts(measured, pair_from, pair_to, image_count, tiles_x, tiles_y, reference_image, pair_valid=None):
    measured = _a(measured, np.float32)
    nt, m = measured.shape[:2]
    n1 = image_count - 1
    one = np.zeros((nt, n1, 2), np.float32)
    fs = np.zeros((image_count, tiles_y, tiles_x, 2), np.float32)
    status = np.zeros((nt,), np.int32)
    pv = _a(pair_valid, np.uint8) if pair_valid is not None else None
    lib().orc_consolidate_shifts_masked(_p(measured), _i(pair_from), _i(pair_to), _p(pv), m, image_count, tiles_x, tiles_y, reference_image,
                                        _p(one), _p(fs), _p(status))
    return one, fs, status


def prealign_table():
    """(cos, sin) of the pre-alignment's candidate angles: 0.125 degree steps over +-21 degrees, double -> float32 (the same table
    the product uploads, csrc/pipeline.cu)."""
    i = np.arange(2 * 21 * 8 + 1, dtype=np.float64) - 21 * 8
    th = i * (0.125 * 3.14159265358979323846 / 180.0)
    return np.stack([np.cos(th), np.sin(th)], axis=1).astype(np.float32), 21 * 8


def prealign_search(ref, mov, cs, idx0, step, n_ang, cx=0, cy=0, radius=8, sub=1):
    ref, mov, cs = _a(ref, np.uint8), _a(mov, np.uint8), _a(cs, np.float32)
    h, w = ref.shape
    out = np.zeros((3,), np.int32)
    lib().orc_prealign_search(_p(ref), _p(mov), w, h, _p(cs), int(idx0), int(step), int(n_ang), int(cx), int(cy), int(radius), int(sub), _p(out))
    return out


def prealign_frame(pyr_ref, pyr_mov, w, h):
    """Pose (bx, by, cos, sin) of one frame against the reference: the two search stages of the restated pre-alignment host
    (csrc/prealign.cu header) on the tracking pyramids pyr_*[level] (extended below the matcher's levels by pyramid_down)."""
    cs, zero = prealign_table()
    la, aw, ah = 0, w, h
    while (aw > 192 or ah > 192) and aw // 2 >= 16 and ah // 2 >= 16:
        aw //= 2; ah //= 2; la += 1
    lb = la - 2 if la >= 2 else 0
    ra = prealign_search(pyr_ref[la], pyr_mov[la], cs, zero - 20 * 8, 8, 41, 0, 0, 8, 1)
    ta = zero if ra[0] < 0 else zero - 20 * 8 + int(ra[0]) * 8
    sc = 1 << (la - lb)
    rb = prealign_search(pyr_ref[lb], pyr_mov[lb], cs, ta - 8, 1, 17, int(ra[1]) * sc, int(ra[2]) * sc, 4, 2)
    tb = zero if rb[0] < 0 else ta - 8 + int(rb[0])
    bx, by = (0, 0) if rb[0] < 0 else (int(rb[1]), int(rb[2]))
    return np.array([np.float32(bx * (1 << lb)), np.float32(by * (1 << lb)), cs[tb, 0], cs[tb, 1]], np.float32), (ra, rb)


def pair_pose(pi, pj):
    """Pose of frame j relative to frame i (strict fp32, same expressions as pair_pose_kernel)."""
    f = np.float32
    dbx, dby = f(pj[0] - pi[0]), f(pj[1] - pi[1])
    ci, si, cj, sj = pi[2], pi[3], pj[2], pj[3]
    return np.array([f(f(ci * dbx) - f(si * dby)), f(f(si * dbx) + f(ci * dby)), f(f(cj * ci) + f(sj * si)), f(f(sj * ci) - f(cj * si))], np.float32)


def tile_align_cs(ref, mov, pre, T, M, bs, cf, sf, threshold=0.0):
    ref, mov = _a(ref, np.uint8), _a(mov, np.uint8)
    h, w = ref.shape
    tx, ty = (w - 2 * M) // T, (h - 2 * M) // T
    shift = np.zeros((ty, tx, 2), np.float32)
    arg = np.zeros((ty, tx, 2), np.int32)
    pre_a = _a(pre, np.float32) if pre is not None else None
    lib().orc_tile_align_cs(_p(ref), _p(mov), w, h, _p(pre_a), _p(shift), _p(arg), None, T, M, tx, ty,
                            c_f(bs[0]), c_f(bs[1]), c_f(cf), c_f(sf), c_f(threshold))
    return shift, arg


def flow_from_tiles_cs(tile2, T, w, h, bs, cf, sf):
    tile2 = _a(tile2, np.float32)
    ty, tx = tile2.shape[:2]
    flow = np.zeros((h, w, 2), np.float32)
    lib().orc_flow_from_tiles_cs(_p(tile2), tx, ty, T, _p(flow), w, h, c_f(bs[0]), c_f(bs[1]), c_f(cf), c_f(sf))
    return flow


def flow_from_tiles(tile2, T, w, h, base_shift=(0.0, 0.0), rot=0.0):
    tile2 = _a(tile2, np.float32)
    ty, tx = tile2.shape[:2]
    flow = np.zeros((h, w, 2), np.float32)
    lib().orc_flow_from_tiles(_p(tile2), tx, ty, T, _p(flow), w, h, c_f(base_shift[0]), c_f(base_shift[1]), c_f(rot))
    return flow


def warp(flow, img):
    flow, img = _a(flow, np.float32), _a(img, np.float32)
    h, w = img.shape
    out = np.zeros_like(img)
    lib().orc_warp(_p(flow), _p(img), _p(out), w, h)
    return out


def derivatives(src, tgt):
    src, tgt = _a(src, np.float32), _a(tgt, np.float32)
    h, w = src.shape
    ix, iy, iz = np.zeros_like(src), np.zeros_like(src), np.zeros_like(src)
    lib().orc_derivatives(_p(src), _p(tgt), _p(ix), _p(iy), _p(iz), w, h)
    return ix, iy, iz


def derivatives2(img):
    img = _a(img, np.float32)
    h, w = img.shape
    ix, iy = np.zeros_like(img), np.zeros_like(img)
    lib().orc_derivatives2(_p(img), _p(ix), _p(iy), w, h)
    return ix, iy


def lucas_kanade(flow, ix, iy, it, half_window, min_det):
    flow = _a(flow, np.float32).copy()
    ix, iy, it = _a(ix, np.float32), _a(iy, np.float32), _a(it, np.float32)
    h, w = ix.shape
    lib().orc_lucas_kanade(_p(flow), _p(ix), _p(iy), _p(it), w, h, half_window, c_f(min_det))
    return flow


def lk_iteration(ref, mov, flow, half_window=3, min_det=1e-3):
    ref, mov, flow = _a(ref, np.float32), _a(mov, np.float32), _a(flow, np.float32)
    h, w = ref.shape
    out = np.zeros_like(flow)
    lib().orc_lk_iteration(_p(ref), _p(mov), _p(flow), _p(out), w, h, half_window, c_f(min_det))
    return out


def structure_tensor(ix, iy):
    ix, iy = _a(ix, np.float32), _a(iy, np.float32)
    h, w = ix.shape
    t = np.zeros((h, w, 3), np.float32)
    lib().orc_structure_tensor(_p(ix), _p(iy), _p(t), w, h)
    return t


def kernel_param(t3, Dth, Dtr, kDetail, kDenoise, kStretch, kShrink):
    k = _a(t3, np.float32).copy()
    h, w = k.shape[:2]
    lib().orc_kernel_param(_p(k), w, h, c_f(Dth), c_f(Dtr), c_f(kDetail), c_f(kDenoise), c_f(kStretch), c_f(kShrink))
    return k


def kernel_params(gray, box_radius=2, Dth=0.005, Dtr=0.012, kDetail=0.3, kDenoise=4.0, kStretch=4.0, kShrink=2.0):
    gray = _a(gray, np.float32)
    h, w = gray.shape
    out = np.zeros((h, w, 4), np.float32)
    lib().orc_kernel_params(_p(gray), _p(out), w, h, box_radius, c_f(Dth), c_f(Dtr), c_f(kDetail), c_f(kDenoise), c_f(kStretch), c_f(kShrink))
    return out


def robustness_mask(ref3, mov3, flow, alpha, beta, threshold_m, erode_radius=0):
    ref3, mov3, flow = _a(ref3, np.float32), _a(mov3, np.float32), _a(flow, np.float32)
    h, w = ref3.shape[:2]
    fh, fw = flow.shape[:2]
    mask = np.zeros((h, w, 4), np.float32)
    lib().orc_robustness_mask(_p(ref3), _p(mov3), _p(mask), _p(flow), fw, fh, w, h, c_f(alpha), c_f(beta), c_f(threshold_m))
    if erode_radius > 0:
        out = np.zeros_like(mask)
        lib().orc_mask_erode(_p(mask), _p(out), w, h, erode_radius)
        mask = out
    return mask


def accumulate(raw, sum3, weight3, mask4, kernel4, flow2, geom: Geom, cfa, white, black):
    raw = _a(raw, np.uint16)
    mask4, kernel4, flow2 = _a(mask4, np.float32), _a(kernel4, np.float32), _a(flow2, np.float32)
    assert sum3.dtype == np.float32 and weight3.dtype == np.float32 and sum3.flags.c_contiguous and weight3.flags.c_contiguous
    lib().orc_accumulate(_p(raw), _p(sum3), _p(weight3), _p(mask4), _p(kernel4), _p(flow2), C.byref(geom), _i(cfa), _f(white), _f(black))


def apply_weighting(inout3, final3, weight3, threshold):
    io = _a(inout3, np.float32).copy()
    h, w = io.shape[:2]
    lib().orc_apply_weighting(_p(io), _p(_a(final3, np.float32)), _p(_a(weight3, np.float32)), w, h, c_f(threshold))
    return io


def gamma_srgb(img3):
    im = _a(img3, np.float32).copy()
    h, w = im.shape[:2]
    lib().orc_gamma_srgb(_p(im), w, h)
    return im


def fallback_upsample(rgb3, geom: Geom):
    rgb3 = _a(rgb3, np.float32)
    h, w = rgb3.shape[:2]
    out = np.zeros((geom.out_h, geom.out_w, 3), np.float32)
    lib().orc_fallback_upsample(_p(rgb3), w, h, _p(out), C.byref(geom))
    return out


def merge(raw, mask, flow, kernel4, fallback, geom: Geom, white, black, threshold, cfa, gamma=False, want_accumulators=False):
    """The reference chain: N x accumulate (RMW) -> ApplyWeighting -> optional GammasRGB."""
    n = raw.shape[0]
    s = np.zeros((geom.out_h, geom.out_w, 3), np.float32)
    wt = np.zeros_like(s)
    for f in range(n):
        accumulate(raw[f], s, wt, mask[f], kernel4, flow[f], geom, cfa, white, black)
    fb = fallback if fallback is not None else np.zeros_like(s)
    out = apply_weighting(fb, s, wt, threshold)
    if gamma:
        out = gamma_srgb(out)
    return (out, s, wt) if want_accumulators else out


def default_params(**kw):
    """The defaults the oracle shares with the product (mfsr_default_params, csrc/pipeline.cu), as a plain namespace: the CPU arm
    of bench.py runs from here without loading the product package."""
    from types import SimpleNamespace
    d = dict(scale=2, full_frame=1, cfa=[0, 1, 1, 2], black_level=[64.0] * 3, white_level=[959.0] * 3, tile_size=16, max_shift=4,
             levels=4, pair_span=2, track_bits=7, track_sigma=0.5, min_threshold=1024.0, base_shift=[0.0, 0.0], base_rotation=0.0,
             lk_iterations=3, lk_half_window=3, lk_min_det=1e-3, Dth=0.005, Dtr=0.012, kDetail=0.3, kDenoise=4.0, kStretch=4.0,
             kShrink=2.0, tensor_box_radius=2, alpha=1e-3, beta=1e-5, thresholdM=0.8, mask_erode_radius=2,
             weight_threshold=0.1, merge_flags=0, prealign=0, lk_texture=0)
    d.update(kw)
    return SimpleNamespace(**d)


# ---------------------------------------------------------------------------
# whole pipeline, stage by stage (the restated host of DESIGN.md §3), CPU only
# ---------------------------------------------------------------------------
def run_pipeline(frames, p, ref_idx=0, gray_format=False, keep=False):
    """frames: uint16 [N,H,W]; p: any object with the mfsr_params fields.  Returns (image, intermediates)."""
    frames = _a(frames, np.uint16)
    n, h, w = frames.shape
    cfa = [1, 1, 1, 1] if gray_format else list(p.cfa)
    black, white = list(p.black_level), list(p.white_level)
    scale = [np.float32(1.0) / np.float32(x) for x in white]
    max_val = np.float32(white[1]) + np.float32(black[1])
    T, M = p.tile_size, p.max_shift
    rgb_half = [subsample3(frames[f], max_val, cfa) for f in range(n)]
    gray, pyr = [], []
    rgb_ref = None
    for f in range(n):
        rgb = demosaic(frames[f], black, scale, cfa)
        if f == ref_idx:
            rgb_ref = rgb
        gr, gq = tracking_image(rgb, p.track_sigma, p.track_bits)
        gray.append(gr)
        lv = [gq]
        for _ in range(1, p.levels):
            if (lv[-1].shape[1] // 2 - 2 * M) // T < 1 or (lv[-1].shape[0] // 2 - 2 * M) // T < 1:
                break
            lv.append(pyramid_down(lv[-1]))
        pyr.append(lv)
    L = len(pyr[0])
    pre_on = bool(getattr(p, "prealign", 0)) and n > 1
    # measured pairs: at most pair_span frames apart; with the pre-alignment also every frame against the reference (pipeline.cu: build_pairs)
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n) if j - i <= p.pair_span or (pre_on and (i == ref_idx or j == ref_idx))]
    poses, pa_results, pair_valid = None, None, None
    if getattr(p, "prealign", 0) and n > 1:
        # the tracking pyramid continues below the matcher's levels for the search on a small image
        ext = []
        for f in range(n):
            lv = list(pyr[f])
            while max(lv[-1].shape) > 192 and min(lv[-1].shape) // 2 >= 16:
                lv.append(pyramid_down(lv[-1]))
            ext.append(lv)
        poses, pa_results = [], []
        for f in range(n):
            if f == ref_idx:
                poses.append(np.array([0, 0, 1, 0], np.float32)); pa_results.append(None)
            else:
                ps, rr = prealign_frame(ext[ref_idx], ext[f], w, h)
                poses.append(ps); pa_results.append(rr)
        # pairs rotated against each other by more than 16 degrees stay out of the consolidation (prealign.cu: pair_pose_kernel)
        pair_valid = [1 if pair_pose(poses[i], poses[j])[2] >= np.float32(0.96126169593831886) else 0 for (i, j) in pairs]
    tx, ty = (w - 2 * M) // T, (h - 2 * M) // T
    nt = tx * ty
    argmins = []
    if n > 1:
        measured = np.zeros((nt, len(pairs), 2), np.float32)
        for k, (i, j) in enumerate(pairs):
            pre = None
            for l in range(L - 1, -1, -1):
                lh, lw = pyr[0][l].shape
                ltx, lty = (lw - 2 * M) // T, (lh - 2 * M) // T
                if pre is not None:
                    pre = upsample_shifts(pre, 1 << (l + 1), 1 << l, ltx, lty, T, T)
                if poses is not None:
                    pp = pair_pose(poses[i], poses[j])
                    sc = np.float32(1.0) / np.float32(1 << l)
                    pre, arg = tile_align_cs(pyr[i][l], pyr[j][l], pre, T, M, (np.float32(pp[0] * sc), np.float32(pp[1] * sc)), pp[2], pp[3], p.min_threshold)
                else:
                    bs = (p.base_shift[0] / float(1 << l), p.base_shift[1] / float(1 << l))
                    pre, arg, _ = tile_align(pyr[i][l], pyr[j][l], pre, T, M, bs, p.base_rotation, p.min_threshold)
            measured[:, k, :] = pre.reshape(nt, 2)
            argmins.append(arg)
        one, frame_shift, status = consolidate_shifts(measured, [a for a, _ in pairs], [b for _, b in pairs], n, tx, ty, ref_idx, pair_valid)
    else:
        frame_shift = np.zeros((1, ty, tx, 2), np.float32)
    flows = []
    for f in range(n):
        if poses is not None:
            fl = flow_from_tiles_cs(frame_shift[f], T, w, h, (poses[f][0], poses[f][1]), poses[f][2], poses[f][3])
        else:
            fl = flow_from_tiles(frame_shift[f], T, w, h, tuple(p.base_shift), p.base_rotation)
        if f != ref_idx:
            for _ in range(p.lk_iterations):
                fl = lk_iteration(gray[ref_idx], gray[f], fl, p.lk_half_window, p.lk_min_det)
        flows.append(fl)
    kern = kernel_params(gray[ref_idx], p.tensor_box_radius, p.Dth, p.Dtr, p.kDetail, p.kDenoise, p.kStretch, p.kShrink)
    masks = [robustness_mask(rgb_half[ref_idx], rgb_half[f], flows[f], p.alpha, p.beta, p.thresholdM, p.mask_erode_radius) for f in range(n)]
    if p.full_frame:
        geom = Geom.full_frame(w, h, p.scale)
    else:
        s = p.scale
        ox, oy = w * (s - 1) // 2, h * (s - 1) // 2
        geom = Geom(w, h, s, w, h, ox, oy, ox // s, ox // s + w // s - 1, oy // s, oy // s + h // s - 1)
    fb = fallback_upsample(rgb_ref, geom)
    out = merge(frames, np.stack(masks), np.stack(flows), kern, fb, geom, white, black, p.weight_threshold, cfa,
                gamma=bool(p.merge_flags & 1))
    inter = dict(argmin=argmins, frame_shift=frame_shift, flow=flows, mask=masks, kernel=kern, fallback=fb, gray=gray,
                 gray_q=[pv[0] for pv in pyr], rgb_half=rgb_half, pairs=pairs, poses=poses, prealign=pa_results) if keep else None
    return out, inter

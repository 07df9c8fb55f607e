"""Pins the CPU oracle (oracle/mfsr_oracle.c) against golden vectors produced by the REFERENCE's OWN
kernels (test_opencv/*.cu compiled unmodified, run on a B200 by tests/golden/make_ref_golden.py).

Tolerances: the reference build contracts a*b+c into FMA and uses CUDA's expf/powf/atan2f, the oracle is
strict IEEE with glibc's libm -> values agree to a few ulp; integer results (arg-min) are exact."""
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as O

G = np.load(Path(__file__).resolve().parent / "golden" / "ref_golden.npz")
WHITE, BLACK = [959.0, 959.0, 959.0], [64.0, 64.0, 64.0]
SCALE = [np.float32(1.0) / np.float32(959.0)] * 3
RGGB = [0, 1, 1, 2]


def close(a, b, atol, rtol=0.0):
    assert a.shape == b.shape
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    d = np.abs(a[fin].astype(np.float64) - b[fin].astype(np.float64))
    lim = atol + rtol * np.abs(b[fin])
    assert np.all(d <= lim), f"max excess {float((d - lim).max()):.3e}, max abs {float(d.max()):.3e}"


def test_front_end():
    close(O.demosaic(G["fe_raw"], BLACK, SCALE, RGGB), G["fe_rgb"], 2e-6)
    close(O.demosaic(G["fe_raw"], BLACK, SCALE, [2, 1, 1, 0]), G["fe_rgb_bggr"], 2e-6)
    close(O.subsample3(G["fe_raw"], 1023.0, RGGB), G["fe_half"], 1e-7)
    assert np.all(G["fe_rgb"][:2] == 0) and np.all(G["fe_rgb"][:, -2:] == 0)       # unwritten 2-px border


@pytest.mark.parametrize("name", ["nopre", "pre"])
def test_tile_align_direct_cc(name):
    """Reference chain with the direct cross-correlation: SSD map bit-exact, integer arg-min exact."""
    pre = G["ta_pre"] if name == "pre" else None
    shift, arg, ssd = O.tile_align(G["ta_ref"], G["ta_mov"], pre, 16, 4)
    assert np.array_equal(ssd, G[f"ta_ssd_{name}_fft0"])
    ref_idx = G[f"ta_ssd_{name}_fft0"].argmin(axis=1)
    assert np.array_equal(arg.reshape(-1, 2), np.stack([ref_idx % 9 - 4, ref_idx // 9 - 4], 1))
    coord = G[f"ta_coord_{name}_fft0"]
    disp = np.round(pre) if pre is not None else 0.0          # out_shift = coord + round(pre_shift)
    close(shift - disp, coord, 1e-5)
    if name == "nopre":                                        # ground truth of the fixture: ref(p) = mov(p + (-3, +2))
        assert np.all(arg[..., 0] == -3) and np.all(arg[..., 1] == 2)


@pytest.mark.parametrize("name", ["nopre", "pre"])
def test_tile_align_cufft_cc(name):
    """Reference chain with cuFFT (as upstream): SSD equal up to FFT round-off, same arg-min on this fixture."""
    pre = G["ta_pre"] if name == "pre" else None
    _, arg, ssd = O.tile_align(G["ta_ref"], G["ta_mov"], pre, 16, 4)
    ref = G[f"ta_ssd_{name}_fft1"]
    assert np.abs(ssd - ref).max() <= 2e-6 * np.abs(ref).max() + 4.0
    ref_idx = ref.argmin(axis=1)
    assert np.array_equal(arg.reshape(-1, 2), np.stack([ref_idx % 9 - 4, ref_idx // 9 - 4], 1))


def test_upsample_shifts():
    close(O.upsample_shifts(G["up_in"], 4, 2, 15, 11, 16, 16), G["up_out"], 2e-6)


def test_flow_from_tiles_texture_model():
    got = O.flow_from_tiles(G["ff_tiles"], 16, 72, 56)
    d = np.abs(got - G["ff_flow"])
    assert d.max() <= 6.0 / 256.0 and np.median(d) <= 2e-5        # tile shifts differ by up to 6 px between neighbours


def test_warp_derivatives_lk():
    warped = O.warp(G["of_flow"], G["of_b"])
    d = np.abs(warped - G["of_warped"])
    assert d.max() <= 1.5e-3 and (d > 1e-4).mean() < 0.08          # one LSB of the 1.8 fraction on a few % of fetches
    ix, iy, iz = O.derivatives(G["of_warped"], G["of_a"])
    close(ix, G["of_ix"], 1e-6); close(iy, G["of_iy"], 1e-6); close(iz, G["of_iz"], 1e-7)
    lk = O.lucas_kanade(G["of_flow"], G["of_ix"], G["of_iy"], G["of_iz"], 3, 1e-3)
    close(lk, G["of_lk"], 2e-4)
    assert np.array_equal(lk[:3], G["of_flow"][:3])                 # border untouched (opticalFlow.cu:205)


def test_kernel_params():
    ix, iy = O.derivatives2(G["of_a"])
    close(ix, G["kp_ix"], 1e-6); close(iy, G["kp_iy"], 1e-6)
    close(O.structure_tensor(G["kp_ix"], G["kp_iy"]), G["kp_tensor"], 1e-9, 1e-6)
    got, ref = O.kernel_param(G["kp_tensor"], 0.005, 0.012, 0.3, 4.0, 4.0, 2.0), G["kp_kernel"]
    # the off-diagonal term is a cancelling difference (kernel.cu:780): compare relative to the matrix norm
    norm = np.abs(ref).max(axis=-1, keepdims=True)
    assert np.all(np.abs(got - ref) <= 5e-5 * norm + 1e-6)


def test_robustness_mask():
    got = O.robustness_mask(G["rb_ref"], G["rb_mov"], G["of_flow"], 1e-3, 1e-5, 0.8)
    close(got, G["rb_mask"], 2e-5, 1e-5)
    assert np.all(G["rb_mask"][0] == 0) and got[..., :3].max() <= 1.0


def test_merge_superres_and_1x():
    n, h, w = G["mg_raw"].shape
    out, s, wt = O.merge(G["mg_raw"], G["mg_mask"], G["mg_flow"], G["mg_kernel"], G["mg_fallback"], O.Geom.reference(w, h),
                         WHITE, BLACK, 0.1, RGGB, gamma=True, want_accumulators=True)
    close(s, G["mg_sum"], 2e-5, 2e-5); close(wt, G["mg_weight"], 2e-5, 2e-5)
    close(out, G["mg_out"], 1e-4)                                   # north-star tolerance is 1e-3
    out1 = O.merge(G["mg_raw"], G["mg_mask"], G["mg_flow"], G["mg_kernel"], G["mg_fallback"], O.Geom.full_frame(w, h, 1),
                   WHITE, BLACK, 0.1, RGGB)
    close(out1, G["mg_out_1x"], 1e-4)


def test_texture_model():
    """1.8 fixed-point fraction rounded to nearest reproduces the texture unit: exactly at the quarter
    positions accumulateImagesSuperRes uses (DeBayerKernels.cu:398), to one LSB elsewhere."""
    tw = int(G["tx_width"][0])
    for name, min_exact in (("quarter", 1.0), ("rand", 0.95)):
        xn, hw = G[f"tx_xn_{name}"], G[f"tx_out_{name}"]
        xb = (xn * np.float32(tw)).astype(np.float32) - np.float32(0.5)
        f = np.floor(xb)
        q = np.floor((xb - f).astype(np.float32) * 256 + 0.5) / 256
        model = np.clip(f, 0, tw - 1) * (1 - q) + np.clip(f + 1, 0, tw - 1) * q
        d = np.abs(model - hw)
        assert d.max() <= 1.0 / 256 + 1e-6 and (d < 1e-6).mean() >= min_exact


G2 = np.load(Path(__file__).resolve().parent / "golden" / "ref_golden_r2.npz")


@pytest.mark.parametrize("ci", [0, 1, 2, 3, 4])
def test_consolidate_shifts_vs_reference_kernels(ci):
    """Shift consolidation pinned to the reference's own kernels (copyShiftMatrix / setPointers / transposeShifts / checkForOutliers /
    getOptimalShifts, ShiftMinimizerKernels.cu:29-218) run on a B200 around cuBLAS batched normal equations
    (oracle/ref_driver.cu: ref_consolidate; vectors frozen by tests/golden/make_ref_golden_r2.py): the same measurements are removed
    per tile, the shifts agree to the round-off of two different fp32 inverses (Gauss-Jordan here, cuBLAS matinv there)."""
    n, span, ref, tx, ty = [int(v) for v in G2[f"cs{ci}_cfg"]]
    pairs = [(i, j) for i in range(n) for j in range(i + 1, min(n, i + span + 1))]
    one, fs, st = O.consolidate_shifts(G2[f"cs{ci}_meas"], [a for a, _ in pairs], [b for _, b in pairs], n, tx, ty, ref)
    assert np.all(G2[f"cs{ci}_status"] == -1)                      # the reference loop converged everywhere
    assert np.array_equal(st, G2[f"cs{ci}_removed"])
    close(one, G2[f"cs{ci}_one_to_one"], 2e-5)
    close(fs, G2[f"cs{ci}_frame_shift"], 1e-4)
    if len(pairs) > n:
        assert G2[f"cs{ci}_removed"].sum() == len(range(0, tx * ty, 3))     # one injected outlier in every third tile, each found

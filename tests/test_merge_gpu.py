"""Merge stage parity: mfsr_stage_merge (CUDA, through the C ABI) vs the oracle's restatement of
N x accumulateImagesSuperRes + ApplyWeighting + GammasRGB, and vs the reference's own kernels.

Tolerance (north star): max-abs <= 1e-3 on [0,1] and PSNR >= 60 dB."""
import numpy as np
import pytest
import torch

from multi_frame_super_resolution_b200 import stages
from multi_frame_super_resolution_b200._lib import MergeGeom
from multi_frame_super_resolution_b200.synth import synth_merge_inputs
from oracle import pyoracle as O
from oracle import pyref
from util import BLACK, RGGB, WHITE, max_abs, psnr, u16

pytestmark = pytest.mark.gpu
TOL_MAXABS, TOL_PSNR = 1e-3, 60.0


def _inputs(n, h, w, seed, dev):
    raw, mask, flow, kern = synth_merge_inputs(n, h, w, seed=seed, device="cpu")
    g = torch.Generator().manual_seed(seed)
    return raw, mask, flow, kern, g


def _run_both(raw, mask, flow, kern, fb, geom, dev, cfa=RGGB, threshold=0.1, gamma=False):
    out, s, wt = stages.merge(raw.to(dev), mask.to(dev), flow.to(dev), kern.to(dev), fb.to(dev) if fb is not None else None,
                              geom, WHITE, BLACK, threshold, cfa=cfa, flags=1 if gamma else 0, want_accumulators=True)
    torch.cuda.synchronize()
    exp, es, ew = O.merge(u16(raw), mask.numpy(), flow.numpy(), kern.numpy(), fb.numpy() if fb is not None else None,
                          O.Geom.from_product(geom), WHITE, BLACK, threshold, cfa, gamma=gamma, want_accumulators=True)
    return out.cpu().numpy(), s.cpu().numpy(), wt.cpu().numpy(), exp, es, ew


@pytest.mark.parametrize("n,h,w", [(1, 64, 96), (5, 128, 160), (8, 96, 128)])
def test_merge_reference_geometry_vs_oracle(cuda_device, n, h, w):
    raw, mask, flow, kern, g = _inputs(n, h, w, 11 + n, cuda_device)
    geom = MergeGeom.reference(w, h)
    fb = torch.rand((geom.out_h, geom.out_w, 3), generator=g)
    out, s, wt, exp, es, ew = _run_both(raw, mask, flow, kern, fb, geom, cuda_device)
    assert max_abs(out, exp) <= TOL_MAXABS and psnr(out, exp) >= TOL_PSNR
    # accumulators agree to fp32 round-off of a 25*N term sum
    assert np.allclose(s, es, rtol=2e-5, atol=2e-5) and np.allclose(wt, ew, rtol=2e-5, atol=2e-5)
    # border of the output window is skipped by the reference (DeBayerKernels.cu:391): fallback only
    assert np.array_equal(out[0], fb.numpy()[0]) and np.array_equal(out[:, -1], fb.numpy()[:, -1])


@pytest.mark.parametrize("scale", [1, 2, 3, 4])
def test_merge_full_frame_scales_vs_oracle(cuda_device, scale):
    """Scale 1: accumulateImages; 2: the slot kernel; 3, 4: the lean tap loop (merge_lean_kernel); all against the oracle's tap loop."""
    n, h, w = 4, 64, 96
    raw, mask, flow, kern, g = _inputs(n, h, w, 31 + scale, cuda_device)
    geom = MergeGeom.full_frame(w, h, scale)
    fb = torch.rand((geom.out_h, geom.out_w, 3), generator=g)
    out, s, wt, exp, es, ew = _run_both(raw, mask, flow, kern, fb, geom, cuda_device, gamma=(scale == 2))
    assert max_abs(out, exp) <= TOL_MAXABS and psnr(out, exp) >= TOL_PSNR


def test_merge_rational_scale_vs_oracle(cuda_device):
    """1.5x (MFSR_SCALE_RATIONAL(3, 2)): every "/ s" of the tap arithmetic becomes "* den / num"; and num / 1 is the integer scale, bit for bit."""
    from multi_frame_super_resolution_b200._lib import scale_rational
    n, h, w = 4, 64, 96
    raw, mask, flow, kern, g = _inputs(n, h, w, 91, cuda_device)
    geom = MergeGeom.full_frame(w, h, scale_rational(3, 2))
    assert (geom.out_w, geom.out_h) == (144, 96)
    fb = torch.rand((geom.out_h, geom.out_w, 3), generator=g)
    out, s, wt, exp, es, ew = _run_both(raw, mask, flow, kern, fb, geom, cuda_device)
    assert max_abs(out, exp) <= TOL_MAXABS and psnr(out, exp) >= TOL_PSNR
    assert np.allclose(s, es, rtol=2e-5, atol=2e-5) and np.allclose(wt, ew, rtol=2e-5, atol=2e-5)
    assert float(wt.max()) > 0.5                                   # the taps did land on samples
    # 3 / 1 through the rational path of the plain loop == scale 3 (lean loop) within the merge tolerance
    g3, g31 = MergeGeom.full_frame(w, h, 3), MergeGeom.full_frame(w, h, scale_rational(3, 1))
    fb3 = torch.rand((g3.out_h, g3.out_w, 3), generator=g)
    a = stages.merge(raw.to(cuda_device), mask.to(cuda_device), flow.to(cuda_device), kern.to(cuda_device), fb3.to(cuda_device), g3, WHITE, BLACK, 0.1)
    b = stages.merge(raw.to(cuda_device), mask.to(cuda_device), flow.to(cuda_device), kern.to(cuda_device), fb3.to(cuda_device), g31, WHITE, BLACK, 0.1)
    a = a[0] if isinstance(a, tuple) else a
    b = b[0] if isinstance(b, tuple) else b
    assert max_abs(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5


def test_merge_lean_kernel_large_shifts_and_bad_certainties(cuda_device):
    """Scale 3 (merge_lean_kernel) with large shifts (taps run into the clamp range) and non-finite certainties, against the oracle's
    tap loop, accumulators included."""
    n, h, w = 6, 96, 128
    raw, mask, flow, kern, g = _inputs(n, h, w, 77, cuda_device)
    flow = flow * 6.0                                      # shifts of up to ~ +-18 HR px at 3x: taps run into the clamp range
    mask[1, 5:9, 7:11, 0] = float("nan")
    mask[2, 20:22, 30:40, 1] = float("inf")
    geom = MergeGeom.full_frame(w, h, 3)
    fb = torch.rand((geom.out_h, geom.out_w, 3), generator=g)
    out, s, wt, exp, es, ew = _run_both(raw, mask, flow, kern, fb, geom, cuda_device)
    assert max_abs(out, exp) <= TOL_MAXABS and psnr(out, exp) >= TOL_PSNR
    assert np.allclose(s, es, rtol=5e-5, atol=5e-5) and np.allclose(wt, ew, rtol=5e-5, atol=5e-5)


def test_merge_edge_cases(cuda_device):
    """NaN/Inf certainty -> 0, non-finite kernel -> centre cross only (:429-430), huge shifts clamp, zero frames."""
    n, h, w = 3, 64, 64
    raw, mask, flow, kern, g = _inputs(n, h, w, 77, cuda_device)
    mask[0, 3:9, 4:12, :3] = float("nan")
    mask[1, 10:14, 2:30, 1] = float("inf")
    kern[5:20, 5:20, :3] = float("inf")
    kern[30:40, 10:30, 0] = float("nan")
    kern[41:50, 10:30, :3] = -50.0           # exp overflow -> +inf weights
    flow[2, 20:40, 20:40, :] = 500.0          # taps far outside -> clamped
    flow[1, 0:10, :, 0] = -300.0
    geom = MergeGeom.reference(w, h)
    fb = torch.rand((geom.out_h, geom.out_w, 3), generator=g)
    out, s, wt, exp, es, ew = _run_both(raw, mask, flow, kern, fb, geom, cuda_device)
    both_nan = np.isnan(out) & np.isnan(exp)
    assert np.array_equal(np.isnan(out), np.isnan(exp))
    assert max_abs(np.where(both_nan, 0, out), np.where(both_nan, 0, exp)) <= TOL_MAXABS
    # zero frames: pure fallback path of ApplyWeighting (w = 0 < threshold -> val = inout / 1)
    o0 = stages.merge(raw[:0].to(cuda_device), mask[:0].to(cuda_device), flow[:0].to(cuda_device), kern.to(cuda_device),
                      fb.to(cuda_device), geom, WHITE, BLACK, 0.1)
    assert np.array_equal(o0.cpu().numpy(), fb.numpy())


def test_merge_gray_cfa_and_no_fallback(cuda_device):
    n, h, w = 3, 64, 96
    raw, mask, flow, kern, g = _inputs(n, h, w, 5, cuda_device)
    geom = MergeGeom.reference(w, h)
    out, s, wt, exp, es, ew = _run_both(raw, mask, flow, kern, None, geom, cuda_device, cfa=[1, 1, 1, 1])
    assert max_abs(out, exp) <= TOL_MAXABS
    assert np.all(out[..., 0] == 0) and np.all(out[..., 2] == 0)      # only .y is fed by an all-green CFA


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libmfsr_ref.so not built")
@pytest.mark.parametrize("n,h,w", [(5, 128, 192), (8, 256, 256)])
def test_merge_vs_reference_kernels(cuda_device, n, h, w):
    """Same device buffers through the reference's own accumulateImagesSuperRes/ApplyWeighting/GammasRGB."""
    raw, mask, flow, kern, g = _inputs(n, h, w, 3, cuda_device)
    dev = cuda_device
    geom = MergeGeom.reference(w, h)
    fb = torch.rand((h, w, 3), generator=g).to(dev)
    rawd, maskd, flowd, kernd = raw.to(dev), mask.to(dev), flow.to(dev), kern.to(dev)
    out = stages.merge(rawd, maskd, flowd, kernd, fb, geom, WHITE, BLACK, 0.1, flags=1)
    ref = pyref.merge_superres(rawd, maskd, flowd, kernd, fb, WHITE, BLACK, 0.1, RGGB, gamma=True)
    o, r = out.cpu().numpy(), ref.cpu().numpy()
    bad = np.abs(o - r) > TOL_MAXABS
    # the texture unit's 1.8 fixed-point bilinear can flip roundf(2*shift) at exact .5 ties;
    # such pixels are counted, everything else must hold the tolerance.
    assert bad.mean() < 1e-4, f"{bad.sum()} of {bad.size} samples beyond 1e-3"
    assert psnr(o, r) >= TOL_PSNR


def test_merge_full_size_properties(cuda_device):
    """Size-independent properties at BASELINE config-2 size (12 MP x 8, 2x, full frame)."""
    n, h, w = 8, 3024, 4032
    dev = cuda_device
    raw, mask, flow, kern = synth_merge_inputs(n, h, w, seed=1234, device=dev)
    geom = MergeGeom.full_frame(w, h, 2)
    fb = torch.full((geom.out_h, geom.out_w, 3), 0.25, device=dev)
    out, s, wt = stages.merge(raw, mask, flow, kern, fb, geom, WHITE, BLACK, 0.1, want_accumulators=True)
    # (1) linearity in the certainty: doubling every mask doubles sum and weight exactly (power of two)
    out2, s2, wt2 = stages.merge(raw, mask * 2.0, flow, kern, fb, geom, WHITE, BLACK, 1e9, want_accumulators=True)
    assert torch.equal(s2, s * 2.0) and torch.equal(wt2, wt * 2.0)
    # (2) frame-permutation invariance of the frame loop up to fp32 re-association
    perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4], device=dev)
    outp = stages.merge(raw[perm].contiguous(), mask[perm].contiguous(), flow[perm].contiguous(), kern, fb, geom, WHITE, BLACK, 0.1)
    assert float((outp - out).abs().max()) <= 1e-4
    # (3) convex combination: where weight >= threshold the output lies within the normalised raw range
    lo, hi = (0 - 64.0) / 959.0, (1023 - 64.0) / 959.0
    ok = wt >= 0.1
    assert float(out[ok].min()) >= lo - 1e-4 and float(out[ok].max()) <= hi + 1e-4
    # (4) a rectangle of the full-frame result equals the same rectangle computed as its own window
    sub = MergeGeom(w, h, 2, 512, 256, 3000, 2000, 0, w - 1, 0, h - 1)
    fbs = fb[2000:2256, 3000:3512].contiguous()
    outs = stages.merge(raw, mask, flow, kern, fbs, sub, WHITE, BLACK, 0.1)
    assert torch.equal(outs[1:-1, 1:-1], out[2001:2255, 3001:3511])


@pytest.mark.parametrize("n", [12, 20, 33])
def test_merge_many_frames_tile_variants(cuda_device, n):
    """More frames than the 16-row tile variant can keep resident in shared memory: the 8- and 4-row variants of
    merge_s2_dyn (and, beyond them, the generic kernel) must give the same image."""
    h, w = 64, 96
    raw, mask, flow, kern, g = _inputs(n, h, w, 100 + n, cuda_device)
    flow[1, 10:30, 20:60, 0] += 9.0            # alignment outliers: window not staged -> global raw fetch
    flow[2, 40:50, :, 1] -= 40.0
    geom = MergeGeom.full_frame(w, h, 2)
    fb = torch.rand((geom.out_h, geom.out_w, 3), generator=g)
    # with sum / weight images requested the burst is merged in chunks of frames by the 16-row kernel (partial sums in place) ...
    out, s, wt, exp, es, ew = _run_both(raw, mask, flow, kern, fb, geom, cuda_device)
    assert max_abs(out, exp) <= TOL_MAXABS and psnr(out, exp) >= TOL_PSNR
    assert max_abs(s, es) <= 1e-3 * max(1.0, float(np.abs(es).max())) and max_abs(wt, ew) <= 1e-3 * max(1.0, float(np.abs(ew).max()))
    # ... without them every frame is resident at once in the 8- / 4-row variants (or the generic kernel runs)
    dev = cuda_device
    out2 = stages.merge(raw.to(dev), mask.to(dev), flow.to(dev), kern.to(dev), fb.to(dev), geom, WHITE, BLACK, 0.1).cpu().numpy()
    assert max_abs(out2, exp) <= TOL_MAXABS and psnr(out2, exp) >= TOL_PSNR
    assert max_abs(out2, out) <= 1e-5            # same arithmetic, only the association of the frame sums differs


def test_merge_window_offsets_and_odd_geometry(cuda_device):
    """Output windows whose origin is not a multiple of 4 (tile grid starts left of / above the window) and whose size is
    not a multiple of the tile; shifts near the +-127 limit of the char2 encoding and beyond it (sentinel path)."""
    n, h, w = 3, 96, 128
    raw, mask, flow, kern, g = _inputs(n, h, w, 321, cuda_device)
    flow[1, 20:30, 30:50, 0] = 63.4            # 2 * 63.4 -> 127
    flow[1, 30:40, 30:50, 0] = 64.0            # 128 -> outsized: recomputed per pixel
    flow[2, 50:60, 10:40, 1] = -70.0
    for (ow, oh, ox, oy) in ((101, 67, 3, 5), (64, 40, 130, 61), (200, 150, 1, 2)):
        geom = MergeGeom(w, h, 2, ow, oh, ox, oy, 0, w - 1, 0, h - 1)
        fb = torch.rand((oh, ow, 3), generator=g)
        out, s, wt, exp, es, ew = _run_both(raw, mask, flow, kern, fb, geom, cuda_device)
        assert max_abs(out, exp) <= TOL_MAXABS, (ow, oh, ox, oy)


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libmfsr_ref.so not built")
def test_merge_config2_size_vs_reference_kernels(cuda_device):
    """BASELINE configs[1] size (4032 x 3024 x 8 frames, 2x) through the reference's own accumulateImagesSuperRes x 8 +
    ApplyWeighting + GammasRGB (reference geometry: central crop, output dims == raw dims, DeBayerKernels.cu:398-423)
    and through mfsr_stage_merge on the same device buffers."""
    n, h, w = 8, 3024, 4032
    dev = cuda_device
    raw, mask, flow, kern = synth_merge_inputs(n, h, w, seed=4242, device=dev)
    geom = MergeGeom.reference(w, h)
    fb = torch.rand((h, w, 3), generator=torch.Generator().manual_seed(7)).to(dev)
    out = stages.merge(raw, mask, flow, kern, fb, geom, WHITE, BLACK, 0.1, flags=1)
    ref = pyref.merge_superres(raw, mask, flow, kern, fb, WHITE, BLACK, 0.1, RGGB, gamma=True)
    d = (out - ref).abs()
    bad = float((d > TOL_MAXABS).float().mean())
    mse = float((d.double() ** 2).mean())
    assert bad < 1e-4, f"{bad:.2e} of samples beyond 1e-3"
    assert 10.0 * np.log10(1.0 / max(mse, 1e-30)) >= TOL_PSNR

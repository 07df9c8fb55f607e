"""End-to-end: mfsr_create / mfsr_set_frames / mfsr_run vs the oracle's stage chain on the same burst."""
import numpy as np
import pytest
import torch

from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params, measured_pairs
from multi_frame_super_resolution_b200.synth import synth_burst
from oracle import pyoracle as O
from util import max_abs, psnr, u16

pytestmark = pytest.mark.gpu


def _run(p, fr, dev, ref_idx=0, host=False):
    n, h, w = fr.shape
    sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n)
    sr.set_input(u16(fr).copy() if host else fr.to(dev), ref_idx=ref_idx)
    out = sr.next_frame(host=host)
    sr.synchronize()
    return sr, (out.numpy() if host else out.cpu().numpy())


@pytest.mark.parametrize("full_frame", [1, 0])
def test_pipeline_vs_oracle(cuda_device, full_frame):
    p = default_params()
    p.full_frame = full_frame
    p.levels = 3
    fr, sh = synth_burst(5, 256, 320, seed=77)
    sr, out = _run(p, fr, cuda_device, ref_idx=2)
    exp, it = O.run_pipeline(u16(fr), p, ref_idx=2, keep=True)
    tx, ty, m = sr.tile_grid()
    assert m == len(measured_pairs(5, p.pair_span)) == len(it["pairs"])
    # integer tile shifts: bit-exact
    for k in range(m):
        assert np.array_equal(sr.tile_argmin(k), it["argmin"][k]), f"pair {k}"
    # consolidated tile shifts: bit-exact (strict fp32 on both sides)
    for f in range(5):
        assert np.array_equal(sr.tile_shifts(f), it["frame_shift"][f])
    # recovered motion ~ -ground truth (synth: frame(x) = scene(x + d)); sanity of the restated host, not parity
    for f in range(5):
        med = np.median(sr.tile_shifts(f).reshape(-1, 2), axis=0)
        gt = -(sh[f] - sh[2]).numpy()
        assert np.abs(med - gt).max() < 0.8, (f, med, gt)
    # intermediates, stage by stage
    h, w = fr.shape[1:]
    for f in range(5):
        gq = sr.buffer("gray_q0", h, w, f)
        assert np.array_equal(gq, it["gray_q"][f]), f"tracking image of frame {f}"
        half = sr.buffer("rgb_half", h // 2, (w // 2) * 12, f).view(np.float32).reshape(h // 2, w // 2, 3)
        assert np.array_equal(half, it["rgb_half"][f])
        flow = sr.buffer("flow", h, w * 8, f).view(np.float32).reshape(h, w, 2)
        dfl = np.abs(flow - it["flow"][f])
        assert np.percentile(dfl, 99.9) < 5e-3, (f, float(dfl.max()))
        mask = sr.buffer("mask", h // 2, (w // 2) * 16, f).view(np.float32).reshape(h // 2, w // 2, 4)
        dm = np.abs(mask - it["mask"][f])
        assert (dm > 1e-3).mean() < 5e-3, (f, float(dm.max()))
    kern = sr.buffer("kernel", h, w * 16).view(np.float32).reshape(h, w, 4)
    fin = np.isfinite(it["kernel"])
    assert np.array_equal(np.isfinite(kern), fin)
    assert np.percentile(np.abs(kern[fin] - it["kernel"][fin]) / (np.abs(it["kernel"][fin]) + 1e-3), 99.9) < 1e-3
    fb = sr.buffer("fallback", exp.shape[0], exp.shape[1] * 12).view(np.float32).reshape(exp.shape)
    assert np.array_equal(fb, it["fallback"])
    # merged image: tolerance of the north star, allowing isolated round(2*shift) flips caused by
    # ulp-level differences of atan2f/sinf/cosf inside the LK refinement (counted, must be rare)
    both = np.isfinite(out) & np.isfinite(exp)
    assert np.array_equal(np.isfinite(out), np.isfinite(exp))
    bad = np.abs(np.where(both, out, 0) - np.where(both, exp, 0)) > 1e-3
    # measured (tools/parity_margins.py, three seeds): <= 7e-6 of the samples beyond 1e-3, PSNR 90 - 148 dB
    assert bad.mean() < 1e-4, f"{bad.mean():.2e} of samples beyond 1e-3"
    assert psnr(np.where(both, out, 0), np.where(both, exp, 0)) >= 60.0            # the north star's bar
    st = sr.stage_ms()
    assert set(st) >= {"frontend", "align", "consolidate", "flow", "kernel_params", "robustness", "fallback", "merge"}
    assert sr.launch_count() >= 15            # per-frame kernels run once per burst (grid.z = frame): about 20 launches, not ~100
    sr.close()


def test_pipeline_host_buffers_and_reuse(cuda_device):
    """Host (pinned/pageable) frames in, host image out, handle reused for a second burst; single frame burst."""
    p = default_params()
    p.levels = 2
    fr, _ = synth_burst(3, 128, 192, seed=5)
    sr, out_dev = _run(p, fr, cuda_device)
    sr.set_input(u16(fr).copy(), ref_idx=0)
    out_host = sr.next_frame(host=True).numpy()
    assert np.array_equal(out_host, out_dev)
    fr1, _ = synth_burst(1, 128, 192, seed=6)
    sr.set_input(fr1.to(cuda_device))
    o1 = sr.next_frame().cpu().numpy()
    e1, _ = O.run_pipeline(u16(fr1), p)
    assert max_abs(o1, e1) <= 1e-3
    sr.close()


def test_pipeline_gray_format(cuda_device):
    p = default_params()
    p.levels = 2
    fr, _ = synth_burst(3, 128, 160, seed=9, bayer=False)
    n, h, w = fr.shape
    sr = BurstSuperResolution(p, 0, w, h, n)
    sr.set_input(fr.to(cuda_device), fmt=1)
    out = sr.next_frame().cpu().numpy()
    exp, _ = O.run_pipeline(u16(fr), p, gray_format=True)
    both = np.isfinite(out) & np.isfinite(exp)
    bad = np.abs(np.where(both, out, 0) - np.where(both, exp, 0)) > 1e-3
    assert np.array_equal(np.isfinite(out), np.isfinite(exp)) and bad.mean() < 2e-3, float(bad.mean())
    sr.close()


def test_api_errors(cuda_device):
    from multi_frame_super_resolution_b200._lib import MfsrError
    p = default_params()
    sr = BurstSuperResolution(p, 0, 128, 128, 2)
    with pytest.raises(RuntimeError):
        sr.next_frame()
    fr, _ = synth_burst(3, 128, 128, seed=1)
    with pytest.raises(MfsrError):
        sr.set_input(fr.to(cuda_device))          # more frames than the handle was sized for
    p2 = default_params(); p2.track_bits = 8      # violates the exact-sum contract for T=16
    with pytest.raises(MfsrError):
        BurstSuperResolution(p2, 0, 128, 128, 2).workspace_bytes
    sr.close()


def test_pipeline_async_two_handles(cuda_device):
    """mfsr_run_async on two alternating handles (the bench's e2e pattern) returns the same images as the blocking call."""
    p = default_params()
    p.levels = 2
    bursts = [synth_burst(3, 128, 192, seed=40 + i)[0] for i in range(4)]
    n, h, w = bursts[0].shape
    ref = []
    sr = BurstSuperResolution(p, 0, w, h, n)
    for b in bursts:
        sr.set_input(u16(b).copy())
        ref.append(sr.next_frame(host=True).numpy().copy())
    hs = [sr, BurstSuperResolution(p, 0, w, h, n)]
    pinned_in = [torch.from_numpy(u16(b).copy().view(np.int16)).pin_memory() for b in bursts]
    outs = [torch.empty_like(torch.from_numpy(ref[0])).pin_memory() for _ in bursts]
    for i, b in enumerate(bursts):
        hd = hs[i % 2]
        hd.synchronize()
        hd.set_input(pinned_in[i].numpy().view(np.uint16))
        hd.next_frame(out=outs[i], host=True, sync=False)
    for hd in hs:
        hd.synchronize()
    for i in range(4):
        assert np.array_equal(outs[i].numpy(), ref[i]), i
    for hd in hs:
        hd.close()


def test_output_formats(cuda_device):
    """mfsr_run_format: the half3 image is the float image rounded to nearest, the 8-bit image is floor(v*255+.5) saturated
    (what the reference program writes to its PNGs, multi_frame_sr.cpp:207); device and host delivery agree."""
    p = default_params()
    p.levels = 3
    p.merge_flags = 1                     # GammasRGB, like the reference's 8-bit result
    fr, _ = synth_burst(4, 192, 256, seed=5)
    sr = BurstSuperResolution(p, 0, 256, 192, 4)
    sr.set_input(fr.to(cuda_device))
    f32 = sr.next_frame().clone()
    f16 = sr.next_frame(dtype=torch.float16).clone()
    u8 = sr.next_frame(dtype=torch.uint8).clone()
    u8_host = sr.next_frame(host=True, dtype=torch.uint8)
    f16_host = sr.next_frame(host=True, sync=False, dtype=torch.float16)
    sr.synchronize()
    assert torch.equal(f16, f32.half())
    assert float((f16.float() - f32).abs().max()) <= 2.5e-4
    exp8 = torch.clamp(torch.floor(torch.nan_to_num(f32) * 255.0 + 0.5), 0, 255).to(torch.uint8)
    assert torch.equal(u8, exp8)
    assert torch.equal(u8_host, exp8.cpu()) and torch.equal(f16_host, f16.cpu())
    lib = sr._lib
    import ctypes as C
    assert lib.mfsr_run_format(sr._h, C.c_void_p(u8.data_ptr()), 256 * 2 * 3, 0, 7, 0) == -1          # unknown format
    assert lib.mfsr_run_format(sr._h, C.c_void_p(u8.data_ptr()), 10, 0, 2, 0) == -1                   # pitch below a row
    sr.close()


def test_pipeline_edge_bursts(cuda_device):
    """Edge cases of the burst path against the oracle: a single frame (no pairs, no flow: the merge of the reference alone),
    the smallest accepted frame (64 x 64, pyramid cut to the levels that fit), the LAST frame as reference, and a static burst
    (identical frames: every integer tile shift is exactly zero; the sub-pixel part is findMinimum's parabola through an asymmetric
    SSD valley, kernel.cu:560-636, so it is small but not zero, and it must equal the oracle's bit for bit)."""
    p = default_params()
    # single frame
    fr, _ = synth_burst(1, 128, 160, seed=9)
    p1 = default_params(); p1.levels = 2
    sr, out = _run(p1, fr, cuda_device)
    exp, _ = O.run_pipeline(u16(fr), p1)
    assert sr.tile_grid()[2] == 0
    assert np.array_equal(np.isfinite(out), np.isfinite(exp)) and max_abs(np.nan_to_num(out), np.nan_to_num(exp)) <= 1e-3
    sr.close()
    # smallest frame, reference = last frame
    fr, _ = synth_burst(3, 64, 64, seed=10)
    p2 = default_params(); p2.levels = 1
    sr, out = _run(p2, fr, cuda_device, ref_idx=2)
    exp, it = O.run_pipeline(u16(fr), p2, ref_idx=2, keep=True)
    for k in range(sr.tile_grid()[2]):
        assert np.array_equal(sr.tile_argmin(k), it["argmin"][k])
    bad = np.abs(np.nan_to_num(out) - np.nan_to_num(exp)) > 1e-3
    assert bad.mean() < 2e-3
    sr.close()
    # static burst: identical frames
    fr, _ = synth_burst(1, 192, 256, seed=11)
    st = fr.repeat(4, 1, 1).contiguous()
    p3 = default_params(); p3.levels = 3
    sr, out = _run(p3, st, cuda_device, ref_idx=1)
    tx, ty, m = sr.tile_grid()
    for k in range(m):
        assert not sr.tile_argmin(k).any(), f"pair {k}: non-zero shift between identical frames"
    _, it = O.run_pipeline(u16(st), p3, ref_idx=1, keep=True)
    for f in range(4):
        ts = sr.tile_shifts(f)
        assert np.array_equal(ts, it["frame_shift"][f]) and np.abs(ts).max() < 0.5
    assert np.isfinite(out).all()
    sr.close()
    # rejected geometries: odd width, below the minimum size
    from multi_frame_super_resolution_b200._lib import MfsrError
    with pytest.raises(MfsrError):
        BurstSuperResolution(p, 0, 63, 64, 2).workspace_bytes
    sr = BurstSuperResolution(p, 0, 256, 256, 2)
    with pytest.raises(MfsrError):
        sr.set_input(torch.zeros((2, 65, 64), dtype=torch.int16, device=cuda_device))
    sr.close()


def _compare_images(out, exp, frac=1e-4, db=60.0):
    assert np.array_equal(np.isfinite(out), np.isfinite(exp))
    both = np.isfinite(out) & np.isfinite(exp)
    a, b = np.where(both, out, 0), np.where(both, exp, 0)
    bad = np.abs(a - b) > 1e-3
    assert bad.mean() < frac, f"{bad.mean():.2e} of samples beyond 1e-3"
    assert psnr(a, b) >= db


def test_pipeline_bundled_burst_config1(cuda_device):
    """BASELINE configs[0]: the reference's own 5-frame burst (tests/golden/bundled_burst_rggb.npz, made from
    test_opencv/img_00000{0..4}.png by tests/golden/make_bundled_fixture.py; frames 2-4 rotated 5 / 10 / -15 degrees,
    main.cpp:1894-1907) through the CUDA path with the global pre-alignment against the oracle's chain: poses, integer tile
    shifts and consolidated shifts bit-exact, image within the north star's tolerance — and every frame ALIGNS."""
    from pathlib import Path
    fr_np = np.load(Path(__file__).resolve().parent / "golden" / "bundled_burst_rggb.npz")["frames"]
    fr = torch.from_numpy(fr_np.view(np.int16))
    p = default_params()
    p.prealign = 1
    sr, out = _run(p, fr, cuda_device, ref_idx=0)
    exp, it = O.run_pipeline(fr_np, p, ref_idx=0, keep=True)
    tx, ty, m = sr.tile_grid()
    assert m == len(it["pairs"]) == len(measured_pairs(5, p.pair_span, 1, 0))
    pose = sr.buffer("pose", 5, 16).view(np.float32).reshape(5, 4)
    assert np.array_equal(pose, np.stack(it["poses"])), (pose, it["poses"])
    ang = np.degrees(np.arctan2(pose[:, 3], pose[:, 2]))
    assert np.allclose(ang, [0, 0, -5, -10, 15], atol=0.3), ang           # the generator's rotations, in the kernels' sign convention
    for k in range(m):
        assert np.array_equal(sr.tile_argmin(k), it["argmin"][k]), f"pair {k}"
    for f in range(5):
        assert np.array_equal(sr.tile_shifts(f), it["frame_shift"][f])
    h, w = fr_np.shape[1:]
    for f in range(1, 5):
        ts = sr.tile_shifts(f).reshape(-1, 2)
        mask = sr.buffer("mask", h // 2, (w // 2) * 16, f).view(np.float32).reshape(h // 2, w // 2, 4)
        assert np.median(np.hypot(ts[:, 0], ts[:, 1])) < (1.5 if f < 4 else 3.0), f     # residual against the pre-alignment pose
        assert mask[..., :3].mean() > 0.4, (f, float(mask[..., :3].mean()))             # the robustness model accepts the frame
    # The photograph has flat and saturated regions where lucasKanadeOptim's 2x2 system is singular (opticalFlow.cu:232-262: the
    # minDet branch and the pseudo-inverse of a rank-deficient matrix): there the refined flow is decided by round-off, CUDA (MUFU
    # sqrt / rcp) and the oracle (libm) land on different sides of round(2 * flow) for ~1 % of the pixels, and 0.2 % of the image
    # samples move by more than 1e-3 (tools/bundled_debug.py).  Synthetic bursts, textured everywhere, stay below 1e-5.
    _compare_images(out, exp, frac=3e-3, db=40.0)
    sr.close()


def test_pipeline_config5_shape(cuda_device):
    """A config-5-shaped pipeline (3x scale, 4 pyramid levels, 16 frames, large motion) on a small image: integer tile
    shifts and consolidated shifts bit-exact, merged image within the north star's tolerance of the oracle's chain (the
    3x merge is only definable against the restatement generalised the same way)."""
    p = default_params()
    p.scale = 3
    p.levels = 4
    fr, sh = synth_burst(16, 320, 384, seed=505, max_shift=12.0)
    sr, out = _run(p, fr, cuda_device, ref_idx=0)
    exp, it = O.run_pipeline(u16(fr), p, ref_idx=0, keep=True)
    assert out.shape == (960, 1152, 3)
    for k in range(sr.tile_grid()[2]):
        assert np.array_equal(sr.tile_argmin(k), it["argmin"][k]), f"pair {k}"
    for f in range(16):
        assert np.array_equal(sr.tile_shifts(f), it["frame_shift"][f])
    _compare_images(out, exp)
    sr.close()


def test_pipeline_config2_full_size_vs_oracle(cuda_device):
    """BASELINE configs[1] at its full size (4032 x 3024 RGGB x 8, 2x, full frame): the whole CUDA chain against the oracle's
    whole chain (about 20 s of CPU on the box's cores)."""
    p = default_params()
    fr, _ = synth_burst(8, 3024, 4032, seed=1234)
    sr, out = _run(p, fr, cuda_device, ref_idx=0)
    exp, it = O.run_pipeline(u16(fr), p, ref_idx=0, keep=True)
    for k in range(sr.tile_grid()[2]):
        assert np.array_equal(sr.tile_argmin(k), it["argmin"][k]), f"pair {k}"
    for f in range(8):
        assert np.array_equal(sr.tile_shifts(f), it["frame_shift"][f])
    _compare_images(out, exp)
    sr.close()


def test_pipeline_rational_scale_1p5(cuda_device):
    """s = 1.5 (MFSR_SCALE_RATIONAL(3, 2)) through the whole chain against the oracle; rejected where it is not defined."""
    from multi_frame_super_resolution_b200._lib import MfsrError, scale_rational
    fr, _ = synth_burst(4, 128, 192, seed=17)
    p = default_params()
    p.levels = 2
    p.scale = scale_rational(3, 2)
    sr = BurstSuperResolution(p, 0, 192, 128, 4)
    assert sr.output_size(192, 128) == (288, 192)
    sr.set_input(fr.to(cuda_device))
    out = sr.next_frame().cpu().numpy()
    exp, _ = O.run_pipeline(u16(fr), p)
    assert out.shape == exp.shape == (192, 288, 3)
    _compare_images(out, exp)
    sr.close()
    q = default_params()
    q.scale = scale_rational(3, 2)
    q.full_frame = 0                                  # the reference's central crop is defined for integer scales only
    with pytest.raises(MfsrError):
        BurstSuperResolution(q, 0, 192, 128, 4).workspace_bytes
    q = default_params()
    q.scale = scale_rational(9, 2)                    # 4.5x: beyond the supported range
    with pytest.raises(MfsrError):
        BurstSuperResolution(q, 0, 192, 128, 4).workspace_bytes


def test_set_input_accepts_a_view_with_a_larger_frame_stride(cuda_device):
    """Frames that are dense but further apart than H*W (a row band inside the shared band buffer of rowband.PeerHaloExchange) are used
    in place and give the image of the dense stack, bit for bit; rows that are not dense are rejected."""
    n, h, w = 3, 128, 192
    fr, _ = synth_burst(n, h, w, seed=23)
    p = default_params()
    p.levels = 2
    sr = BurstSuperResolution(p, 0, w, h, n)
    sr.set_input(fr.to(cuda_device))
    ref = sr.next_frame().clone()
    big = torch.zeros((n, h + 40, w), dtype=fr.dtype, device=cuda_device)
    big[:, 8:8 + h] = fr.to(cuda_device)
    view = big[:, 8:8 + h]
    assert not view.is_contiguous()
    sr.set_input(view)
    assert torch.equal(sr.next_frame(), ref)
    with pytest.raises(ValueError):
        sr.set_input(big[:, :h, :w - 2])
    sr.close()


def test_pipeline_config5_large_motion_prealigned(cuda_device):
    """Config 5's regime on a small image: inter-frame motion beyond the pyramid matcher's reach (max_shift * (2^levels - 1) = 28 px at
    3 levels; the frames move up to +-24 px each, i.e. up to 48 px against each other), 3x scale.  Without the global pre-alignment the
    flows are garbage; with it the CUDA chain equals the oracle's (poses, integer shifts, consolidated shifts bit-exact; image 1e-3 /
    60 dB) and the recovered flow is within 1 px of the truth at the median ("config 5 aligned with a parity test")."""
    n, h, w = 6, 320, 448
    fr, sh = synth_burst(n, h, w, seed=515, max_shift=24.0)
    p = default_params()
    p.scale = 3
    p.levels = 3
    p.pair_span = 1
    errs = {}
    for pre in (0, 1):
        p.prealign = pre
        sr, out = _run(p, fr, cuda_device, ref_idx=0)
        e = []
        for f in range(1, n):
            fl = sr.buffer("flow", h, w * 8, f).view(np.float32).reshape(h, w, 2)[h // 4:3 * h // 4:4, w // 4:3 * w // 4:4]
            e.append(np.abs(fl + sh[f].numpy()[None, None, :]).max(-1).ravel())
        errs[pre] = float(np.median(np.concatenate(e)))
        if pre:
            exp, it = O.run_pipeline(u16(fr), p, ref_idx=0, keep=True)
            assert np.array_equal(sr.buffer("pose", n, 16).view(np.float32).reshape(n, 4), np.stack(it["poses"]))
            for k in range(sr.tile_grid()[2]):
                assert np.array_equal(sr.tile_argmin(k), it["argmin"][k]), f"pair {k}"
            for f in range(n):
                assert np.array_equal(sr.tile_shifts(f), it["frame_shift"][f])
            # the LK flows of the two sides differ by ~1e-3 px (MUFU sqrt / rcp, re-associated sums); at 3x a sample whose 3 * flow sits
            # within that of a rounding boundary reads a neighbouring raw sample: isolated flips (measured 4.7e-4 of the samples), not drift
            _compare_images(out, exp, frac=1.5e-3, db=45.0)
        sr.close()
    assert errs[1] < 1.0, errs
    assert errs[0] > 4.0, errs          # the reach really is exceeded without it

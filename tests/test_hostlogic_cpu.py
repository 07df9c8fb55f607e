"""Host-side logic that needs no GPU: the reference program's sharpen step, the bench's algorithmic-byte model, the synthetic
burst generator and the row-band parameters."""
import importlib.util
from pathlib import Path

import numpy as np
import torch

from multi_frame_super_resolution_b200 import rowband
from multi_frame_super_resolution_b200.cli import _mosaic_rggb, _sharpen
from multi_frame_super_resolution_b200.pipeline import default_params
from multi_frame_super_resolution_b200.synth import synth_burst

ROOT = Path(__file__).resolve().parent.parent


def test_sharpen_follows_sharpenImg2():
    """multi_frame_sr.cpp:90-119 restated as the loop it is: 5 c - 4 neighbours, saturated, written from the row start."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (9, 11, 3), dtype=np.uint8)
    h, w, ch = img.shape
    exp = np.zeros_like(img)
    flat = img.reshape(h, w * ch).astype(np.int32)
    for row in range(1, h - 1):
        o = 0
        for col in range(ch, (w - 1) * ch):
            v = 5 * flat[row, col] - flat[row, col - ch] - flat[row, col + ch] - flat[row - 1, col] - flat[row + 1, col]
            exp.reshape(h, w * ch)[row, o] = min(max(v, 0), 255)
            o += 1
    exp[:, 0] = 0
    exp[:, w - 1] = 0
    assert np.array_equal(_sharpen(img), exp)


def test_mosaic_maps_8bit_to_the_10bit_range():
    bgr = np.zeros((4, 6, 3), np.uint8)
    bgr[..., 2] = 255                       # pure red
    raw = _mosaic_rggb(bgr)
    assert raw.dtype == np.uint16 and raw.shape == (4, 6)
    assert (raw[0::2, 0::2] == 1023).all() and (raw[0::2, 1::2] == 64).all() and (raw[1::2, :] == 64).all()


def test_merge_algorithmic_bytes_model():
    spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.merge_bytes_per_px(8, 2) == 56.0                    # 14 N / s^2 + 16 / s^2 + 24
    assert bench.merge_bytes_per_px(15, 2) == 80.5
    assert abs(bench.merge_bytes_per_px(30, 3) - (14 * 30 + 16) / 9 - 24) < 1e-12
    assert bench.merge_bytes_per_px(8, 2, gray=True) == 34.0
    assert bench.METRIC == "output_megapixels_per_second" and bench.CFG == dict(frames=8, height=3024, width=4032, scale=2)


def test_synth_burst_is_seeded_and_in_range():
    a, sa = synth_burst(3, 64, 96, seed=5)
    b, sb = synth_burst(3, 64, 96, seed=5)
    c, _ = synth_burst(3, 64, 96, seed=6)
    assert torch.equal(a, b) and torch.equal(sa, sb) and not torch.equal(a, c)
    v = a.numpy().view(np.uint16)
    assert v.min() >= 0 and v.max() <= 1023 and a.shape == (3, 64, 96)
    assert float(sa[0].abs().max()) == 0.0 and float(sa[1:].abs().max()) <= 3.0       # frame 0 is the origin, |shift| <= max_shift


def test_band_params_fields():
    p = default_params()
    bands = rowband.plan_bands(6048, 8, p.tile_size << (p.levels - 1), 256)
    assert [b.rows for b in bands] == [768] * 7 + [672]
    bp = rowband.band_params(p, bands[3], 6048)
    assert (bp.band_global_h, bp.band_row0, bp.band_keep_row0, bp.band_keep_rows, bp.band_margin) == (6048, 3 * 768 - 256, 256, 768, rowband.default_margin(p))
    # stencils 27 rows + the aligner's reach 4 * (2^4 - 1) = 60 rows, rounded up to the LK tile height
    assert rowband.default_margin(p) == 96
    assert p.band_global_h == 0 and rowband.band_params(p, bands[0], 6048, margin=0).band_margin == 0     # the input params are not modified


def test_temporal_area_setters_validate_without_gpu():
    """multi_frame_sr.cpp:182 setTemporalAreaRadius: the setter and the sequence bookkeeping are host logic."""
    import pytest
    from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution
    sr = BurstSuperResolution(default_params(), max_width=64, max_height=64, max_frames=3)
    with pytest.raises(ValueError):
        sr.set_temporal_area_radius(-1)
    sr.set_temporal_area_radius(2)
    with pytest.raises(ValueError):
        sr.set_input(np.zeros((6, 64, 64), np.uint16))          # window 5 > max_frames 3
    sr.set_temporal_area_radius(1)
    sr.set_input(np.zeros((6, 64, 64), np.uint16))              # no device work before the first next_frame()
    assert sr._pos == 0 and sr._shape == (64, 64)
    sr.set_temporal_area_radius(None)
    with pytest.raises(RuntimeError):
        sr.next_frame()                                         # whole-burst mode again: needs set_input first

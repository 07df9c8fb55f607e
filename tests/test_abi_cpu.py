"""The C-ABI library loads on a GPU-less box, exports every symbol include/mfsr.h declares, and
fails LOUDLY (status, not a silent CPU path) when there is no device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from multi_frame_super_resolution_b200 import _lib
from multi_frame_super_resolution_b200._lib import MergeGeom, Params

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "mfsr.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mfsr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mfsr.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_no_oracle_in_product():
    """The product must not route through the oracle (or any CPU fallback)."""
    pkg = ROOT / "multi_frame_super_resolution_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        src = f.read_text()
        assert "pyoracle" not in src and "mfsr_oracle" not in src and "orc_" not in src, f
        assert "pyref" not in src and "libmfsr_ref" not in src, f


def test_default_params_layout():
    lib = _lib.load()
    p = Params()
    assert lib.mfsr_default_params(C.byref(p)) == 0
    assert p.abi_version == 1 and p.scale == 2 and list(p.cfa) == [0, 1, 1, 2]
    assert p.tile_size == 16 and p.max_shift == 4 and p.track_bits == 7 and p.pair_span == 2
    assert abs(p.track_sigma - 0.5) < 1e-7 and abs(p.weight_threshold - 0.1) < 1e-7
    assert abs(p.thresholdM - 0.8) < 1e-7 and p.mask_erode_radius == 2 and p.lk_half_window == 3
    assert p.band_global_h == 0 and p.band_keep_rows == 0 and p.band_margin == 0 and p.prealign == 0 and p.lk_texture == 0 and list(p.reserved) == [0]                      # struct size matches: the tail is still zero
    assert lib.mfsr_default_params(None) == -1


def test_error_strings_and_version():
    lib = _lib.load()
    assert lib.mfsr_abi_version() == 1
    assert lib.mfsr_error_string(0) == b"ok"
    assert b"invalid" in lib.mfsr_error_string(-1)
    assert b"device" in lib.mfsr_error_string(-4)
    assert lib.mfsr_stage_name(8) == b"merge" and lib.mfsr_stage_name(99) is None


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.mfsr_device_count() == 0
    p = Params()
    lib.mfsr_default_params(C.byref(p))
    h = C.c_void_p()
    assert lib.mfsr_create(C.byref(p), 0, 256, 256, 4, C.byref(h)) == -4 and not h


def test_argument_validation_without_gpu():
    lib = _lib.load()
    g = MergeGeom.reference(64, 64)
    # null pointers / bad geometry are rejected before any CUDA call
    assert lib.mfsr_stage_merge(None, 0, 0, None, 0, 0, None, 0, 0, None, 0, None, 0, None, 0, None, None, 0, 1, C.byref(g),
                                _lib.iarr([0, 1, 1, 2]), _lib.farr([1, 1, 1]), _lib.farr([0, 0, 0]), 0.1, 0, None) == -1
    assert lib.mfsr_stage_pyramid_down(None, 0, 8, 8, None, 0, None) == -1
    assert lib.mfsr_destroy(None) == -1
    p = Params()
    lib.mfsr_default_params(C.byref(p))
    p.track_bits = 8                                         # breaks the exact-sum contract of the integer SSD
    h = C.c_void_p()
    assert lib.mfsr_create(C.byref(p), 0, 256, 256, 4, C.byref(h)) == -1


def test_geometry_helpers():
    g = MergeGeom.reference(4032, 3024)
    assert (g.org_x, g.org_y, g.out_w, g.out_h) == (2016, 1512, 4032, 3024)
    assert (g.clamp_x0, g.clamp_x1, g.clamp_y0, g.clamp_y1) == (1008, 3023, 756, 2267)
    f = MergeGeom.full_frame(4032, 3024, 2)
    assert (f.out_w, f.out_h, f.clamp_x1, f.clamp_y1) == (8064, 6048, 4031, 3023)

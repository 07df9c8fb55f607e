"""ctypes loader for libmfsr_b200.so (the C ABI declared in include/mfsr.h).

There is NO CPU fallback: if the CUDA library is missing or a call fails the
error is raised, never papered over.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libmfsr_b200.so"

c_i64 = C.c_int64
c_f = C.c_float
c_i = C.c_int
vp = C.c_void_p
ip = C.POINTER(C.c_int)
fp = C.POINTER(C.c_float)


class MfsrError(RuntimeError):
    def __init__(self, status: int, where: str, msg: str):
        super().__init__(f"{where} failed with status {status}: {msg}")
        self.status = status


class Params(C.Structure):
    """Mirror of `mfsr_params` (include/mfsr.h)."""
    _fields_ = [
        ("abi_version", c_i), ("scale", c_i), ("full_frame", c_i), ("cfa", c_i * 4),
        ("black_level", c_f * 3), ("white_level", c_f * 3),
        ("tile_size", c_i), ("max_shift", c_i), ("levels", c_i), ("pair_span", c_i),
        ("track_bits", c_i), ("track_sigma", c_f), ("min_threshold", c_f),
        ("base_shift", c_f * 2), ("base_rotation", c_f),
        ("lk_iterations", c_i), ("lk_half_window", c_i), ("lk_min_det", c_f),
        ("Dth", c_f), ("Dtr", c_f), ("kDetail", c_f), ("kDenoise", c_f), ("kStretch", c_f), ("kShrink", c_f),
        ("tensor_box_radius", c_i),
        ("alpha", c_f), ("beta", c_f), ("thresholdM", c_f), ("mask_erode_radius", c_i),
        ("weight_threshold", c_f), ("merge_flags", c_i),
        ("band_global_h", c_i), ("band_row0", c_i), ("band_keep_row0", c_i), ("band_keep_rows", c_i), ("band_margin", c_i), ("prealign", c_i), ("lk_texture", c_i), ("reserved", c_i * 1),
    ]


def scale_rational(num: int, den: int) -> int:
    """include/mfsr.h MFSR_SCALE_RATIONAL: num / den as a scale value (scale_rational(3, 2) = 1.5x)."""
    return (den << 16) | num


class MergeGeom(C.Structure):
    """Mirror of `mfsr_merge_geom`."""
    _fields_ = [("raw_w", c_i), ("raw_h", c_i), ("scale", c_i), ("out_w", c_i), ("out_h", c_i),
                ("org_x", c_i), ("org_y", c_i), ("clamp_x0", c_i), ("clamp_x1", c_i),
                ("clamp_y0", c_i), ("clamp_y1", c_i)]

    @classmethod
    def reference(cls, raw_w: int, raw_h: int) -> "MergeGeom":
        """Geometry hard-coded in accumulateImagesSuperRes (DeBayerKernels.cu:398-423)."""
        return cls(raw_w, raw_h, 2, raw_w, raw_h, raw_w // 2, raw_h // 2,
                   raw_w // 4, raw_w // 2 - 1 + raw_w // 4, raw_h // 4, raw_h // 2 - 1 + raw_h // 4)

    @classmethod
    def full_frame(cls, raw_w: int, raw_h: int, scale: int) -> "MergeGeom":
        num, den = scale & 0xffff, ((scale >> 16) & 0xffff) or 1      # scale_rational(num, den)
        return cls(raw_w, raw_h, scale, raw_w * num // den, raw_h * num // den, 0, 0, 0, raw_w - 1, 0, raw_h - 1)

    @classmethod
    def one_to_one(cls, raw_w: int, raw_h: int) -> "MergeGeom":
        """accumulateImages (DeBayerKernels.cu:290): scale 1."""
        return cls.full_frame(raw_w, raw_h, 1)


# name -> (restype, argtypes); every symbol include/mfsr.h declares
SIGNATURES = {
    "mfsr_abi_version": (c_i, []),
    "mfsr_error_string": (C.c_char_p, [c_i]),
    "mfsr_default_params": (c_i, [C.POINTER(Params)]),
    "mfsr_device_count": (c_i, []),
    "mfsr_stage_subsample3": (c_i, [vp, c_i64, vp, c_i64, c_f, c_i, c_i, ip, vp]),
    "mfsr_stage_demosaic": (c_i, [vp, c_i64, vp, c_i64, c_i, c_i, ip, fp, fp, vp]),
    "mfsr_stage_tracking_image": (c_i, [vp, c_i64, vp, c_i64, vp, c_i64, c_i, c_i, ip, fp, fp, c_f, c_i, vp]),
    "mfsr_stage_pyramid_down": (c_i, [vp, c_i64, c_i, c_i, vp, c_i64, vp]),
    "mfsr_stage_tile_align": (c_i, [vp, vp, c_i64, c_i, c_i, vp, c_i64, vp, c_i64, vp, vp,
                                    c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, vp]),
    "mfsr_stage_prealign_search": (c_i, [vp, vp, c_i64, c_i, c_i, vp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, vp, vp]),
    "mfsr_stage_upsample_shifts": (c_i, [vp, c_i64, vp, c_i64, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, vp]),
    "mfsr_stage_consolidate_shifts": (c_i, [vp, ip, ip, c_i, c_i, c_i, c_i, c_i, vp, vp, vp, vp]),
    "mfsr_stage_flow_from_tiles": (c_i, [vp, c_i64, c_i, c_i, c_i, vp, c_i64, c_i, c_i, c_f, c_f, c_f, vp]),
    "mfsr_stage_lk_iteration": (c_i, [vp, vp, c_i64, vp, vp, c_i64, c_i, c_i, c_i, c_f, vp]),
    "mfsr_stage_lk_iteration_tex": (c_i, [vp, vp, c_i64, vp, vp, c_i64, c_i, c_i, c_i, c_f, vp]),
    "mfsr_stage_kernel_params": (c_i, [vp, c_i64, vp, c_i64, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f, vp]),
    "mfsr_stage_robustness": (c_i, [vp, vp, c_i64, vp, c_i64, vp, c_i64, vp, c_i, c_i, c_f, c_f, c_f, c_i, vp]),
    "mfsr_stage_merge": (c_i, [vp, c_i64, c_i64, vp, c_i64, c_i64, vp, c_i64, c_i64, vp, c_i64, vp, c_i64,
                               vp, c_i64, vp, vp, c_i64, c_i, C.POINTER(MergeGeom), ip, fp, fp, c_f, c_i, vp]),
    "mfsr_stage_fallback_upsample": (c_i, [vp, c_i64, c_i, c_i, vp, c_i64, C.POINTER(MergeGeom), vp]),
    "mfsr_create": (c_i, [C.POINTER(Params), c_i, c_i, c_i, c_i, C.POINTER(vp)]),
    "mfsr_destroy": (c_i, [vp]),
    "mfsr_output_size": (c_i, [vp, c_i, c_i, ip, ip]),
    "mfsr_workspace_bytes": (c_i64, [vp]),
    "mfsr_set_frames": (c_i, [vp, C.POINTER(vp), c_i, c_i, c_i, c_i64, c_i, c_i, c_i]),
    "mfsr_run": (c_i, [vp, vp, c_i64, c_i]),
    "mfsr_run_async": (c_i, [vp, vp, c_i64, c_i]),
    "mfsr_run_format": (c_i, [vp, vp, c_i64, c_i, c_i, c_i]),
    "mfsr_synchronize": (c_i, [vp]),
    "mfsr_stream": (vp, [vp]),
    "mfsr_get_tile_grid": (c_i, [vp, ip, ip, ip]),
    "mfsr_get_tile_argmin": (c_i, [vp, c_i, vp]),
    "mfsr_get_tile_shifts": (c_i, [vp, c_i, vp]),
    "mfsr_get_stage_ms": (c_i, [vp, fp, c_i]),
    "mfsr_stage_name": (C.c_char_p, [c_i]),
    "mfsr_get_buffer": (c_i, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(c_i64), C.POINTER(c_i64)]),
    "mfsr_last_launch_count": (c_i, [vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m multi_frame_super_resolution_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mfsr_abi_version() != 1:
        raise ImportError("libmfsr_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, where: str) -> None:
    if status != 0:
        raise MfsrError(status, where, load().mfsr_error_string(status).decode())


def iarr(vals):
    return (c_i * len(vals))(*[int(v) for v in vals])


def farr(vals):
    return (c_f * len(vals))(*[float(v) for v in vals])

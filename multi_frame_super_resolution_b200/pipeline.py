"""Host-side mirror of the reference's pull-style burst API over the C ABI.

The reference's only runnable burst program drives `cv::superres::SuperResolution`
(finalProject/Project/multi_frame_sr.cpp:165-194): create, setScale / setIterations /
setTemporalAreaRadius / setInput, then pull results with nextFrame().  `BurstSuperResolution`
keeps that shape (set_scale, set_iterations, set_temporal_area_radius, set_input, next_frame) on
top of mfsr_create / mfsr_set_frames / mfsr_run.  Two pull modes:

* whole burst (default): set_input(frames) is one burst, next_frame() merges all of it onto
  frame `ref_idx` (BASELINE configs);
* temporal area (multi_frame_sr.cpp:182 setTemporalAreaRadius(1), :185-194 the nextFrame loop):
  set_input(frames) is a frame SEQUENCE, each next_frame() returns the super-resolved frame i
  merged from frames [i - r, i + r] (clipped at the ends of the sequence) and advances i;
  None once the sequence is exhausted (the reference's empty result, :191).

All arithmetic happens in libmfsr_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Params, check

FMT_BAYER_U16 = 0
FMT_GRAY_U16 = 1
OUT_F32, OUT_F16, OUT_U8 = 0, 1, 2

_cudart = None


def _rt():
    global _cudart
    if _cudart is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart = C.CDLL(name)
                break
            except OSError:
                continue
        if _cudart is None:
            raise ImportError("libcudart not found")
        _cudart.cudaMemcpy2D.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        _cudart.cudaMemcpy2D.restype = C.c_int
    return _cudart


def default_params() -> Params:
    p = Params()
    check(_lib.load().mfsr_default_params(C.byref(p)), "mfsr_default_params")
    return p


def measured_pairs(n_frames: int, pair_span: int, prealign: int = 0, ref_idx: int = 0):
    """The (from, to) frame pairs the aligner measures: 0 < to - from <= pair_span, with the global pre-alignment also every
    frame against the reference frame (csrc/pipeline.cu: build_pairs)."""
    return [(i, j) for i in range(n_frames) for j in range(i + 1, n_frames)
            if j - i <= pair_span or (prealign and (i == ref_idx or j == ref_idx))]


class BurstSuperResolution:
    """One handle = one GPU, one stream, one workspace (re-usable for many bursts)."""

    def __init__(self, params: Optional[Params] = None, device: int = 0, max_width: int = 4032, max_height: int = 3024,
                 max_frames: int = 8):
        self._lib = _lib.load()
        self.params = params if params is not None else default_params()
        self.device = device
        self._max = (max_width, max_height, max_frames)
        self._h = C.c_void_p()
        self._shape = None
        self._n = 0
        self._keep = None
        self._keep_ev = None
        self._retired = []
        self._radius = None          # None: whole-burst mode
        self._seq = None             # temporal-area mode: (frames, fmt), position
        self._pos = 0

    # -- cv::superres-style setters (multi_frame_sr.cpp:179-182); take effect at the next (re)create
    def set_scale(self, scale: int):
        self._destroy()
        self.params.scale = int(scale)

    def set_iterations(self, iterations: int):
        """Reference: BTV-L1 iterations; here: Lucas-Kanade refinement sweeps."""
        self._destroy()
        self.params.lk_iterations = int(iterations)

    def set_temporal_area_radius(self, radius: Optional[int]):
        """multi_frame_sr.cpp:182.  radius >= 0 switches to the sliding mode (one output per input frame, window 2 r + 1);
        None goes back to whole-burst mode.  Takes effect at the next set_input()."""
        if radius is not None and radius < 0:
            raise ValueError("temporal area radius must be >= 0 (None = whole burst)")
        self._radius = None if radius is None else int(radius)
        self._seq = None
        self._shape = None

    def reset(self):
        """Rewind the frame sequence (FrameSource::reset, multi_frame_sr.cpp:46-49)."""
        self._pos = 0

    def _ensure(self):
        if not self._h:
            h = C.c_void_p()
            check(self._lib.mfsr_create(C.byref(self.params), self.device, *self._max, C.byref(h)), "mfsr_create")
            self._h = h

    def _ext(self):
        return torch.cuda.ExternalStream(self.stream, device=torch.device("cuda", self.device))

    def _destroy(self):
        if self._h:
            self._lib.mfsr_destroy(self._h)      # synchronises the handle's stream
            self._h = C.c_void_p()
            self._retired = []
            self._keep_ev = None

    def close(self):
        self._destroy()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    @property
    def workspace_bytes(self) -> int:
        self._ensure()
        return int(self._lib.mfsr_workspace_bytes(self._h))

    @property
    def stream(self) -> int:
        self._ensure()
        return int(self._lib.mfsr_stream(self._h) or 0)

    def output_size(self, width: int, height: int):
        self._ensure()
        ow, oh = C.c_int(), C.c_int()
        check(self._lib.mfsr_output_size(self._h, width, height, C.byref(ow), C.byref(oh)), "mfsr_output_size")
        return ow.value, oh.value

    # -- setInput: frames is a CUDA tensor [N,H,W] (uint16/int16) or a host numpy/pinned tensor of the same shape
    def set_input(self, frames, ref_idx: int = 0, fmt: int = FMT_BAYER_U16):
        if self._radius is not None:
            if 2 * self._radius + 1 > self._max[2]:
                raise ValueError(f"temporal area 2*{self._radius}+1 exceeds max_frames={self._max[2]}")
            if len(frames.shape) != 3:
                raise ValueError("frames must be [N,H,W]")
            self._seq, self._pos = (frames, fmt), 0
            self._shape = tuple(frames.shape[1:])
            return
        self._set_window(frames, ref_idx, fmt)

    def _set_window(self, frames, ref_idx: int, fmt: int):
        self._ensure()
        if isinstance(frames, np.ndarray):
            if frames.dtype != np.uint16 or frames.ndim != 3 or not frames.flags.c_contiguous:
                raise ValueError("host frames must be a C-contiguous uint16 array [N,H,W]")
            n, h, w = frames.shape
            base, on_host, fstride = frames.ctypes.data, 1, h * w * 2
        else:
            # every frame dense ([H,W] contiguous); the frames themselves may be further apart than H*W (a view into a larger stack)
            if frames.dim() != 3 or frames.element_size() != 2 or frames.stride(2) != 1 or frames.stride(1) != frames.shape[2] or \
                    (frames.shape[0] > 1 and frames.stride(0) < frames.shape[1] * frames.shape[2]):
                raise ValueError("frames must be a 16-bit tensor [N,H,W] with dense frames")
            n, h, w = frames.shape
            base, on_host = frames.data_ptr(), 0 if frames.is_cuda else 1
            fstride = frames.stride(0) * 2 if n > 1 else h * w * 2
        if not on_host:
            # the frames were produced on torch's current stream: order the handle's stream after it
            self._ext().wait_stream(torch.cuda.current_stream(self.device))
        ptrs = (C.c_void_p * n)(*[base + i * fstride for i in range(n)])
        old = self._keep
        if isinstance(old, torch.Tensor) and old.is_cuda:
            # device frames are read in place until the end of the run: their memory must not go back to the caching allocator
            # before the handle's stream has passed that work.  (Not record_stream(): the allocator would later record an event
            # on the handle's stream, which may be destroyed by then.)
            self._retired = [(t, e) for (t, e) in self._retired if not e.query()]
            if old is not frames:
                ev = torch.cuda.Event()
                ev.record(self._ext())
                self._retired.append((old, ev))
        elif old is not None and self._keep_ev is not None:
            self._keep_ev.synchronize()      # host source of the async H2D copies: wait for the copies (not for the run)
        check(self._lib.mfsr_set_frames(self._h, ptrs, n, w, h, w * 2, fmt, ref_idx, on_host), "mfsr_set_frames")
        self._shape = (h, w)
        self._n = n
        self._keep = frames          # the async H2D copies read it until the stream reaches them
        self._keep_ev = None
        if on_host:
            self._keep_ev = torch.cuda.Event()
            self._keep_ev.record(self._ext())

    # -- nextFrame: run the whole chain, return the float3 image
    def next_frame(self, out=None, host: bool = False, sync: bool = True, dtype: torch.dtype = torch.float32):
        """host=True: `out` is host memory; sync=False only enqueues the D2H copy (pinned `out`), call synchronize().
        dtype: torch.float32 (mfsr_run), torch.float16 or torch.uint8 (mfsr_run_format: half3 / 8-bit image)."""
        if self._shape is None:
            raise RuntimeError("set_input() first")
        if self._radius is not None:
            frames, fmt = self._seq
            n, i, r = frames.shape[0], self._pos, self._radius
            if i >= n:
                return None
            lo, hi = max(0, i - r), min(n, i + r + 1)
            self._set_window(frames[lo:hi], i - lo, fmt)
            self._pos = i + 1
        ow, oh = self.output_size(self._shape[1], self._shape[0])
        if out is not None and isinstance(out, torch.Tensor) and out.is_cuda:
            # a consumer on torch's current stream may still be reading the previous result from `out`
            self._ext().wait_stream(torch.cuda.current_stream(self.device))
        if dtype != torch.float32:
            fmt = {torch.float16: OUT_F16, torch.uint8: OUT_U8}[dtype]
            if out is None:
                out = torch.empty((oh, ow, 3), dtype=dtype, pin_memory=True) if host else torch.empty((oh, ow, 3), dtype=dtype, device=f"cuda:{self.device}")
            if out.dtype != dtype or tuple(out.shape) != (oh, ow, 3) or not out.is_contiguous():
                raise ValueError("out must be a contiguous [out_h, out_w, 3] tensor of the requested dtype")
            check(self._lib.mfsr_run_format(self._h, C.c_void_p(out.data_ptr()), ow * 3 * out.element_size(), 1 if host else 0, fmt, 0 if (sync and host) else 1),
                  "mfsr_run_format")
            if host:
                self._keep_out = out
            else:
                torch.cuda.current_stream(self.device).wait_stream(self._ext())
            return out
        if host:
            if out is None:
                out = torch.empty((oh, ow, 3), dtype=torch.float32, pin_memory=True)
            ptr = out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data
            fn = self._lib.mfsr_run if sync else self._lib.mfsr_run_async
            check(fn(self._h, C.c_void_p(ptr), ow * 12, 1), "mfsr_run")
            self._keep_out = out
            return out
        if out is None:
            out = torch.empty((oh, ow, 3), dtype=torch.float32, device=f"cuda:{self.device}")
        check(self._lib.mfsr_run(self._h, C.c_void_p(out.data_ptr()), ow * 12, 0), "mfsr_run")
        # asynchronous, but ordered: whoever consumes `out` on torch's current stream sees the finished image
        torch.cuda.current_stream(self.device).wait_stream(self._ext())
        return out

    def synchronize(self):
        if self._h:
            check(self._lib.mfsr_synchronize(self._h), "mfsr_synchronize")

    # -- introspection used by tests / bench
    def stage_ms(self) -> dict:
        buf = (C.c_float * 16)()
        n = self._lib.mfsr_get_stage_ms(self._h, buf, 16)
        if n < 0:
            check(n, "mfsr_get_stage_ms")
        return {self._lib.mfsr_stage_name(i).decode(): float(buf[i]) for i in range(n)}

    def launch_count(self) -> int:
        return int(self._lib.mfsr_last_launch_count(self._h))

    def tile_grid(self):
        tx, ty, m = C.c_int(), C.c_int(), C.c_int()
        check(self._lib.mfsr_get_tile_grid(self._h, C.byref(tx), C.byref(ty), C.byref(m)), "mfsr_get_tile_grid")
        return tx.value, ty.value, m.value

    def tile_argmin(self, pair: int) -> np.ndarray:
        tx, ty, _ = self.tile_grid()
        a = np.empty((ty, tx, 2), dtype=np.int32)
        check(self._lib.mfsr_get_tile_argmin(self._h, pair, C.c_void_p(a.ctypes.data)), "mfsr_get_tile_argmin")
        return a

    def tile_shifts(self, frame: int) -> np.ndarray:
        tx, ty, _ = self.tile_grid()
        a = np.empty((ty, tx, 2), dtype=np.float32)
        check(self._lib.mfsr_get_tile_shifts(self._h, frame, C.c_void_p(a.ctypes.data)), "mfsr_get_tile_shifts")
        return a

    def buffer(self, name: str, rows: int, row_bytes: int, frame: int = 0) -> np.ndarray:
        """Copy an intermediate buffer of the last run to the host as raw bytes [rows, row_bytes]."""
        ptr, pitch, fs = C.c_void_p(), C.c_int64(), C.c_int64()
        check(self._lib.mfsr_get_buffer(self._h, name.encode(), C.byref(ptr), C.byref(pitch), C.byref(fs)), "mfsr_get_buffer")
        self.synchronize()
        out = np.empty((rows, row_bytes), dtype=np.uint8)
        rc = _rt().cudaMemcpy2D(C.c_void_p(out.ctypes.data), row_bytes, C.c_void_p(ptr.value + fs.value * frame), pitch.value,
                                row_bytes, rows, 2)
        if rc != 0:
            raise _lib.MfsrError(rc, "cudaMemcpy2D", "copy of intermediate buffer failed")
        return out

#!/usr/bin/env python
"""`multi_frame_sr`-compatible command line for the B200 burst path.

Mirrors finalProject/Project/multi_frame_sr.cpp:122-209:

    python -m multi_frame_super_resolution_b200.cli optFlowName inputName iterations

* inputName city / car / iso selects `img_%06d.png` (5 frames; the repo ships 000000..000004, the program asks
  for 000001..000005 — both numberings are tried), `car/%d.jpg` (4), `iso/%06d.png` (4) in the current directory
  (:151-163); any other value is taken as a printf-style pattern with `--frames N`.
* optFlowName is accepted for compatibility (farneback / tvl1 / brox / pyrlk select OpenCV flows in the reference,
  :61-87); this path always uses its own tile aligner + Lucas-Kanade refinement, `iterations` sets the LK sweeps.
* scale 2, the burst is processed `num_times = 10` times and the last `real_times = 5` repetitions are timed (:146-149);
  prints "<t> sec" and "<fps> FPS" (:204-206), writes <input>_<flow>_sr_result.png and the Laplacian-sharpened
  <input>_<flow>_sr2_result.png (:90-119, :207-209), plus one JSON line with MP/s.
  (`host/multi_frame_sr_b200.cpp` is the same program in C++ over the C ABI, with Netpbm files.)
* `--radius R` selects the reference's temporal-area pull mode (:182 setTemporalAreaRadius(1), :185-194): the frame set repeated
  num_times is one sequence, every nextFrame() is the super-resolved frame i merged from frames [i - R, i + R], and the frames
  from start_i = (num_times - real_times) * num_images on are timed.  Without it the frame set is one burst merged onto frame 0.

8-bit colour frames are mosaiced to RGGB and mapped to the 10-bit range of the default parameters
(raw = round(v8 * 959 / 255) + 64), exactly like tests/golden/make_bundled_fixture.py.
All arithmetic runs in libmfsr_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

import json
import sys
import time

import numpy as np


def _mosaic_rggb(bgr: np.ndarray) -> np.ndarray:
    rgb = bgr[..., ::-1].astype(np.float64)
    h, w = rgb.shape[:2]
    h, w = h & ~1, w & ~1
    raw = np.empty((h, w), np.float64)
    raw[0::2, 0::2] = rgb[0:h:2, 0:w:2, 0]
    raw[0::2, 1::2] = rgb[0:h:2, 1:w:2, 1]
    raw[1::2, 0::2] = rgb[1:h:2, 0:w:2, 1]
    raw[1::2, 1::2] = rgb[1:h:2, 1:w:2, 2]
    return (np.round(raw * 959.0 / 255.0) + 64).astype(np.uint16)


def _sharpen(img8: np.ndarray) -> np.ndarray:
    """sharpenImg2 (multi_frame_sr.cpp:90-119): 5 c - left - right - up - down, saturated.  The reference advances its output
    pointer from the START of each row while reading from column 1, so the result sits one pixel to the left; border rows and
    columns are 0 (the one column the reference leaves uninitialised is 0 here).  Same arithmetic as host/multi_frame_sr_b200.cpp."""
    a = img8.astype(np.int32)
    h, w = a.shape[:2]
    out = np.zeros_like(img8)
    v = 5 * a[1:-1, 1:-1] - a[1:-1, :-2] - a[1:-1, 2:] - a[:-2, 1:-1] - a[2:, 1:-1]
    out[1:-1, 0:w - 2] = np.clip(v, 0, 255).astype(np.uint8)
    out[:, 0] = 0
    out[:, w - 1] = 0
    return out


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    n_override = None
    radius = None
    if "--radius" in argv:
        i = argv.index("--radius")
        radius = int(argv[i + 1])
        del argv[i:i + 2]
    if "--frames" in argv:
        i = argv.index("--frames")
        n_override = int(argv[i + 1])
        del argv[i:i + 2]
    if len(argv) == 0:
        flow, name, iterations = "farneback", "city", 10
    elif len(argv) == 3:
        flow, name, iterations = argv[0], argv[1], max(1, int(argv[2]))
    else:
        print("./multi_frame_sr optFlowName inputName iterations")
        print("\toptFlowName: farneback, tvl1, brox, pyrlk")
        print("\tinputName: city, car, iso")
        print("\titerations: integer, 1, 10, etc.")
        return -1
    patterns = {"city": (5, "img_%06d.png"), "car": (4, "car/%d.jpg"), "iso": (4, "iso/%06d.png")}
    if name in patterns:
        n, fmt = patterns[name]
    elif n_override:
        n, fmt = n_override, name
    else:
        print("wrong input")
        return -1
    import cv2
    import torch
    from .pipeline import BurstSuperResolution, default_params

    frames = []
    for base in (1, 0):                       # the program asks for 1-based names, the repo ships 0-based city frames
        frames = []
        for i in range(n):
            path = fmt % (i + base)
            img = cv2.imread(path, cv2.IMREAD_COLOR)
            if img is None:
                frames = []
                break
            print(f"{path}, [{img.shape[1]} x {img.shape[0]}]")
            frames.append(_mosaic_rggb(img))
        if frames:
            break
    if not frames:
        print(f"cannot read {fmt % 1}")
        return -1
    raw = np.stack(frames)
    n, h, w = raw.shape
    p = default_params()
    p.scale = 2
    p.lk_iterations = iterations
    p.merge_flags = 1                         # GammasRGB: the result is written as an 8-bit image
    while p.levels > 1 and min(h, w) >> (p.levels - 1) < 2 * p.max_shift + p.tile_size:
        p.levels -= 1
    num_times, real_times = 10, 5
    sr = BurstSuperResolution(p, device=0, max_width=w, max_height=h, max_frames=n if radius is None else 2 * radius + 1)
    dev = torch.from_numpy(raw.view(np.int16)).cuda()
    out = None
    if radius is None:
        for t in range(num_times):
            if t == num_times - real_times:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            sr.set_input(dev)
            out = sr.next_frame()
    else:
        sr.set_temporal_area_radius(radius)
        sr.set_input(dev.repeat(num_times, 1, 1))             # :168-176: the frame set, num_times over
        start_i = (num_times - real_times) * n
        for i in range(n * num_times):
            if i == start_i:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            res = sr.next_frame()
            if res is None:                                   # :191 result.empty()
                break
            out = res
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    print(f"{sec:.6g} sec")
    print(f"{real_times * n / sec:.6g} FPS")
    img = out.cpu().numpy()
    img8 = np.clip(np.nan_to_num(img) * 255.0 + 0.5, 0, 255).astype(np.uint8)[..., ::-1]
    cv2.imwrite(f"{name if name in patterns else 'burst'}_{flow}_sr_result.png", img8)
    cv2.imwrite(f"{name if name in patterns else 'burst'}_{flow}_sr2_result.png", _sharpen(img8))
    print(json.dumps({"output_megapixels_per_second": round(real_times * img.shape[0] * img.shape[1] / 1e6 / sec, 2),
                      "frames": n, "raw": [w, h], "scale": 2, "lk_iterations": iterations}))
    sr.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

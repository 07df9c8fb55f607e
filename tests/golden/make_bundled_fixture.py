#!/usr/bin/env python
"""Builds tests/golden/bundled_burst_rggb.npz from the reference's bundled 5-frame burst
(/root/reference/test_opencv/img_00000{0..4}.png, 512x256 BGR u8; generator documented in
test_opencv/main.cpp:1877-1913).  Runs only in the authoring container (needs /root/reference and cv2).

Each frame is mosaiced to RGGB and mapped to the 10-bit range the pipeline defaults assume:
raw = round(v8 * 959 / 255) + 64 (black 64, white 1023)."""
import sys
from pathlib import Path

import cv2
import numpy as np

REF = Path("/root/reference/test_opencv")
OUT = Path(__file__).resolve().parent / "bundled_burst_rggb.npz"


def main():
    frames = []
    for i in range(5):
        bgr = cv2.imread(str(REF / f"img_{i:06d}.png"), cv2.IMREAD_COLOR)
        assert bgr is not None and bgr.shape == (256, 512, 3), bgr
        rgb = bgr[..., ::-1].astype(np.float64)
        raw = np.empty((256, 512), np.float64)
        raw[0::2, 0::2] = rgb[0::2, 0::2, 0]
        raw[0::2, 1::2] = rgb[0::2, 1::2, 1]
        raw[1::2, 0::2] = rgb[1::2, 0::2, 1]
        raw[1::2, 1::2] = rgb[1::2, 1::2, 2]
        frames.append((np.round(raw * 959.0 / 255.0) + 64).astype(np.uint16))
    np.savez_compressed(OUT, frames=np.stack(frames))
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    sys.exit(main())

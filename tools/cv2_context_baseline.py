"""Context baseline, NOT the same algorithm: the closest stand-in this image has for the reference's OpenCV path
(multi_frame_sr.cpp drives cv::superres BTVL1 with a Farneback / TV-L1 flow; the superres / optflow contrib modules are not in this
cv2 build): per frame cv2.calcOpticalFlowFarneback to the reference frame, cv2.remap, mean over the burst, 2x cubic resize.
Timed on all host cores on the same bounded sample as bench.py's CPU arm (8 frames of 1024x768).  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np


def main():
    try:
        import cv2
    except ImportError:
        print(json.dumps({"unavailable": "cv2 not importable"})); return 0
    from multi_frame_super_resolution_b200.synth import synth_burst
    n, h, w = 8, 768, 1024
    cv2.setNumThreads(os.cpu_count() or 1)
    fr, _ = synth_burst(n, h, w, seed=1234, bayer=False)
    gray = [np.clip((f.numpy().view(np.uint16).astype(np.float32) - 64.0) / 959.0, 0, 1) for f in fr]
    g8 = [(g * 255).astype(np.uint8) for g in gray]
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))

    def run():
        acc = gray[0].copy()
        for f in range(1, n):
            flow = cv2.calcOpticalFlowFarneback(g8[0], g8[f], None, 0.5, 4, 15, 3, 5, 1.2, 0)
            acc += cv2.remap(gray[f], xs + flow[..., 0], ys + flow[..., 1], cv2.INTER_LINEAR)
        return cv2.resize(acc / n, (2 * w, 2 * h), interpolation=cv2.INTER_CUBIC)

    run()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        out = run()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"baseline": "cv2 Farneback + remap + mean + 2x cubic resize (context only: not the reference's algorithm)",
                      "sample": f"{n} gray frames of {w}x{h}", "threads": cv2.getNumThreads(), "ms": round(dt * 1e3, 1),
                      "output_megapixels_per_second": round(4 * w * h / 1e6 / dt, 2), "cv2": cv2.__version__}))
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Generates tests/golden/ref_golden.npz by running the REFERENCE's OWN kernels (compiled unmodified
into oracle/_ref/libmfsr_ref.so, see oracle/Makefile + oracle/ref_driver.cu) on a B200.

    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/ref_golden.npz'
    cp gpurun_out/ref_golden.npz tests/golden/ref_golden.npz

Inputs are seeded numpy arrays; both inputs and reference outputs are stored, so the CPU test
(tests/test_oracle_golden.py) needs neither a GPU nor /root/reference.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyref  # noqa: E402

WHITE, BLACK = [959.0, 959.0, 959.0], [64.0, 64.0, 64.0]
SCALE = [float(np.float32(1.0) / np.float32(959.0))] * 3
RGGB = [0, 1, 1, 2]


def main(out_path):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(20261018)
    G = {}

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def smooth(h, w, amp, k=3):
        ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
        acc = np.zeros((h, w), np.float32)
        for _ in range(k):
            fx, fy, ph = rng.uniform(-0.08, 0.08), rng.uniform(-0.08, 0.08), rng.uniform(0, 6.28)
            acc += np.sin(2 * np.pi * (fx * xs + fy * ys) + ph).astype(np.float32)
        return (amp * acc / k).astype(np.float32)

    # ---- front end
    h, w = 40, 56
    raw = rng.integers(64, 1023, size=(h, w), dtype=np.uint16)
    raw = (raw * 0.3 + 0.7 * (300 + 250 * (smooth(h, w, 1.0) + 1))).astype(np.uint16)
    G["fe_raw"] = raw
    d = T(raw.view(np.int16))
    G["fe_rgb"] = pyref.debayer(d, BLACK, SCALE, RGGB).cpu().numpy()
    G["fe_half"] = pyref.subsample3(d, 1023.0, RGGB).cpu().numpy()
    G["fe_rgb_bggr"] = pyref.debayer(d, BLACK, SCALE, [2, 1, 1, 0]).cpu().numpy()

    # ---- tile alignment (7-bit integer valued images, T=16, M=4)
    h, w = 72, 88
    base = rng.integers(0, 128, size=(h + 16, w + 16)).astype(np.float32)
    k = np.ones((3, 3), np.float32) / 9
    base = np.stack([np.roll(np.roll(base, i, 0), j, 1) for i in (-1, 0, 1) for j in (-1, 0, 1)]).mean(0)
    a = np.floor(base[8:8 + h, 8:8 + w]).astype(np.uint8)
    b = np.floor(base[6:6 + h, 11:11 + w]).astype(np.uint8)          # shifted by (+3, -2)
    pre = rng.uniform(-2, 2, size=(4, 5, 2)).astype(np.float32)
    G["ta_ref"], G["ta_mov"], G["ta_pre"] = a, b, pre
    for name, p in (("nopre", None), ("pre", T(pre))):
        for fft in (0, 1):
            coord, ssd = pyref.tile_align(T(a.astype(np.float32)), T(b.astype(np.float32)), p, 16, 4, use_fft=bool(fft))
            G[f"ta_coord_{name}_fft{fft}"] = coord.cpu().numpy()
            G[f"ta_ssd_{name}_fft{fft}"] = ssd.cpu().numpy()
    G["up_in"] = rng.uniform(-4, 4, size=(5, 7, 2)).astype(np.float32)
    G["up_out"] = pyref.upsample_shifts(T(G["up_in"]), 4, 2, 15, 11, 16, 16).cpu().numpy()

    # ---- flow
    tiles = rng.uniform(-3, 3, size=(3, 4, 2)).astype(np.float32)
    G["ff_tiles"] = tiles
    G["ff_flow"] = pyref.flow_from_tiles(T(tiles), 16, 72, 56).cpu().numpy()
    h, w = 48, 64
    img_a = (0.5 + smooth(h, w, 0.3, 4)).astype(np.float32)
    img_b = np.roll(img_a, (1, -1), (0, 1)) + rng.normal(0, 0.003, (h, w)).astype(np.float32)
    flow = np.stack([smooth(h, w, 1.5), smooth(h, w, 1.5)], -1).astype(np.float32)
    G["of_a"], G["of_b"], G["of_flow"] = img_a, img_b, flow
    warped = pyref.warp(T(flow), T(img_b))
    G["of_warped"] = warped.cpu().numpy()
    ix, iy, iz = pyref.derivatives(warped, T(img_a))
    G["of_ix"], G["of_iy"], G["of_iz"] = ix.cpu().numpy(), iy.cpu().numpy(), iz.cpu().numpy()
    G["of_lk"] = pyref.lucas_kanade(T(flow), ix, iy, iz, 3, 1e-3).cpu().numpy()
    ix2, iy2 = pyref.derivatives2(T(img_a))
    G["kp_ix"], G["kp_iy"] = ix2.cpu().numpy(), iy2.cpu().numpy()
    t3 = pyref.structure_tensor(ix2, iy2)
    G["kp_tensor"] = t3.cpu().numpy()
    G["kp_kernel"] = pyref.kernel_param(t3, 0.005, 0.012, 0.3, 4.0, 4.0, 2.0).cpu().numpy()

    # ---- robustness (half-res 24 x 32, flow 48 x 64)
    ref3 = np.stack([0.4 + smooth(24, 32, 0.2), 0.5 + smooth(24, 32, 0.2), 0.3 + smooth(24, 32, 0.2)], -1).astype(np.float32)
    mov3 = (ref3 + rng.normal(0, 0.02, ref3.shape)).astype(np.float32)
    mov3[8:14, 10:20] += 0.3
    G["rb_ref"], G["rb_mov"] = ref3, mov3
    G["rb_mask"] = pyref.robustness_mask(T(ref3), T(mov3), T(flow), 1e-3, 1e-5, 0.8).cpu().numpy()

    # ---- merge (reference geometry; 2x) and 1x
    n, h, w = 3, 48, 64
    mraw = rng.integers(64, 1023, size=(n, h, w), dtype=np.uint16)
    mask = rng.uniform(0, 1, size=(n, h // 2, w // 2, 4)).astype(np.float32)
    mask[0, 4:8, 6:12, :3] = 0
    mflow = np.stack([np.stack([smooth(h, w, 2.5) + rng.uniform(-2, 2), smooth(h, w, 2.5) + rng.uniform(-2, 2)], -1) for _ in range(n)]).astype(np.float32)
    k1, k2, ang = 0.3 + np.abs(smooth(h, w, 0.3)), 0.1 + np.abs(smooth(h, w, 0.1)), smooth(h, w, 3.0)
    c, s = np.cos(ang), np.sin(ang)
    b11, b22, b12 = k1 * c * c + k2 * s * s, k1 * s * s + k2 * c * c, (k1 - k2) * c * s
    det = b11 * b22 - b12 * b12
    kern = np.stack([b22 / det, b11 / det, -b12 / det, np.zeros_like(det)], -1).astype(np.float32)
    fb = rng.uniform(0, 1, size=(h, w, 3)).astype(np.float32)
    G["mg_raw"], G["mg_mask"], G["mg_flow"], G["mg_kernel"], G["mg_fallback"] = mraw, mask, mflow, kern, fb
    out, ssum, wsum = pyref.merge_superres(T(mraw.view(np.int16)), T(mask), T(mflow), T(kern), T(fb), WHITE, BLACK, 0.1, RGGB,
                                           gamma=True, want_accumulators=True)
    G["mg_out"], G["mg_sum"], G["mg_weight"] = out.cpu().numpy(), ssum.cpu().numpy(), wsum.cpu().numpy()
    G["mg_out_1x"] = pyref.merge_1x(T(mraw.view(np.int16)), T(mask), T(mflow), T(np.ascontiguousarray(kern[..., :3])), T(fb),
                                    WHITE, BLACK, 0.1, RGGB).cpu().numpy()

    # ---- texture unit probe: ramp texture, quarter positions of the merge + arbitrary positions
    tw = 504
    tex = np.arange(tw, dtype=np.float32)
    x = np.arange(1, 2 * tw - 1, dtype=np.float32)
    xn_quarter = ((x + 0.5) / 2.0 / tw).astype(np.float32)
    xn_rand = rng.uniform(0.01, 0.99, size=4096).astype(np.float32)
    G["tx_xn_quarter"], G["tx_xn_rand"] = xn_quarter, xn_rand
    G["tx_out_quarter"] = pyref.texture_probe(T(tex), T(xn_quarter)).cpu().numpy()
    G["tx_out_rand"] = pyref.texture_probe(T(tex), T(xn_rand)).cpu().numpy()
    G["tx_width"] = np.array([tw], np.int32)

    Path(out_path).parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out_path, **G)
    print("wrote", out_path, sum(v.nbytes for v in G.values()), "bytes raw")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_golden.npz")

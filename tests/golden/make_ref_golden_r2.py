#!/usr/bin/env python
"""Generates tests/golden/ref_golden_r2.npz: shift consolidation run through the REFERENCE's OWN kernels
(copyShiftMatrix / setPointers / transposeShifts / checkForOutliers / getOptimalShifts, ShiftMinimizerKernels.cu,
compiled unmodified into oracle/_ref/libmfsr_ref.so) around cuBLAS batched normal equations (oracle/ref_driver.cu:
ref_consolidate), on a B200.

    gpurun -- 'python tests/golden/make_ref_golden_r2.py gpurun_out/ref_golden_r2.npz'
    cp gpurun_out/ref_golden_r2.npz tests/golden/ref_golden_r2.npz

Inputs and reference outputs are both stored: tests/test_oracle_golden.py needs neither a GPU nor /root/reference.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyref  # noqa: E402

CASES = [(5, 2, 2), (8, 2, 4), (8, 7, 0), (15, 2, 7), (3, 2, 2)]      # (frames, pair span, reference frame)


def pairs_of(n, span):
    return [(i, j) for i in range(n) for j in range(i + 1, min(n, i + span + 1))]


def make_case(rng, n, span, tx=9, ty=7):
    pairs = pairs_of(n, span)
    nt = tx * ty
    seq = rng.uniform(-2, 2, size=(nt, n - 1, 2)).astype(np.float32)
    meas = np.zeros((nt, len(pairs), 2), np.float32)
    for k, (i, j) in enumerate(pairs):
        meas[:, k] = seq[:, i:j].sum(axis=1)
    meas += rng.normal(0, 0.02, size=meas.shape).astype(np.float32)
    # Gross outliers only where they are IDENTIFIABLE: an unknown that appears in just two measurements (the first and last
    # sequential shift when the pair span is 2) gives those two rows residuals of exactly equal magnitude, so which of them
    # `dist > max` (ShiftMinimizerKernels.cu:118) removes is decided by the round-off of the linear algebra, not by the data.
    cover = np.zeros(n - 1, int)
    for (i, j) in pairs:
        cover[i:j] += 1
    ok = [k for k, (i, j) in enumerate(pairs) if cover[i:j].min() >= 3]
    if ok:
        bad = rng.integers(0, len(ok), size=(nt, 2))
        for t in range(0, nt, 3):
            meas[t, ok[bad[t, 0]]] += 7.0
    return pairs, meas


def main(out_path):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(20261019)
    G = {}
    for ci, (n, span, ref) in enumerate(CASES):
        pairs, meas = make_case(rng, n, span)
        pf, pt = [a for a, _ in pairs], [b for _, b in pairs]
        o2o, fs, st, rm = pyref.consolidate(torch.from_numpy(meas).to(dev), pf, pt, n, 9, 7, ref)
        G[f"cs{ci}_cfg"] = np.array([n, span, ref, 9, 7], np.int32)
        G[f"cs{ci}_meas"] = meas
        G[f"cs{ci}_one_to_one"] = o2o.cpu().numpy()
        G[f"cs{ci}_frame_shift"] = fs.cpu().numpy()
        G[f"cs{ci}_status"] = st.cpu().numpy()
        G[f"cs{ci}_removed"] = rm.cpu().numpy()
        print(f"case {ci}: n={n} span={span} pairs={len(pairs)} removed total {int(rm.sum())} status min/max {int(st.min())}/{int(st.max())}")
    Path(out_path).parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out_path, **G)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_golden_r2.npz")

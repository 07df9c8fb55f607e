"""Shared helpers for the parity tests."""
import numpy as np
import torch


def u16(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint16)


def psnr(a: np.ndarray, b: np.ndarray, peak: float = 1.0) -> float:
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 200.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def max_abs(a, b) -> float:
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))


WHITE = [959.0, 959.0, 959.0]
BLACK = [64.0, 64.0, 64.0]
SCALE = [np.float32(1.0) / np.float32(959.0)] * 3
RGGB = [0, 1, 1, 2]

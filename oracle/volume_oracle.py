"""TEST INFRASTRUCTURE (checker, not product): CPU restatement of the volume-prediction front / back end of the
reference, engine/test_volume.py:135-191 and :269-294.  Imported only by tests/, like oracle/mudiff_oracle.py.

Parity status: the arithmetic is numpy's / ATen's own (np.percentile, np.clip, F.interpolate), called the way the
reference calls them; the reference has no tests or golden vectors for these functions, so this file is
pinned by construction (a line-by-line restatement) and by tests/test_oracle.py::test_volume_oracle_matches_reference
(which imports the reference module text when /root/reference is present and compares on random volumes).
"""
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def robust_minmax_to_minus1_1(vol: np.ndarray, mask: Optional[np.ndarray] = None, pmin: float = 1.0,
                              pmax: float = 99.0) -> np.ndarray:
    """engine/test_volume.py:135-157."""
    data = vol.astype(np.float32, copy=False)
    m = (data != 0) if mask is None else (mask.astype(bool) & (data == data))
    if not np.any(m):
        return np.zeros_like(data, dtype=np.float32)
    vals = data[m]
    lo = np.percentile(vals, pmin)
    hi = np.percentile(vals, pmax)
    if not np.isfinite(lo) or not np.isfinite(hi) or hi <= lo:
        lo, hi = float(vals.min()), float(vals.max())
        if hi <= lo:
            return np.zeros_like(data, dtype=np.float32)
    x01 = np.clip((data - lo) / (hi - lo), 0.0, 1.0)
    return x01 * 2.0 - 1.0


def extract_center_slices(volume: np.ndarray, half_range: int) -> Tuple[List[np.ndarray], int, int]:
    """engine/test_volume.py:159-168."""
    z = volume.shape[2]
    c = z // 2
    start = max(0, c - half_range)
    end = min(z - 1, c + half_range)
    return [volume[:, :, idx] for idx in range(start, end + 1)], start, end


def reconstruct_volume_from_slices(predicted_slices, original_shape, start_slice: int, end_slice: int) -> np.ndarray:
    """engine/test_volume.py:170-181."""
    vol = np.zeros(original_shape, dtype=np.float32)
    for i, sl in enumerate(predicted_slices):
        k = start_slice + i
        if start_slice <= k <= end_slice and k < original_shape[2]:
            vol[:, :, k] = np.asarray(sl).astype(np.float32, copy=False)
    return vol


def preprocess_volume(vol: np.ndarray, half_range: int, image_size: int) -> Tuple[torch.Tensor, int, int]:
    """load_and_preprocess_volume (:183-192) + the per-slice tensor construction of predict_volume (:270-276):
    [n, 1, image_size, image_size] fp32 conditioning slices in [-1, 1]."""
    vol_norm = robust_minmax_to_minus1_1(vol)
    slices, s0, s1 = extract_center_slices(vol_norm, half_range)
    out = []
    for sl in slices:
        t = torch.from_numpy(np.ascontiguousarray(sl).astype(np.float32, copy=False)).unsqueeze(0).unsqueeze(0)
        if t.shape[-2:] != (image_size, image_size):
            t = F.interpolate(t, size=(image_size, image_size), mode='bilinear', align_corners=False)
        out.append(t)
    return torch.cat(out, 0), s0, s1


def postprocess_slices(fake: torch.Tensor) -> np.ndarray:
    """:285: ((fake + 1) / 2).clamp(0, 1) -> numpy [n, H, W]."""
    return ((fake + 1.0) / 2.0).clamp(0.0, 1.0).cpu().numpy()[:, 0]

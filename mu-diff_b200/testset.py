"""Slice-test driver of MU-Diff (engine/test.py:265-400 `sample_and_test`) on the fast path.

The reference walks the BraTS test split with a DataLoader of batch_size=1 (:292-300), samples every slice on its own,
keeps all predictions / ground truths on the host and finally writes 8-bit PNGs scaled with one GLOBAL intensity window
(:366-390).  Here the slices are sampled in batches (a `volume.GraphSliceSampler` replays one CUDA graph per batch), the
dataset normalisation (dataset/dataset_brats.py:83,91: clamp(z, -3, 3) / 3), the global min / max and the uint8 export
run as CUDA kernels, and only the finished uint8 images go back to the host.  File I/O (np.load of the `.npy`
volumes, PNG writing through PIL) stays host-side.

RNG: the reference draws x_T / z / posterior noise from one sequential device stream, slice after slice, which a
batched sampler cannot reproduce; like `volume.py`, every slice has its own stream seeded by (seed, 0, slice index).
"""
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import volume as V

ORDERS: Dict[str, List[str]] = {          # dataset/dataset_brats.py:29-34
    "T1CE": ["FLAIR", "T2", "T1", "T1CE"],
    "FLAIR": ["T1CE", "T1", "T2", "FLAIR"],
    "T2": ["T1CE", "T1", "FLAIR", "T2"],
    "T1": ["FLAIR", "T1CE", "T2", "T1"],
}


def load_split(base_path: str, split: str = 'test', target_modality: str = 'T1CE') -> List[np.ndarray]:
    """The four `(N, H, W)` z-score arrays of a split in the reference's order [cond1, cond2, cond3, target]
    (dataset/dataset_brats.py:52-66)."""
    if target_modality not in ORDERS:
        raise ValueError(f"Invalid target_modality {target_modality}.")
    out = []
    for mod in ORDERS[target_modality]:
        fp = os.path.join(base_path, split, f"{mod}.npy")
        if not os.path.isfile(fp):
            raise FileNotFoundError(fp)
        out.append(np.ascontiguousarray(np.load(fp, allow_pickle=False), dtype=np.float32))
    return out


def zscore_to_unit(x: torch.Tensor) -> torch.Tensor:
    """clamp(x, -3, 3) / 3 on the device (dataset/dataset_brats.py:83,91)."""
    L.require_cuda(x)
    x = x.float().contiguous()
    out = torch.empty_like(x)
    L.check(L.lib().mudiff_zscore_to_unit(x.data_ptr(), out.data_ptr(), x.numel(), L.stream_ptr(x.device)), 'zscore_to_unit')
    return out


def global_window(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """Device uint32[2] keys of the exact global (min, max) over all `tensors` (engine/test.py:366-371)."""
    dev = tensors[0].device
    keys = torch.empty(2, dtype=torch.int32, device=dev)
    for i, t in enumerate(tensors):
        L.require_cuda(t)
        t = t.float().contiguous()
        L.check(L.lib().mudiff_minmax_keys(t.data_ptr(), t.numel(), 1 if i else 0, keys.data_ptr(), L.stream_ptr(dev)), 'minmax_keys')
    return keys


def window_values(keys: torch.Tensor) -> Tuple[float, float]:
    out = torch.empty(2, dtype=torch.float32, device=keys.device)
    L.check(L.lib().mudiff_minmax_read(keys.data_ptr(), out.data_ptr(), L.stream_ptr(keys.device)), 'minmax_read')
    lo, hi = out.tolist()
    return lo, hi


def scale_to_u8(x: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
    """np.clip((x - gmin) / (gmax - gmin) * 255, 0, 255).astype(uint8) with the global window (engine/test.py:378-388)."""
    L.require_cuda(x, keys)
    x = x.float().contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    L.check(L.lib().mudiff_scale_to_u8(x.data_ptr(), x.numel(), keys.data_ptr(), out.data_ptr(), L.stream_ptr(x.device)), 'scale_to_u8')
    return out


def sample_test_split(sample_fn: Callable, arrays: Sequence[np.ndarray], *, batch: int = 64, seed: int = 42, nz: int = 100,
                      n_time: int = 4, device='cuda') -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched `sample_and_test` loop (engine/test.py:324-364): arrays = [cond1, cond2, cond3, target] z-score `(N, H, W)`.
    Returns (pred, gt) as device tensors [N, 1, H, W] in the model's [-1, 1] range."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError("mu-diff_b200: the slice-test driver runs on the GPU only")
    *conds_np, target_np = arrays
    n = target_np.shape[0]
    hw = tuple(target_np.shape[-2:])
    pinned = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory() for a in arrays]
    pred = torch.empty((n, 1) + hw, dtype=torch.float32, device=device)
    gt = zscore_to_unit(pinned[-1].to(device, non_blocking=True)).view(n, 1, *hw)
    bsz = V.balanced_batch(n, batch)
    for b0 in range(0, n, bsz):
        b1 = min(n, b0 + bsz)
        cb = [zscore_to_unit(p[b0:b1].to(device, non_blocking=True)).view(b1 - b0, 1, *hw) for p in pinned[:-1]]
        x_init, latents, noises = V.draw_slice_noise(seed, 0, list(range(b0, b1)), hw, nz, n_time, device)
        pred[b0:b1] = sample_fn(cb, x_init, latents, noises)
    return pred, gt


def export_uint8(pred: torch.Tensor, gt: torch.Tensor) -> Tuple[np.ndarray, np.ndarray, Tuple[float, float]]:
    """Global-window 8-bit images of predictions and ground truth (engine/test.py:366-388) as numpy uint8 [N, H, W]."""
    keys = global_window([pred, gt])
    p8, g8 = scale_to_u8(pred, keys), scale_to_u8(gt, keys)
    return p8[:, 0].cpu().numpy(), g8[:, 0].cpu().numpy(), window_values(keys)


def save_pngs(pred_u8: np.ndarray, gt_u8: np.ndarray, save_dir: str) -> None:
    """engine/test.py:311-316,385-390: <save_dir>/pred/pred_%05d.png and <save_dir>/gt/gt_%05d.png (host I/O, PIL)."""
    from PIL import Image
    pred_dir, gt_dir = os.path.join(save_dir, 'pred'), os.path.join(save_dir, 'gt')
    os.makedirs(pred_dir, exist_ok=True)
    os.makedirs(gt_dir, exist_ok=True)
    for i in range(pred_u8.shape[0]):
        Image.fromarray(pred_u8[i]).save(os.path.join(pred_dir, f"pred_{i:05d}.png"))
        Image.fromarray(gt_u8[i]).save(os.path.join(gt_dir, f"gt_{i:05d}.png"))

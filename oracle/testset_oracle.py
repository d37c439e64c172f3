"""TEST INFRASTRUCTURE (checker, not product): CPU restatement of the slice-test driver's arithmetic,
dataset/dataset_brats.py:73-92 (z-score -> [-1, 1]) and engine/test.py:366-388 (global window, 8-bit export).
Pinned by construction: these ARE the reference's numpy / torch expressions, line by line; the reference has no tests
for them.  Imported only by tests/."""
import numpy as np
import torch


def zscore_to_unit(img: np.ndarray) -> torch.Tensor:
    """dataset_brats.py:79-83 / :89-91."""
    t = torch.from_numpy(img.astype(np.float32))
    return torch.clamp(t, -3.0, 3.0) / 3.0


def export_uint8(all_pred_slices, all_gt_slices):
    """engine/test.py:366-388, without the file writes."""
    all_pred_array = np.concatenate([p.flatten() for p in all_pred_slices])
    all_gt_array = np.concatenate([g.flatten() for g in all_gt_slices])
    global_min = float(min(all_pred_array.min(), all_gt_array.min()))
    global_max = float(max(all_pred_array.max(), all_gt_array.max()))
    if global_max <= global_min:
        global_min, global_max = 0.0, 1.0
    preds, gts = [], []
    for pred_slice, gt_slice in zip(all_pred_slices, all_gt_slices):
        preds.append(np.clip((pred_slice - global_min) / (global_max - global_min) * 255.0, 0, 255).astype(np.uint8))
        gts.append(np.clip((gt_slice - global_min) / (global_max - global_min) * 255.0, 0, 255).astype(np.uint8))
    return np.stack(preds), np.stack(gts), (global_min, global_max)

"""Generates tests/golden/volume.npz by RUNNING THE REFERENCE functions of engine/test_volume.py (build container only:
needs /root/reference).  The module itself imports nibabel (absent here), so the three pure functions are exec'd from
its source text, unmodified; the per-slice tensor construction + F.interpolate of predict_volume (:270-276) and the
post-processing of :285 are the reference's own torch calls.
    python tests/golden/make_volume_golden.py
"""
import ast
import os
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

REF = '/root/reference/engine/test_volume.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'volume.npz')
WANT = ('robust_minmax_to_minus1_1', 'extract_center_slices', 'reconstruct_volume_from_slices')

src = open(REF).read()
tree = ast.parse(src)
ns = dict(np=np, Optional=Optional, Tuple=Tuple, List=List, Dict=Dict)
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in WANT:
        exec(compile(ast.Module(body=[node], type_ignores=[]), REF, 'exec'), ns)
robust, extract, rebuild = (ns[k] for k in WANT)

rng = np.random.default_rng(7)
store = {}
cases = {
    # MRI-like: integer intensities stored as float64 (nibabel get_fdata), background exactly 0, a few negatives
    'mri': (np.where(rng.random((24, 20, 31)) < 0.35, 0.0, np.round(rng.gamma(2.0, 180.0, (24, 20, 31)))) - 3.0 * (rng.random((24, 20, 31)) < 0.01)),
    'smooth': rng.normal(50.0, 20.0, (16, 16, 9)),
    'flat': np.full((8, 8, 5), 7.0),
    'zeros': np.zeros((8, 8, 5)),
    'two_values': np.where(rng.random((12, 12, 7)) < 0.5, 1.0, 2.0),
}
for name, vol in cases.items():
    half, size = (10, 32) if name == 'mri' else (2, vol.shape[0])
    vol_norm = robust(vol)
    slices, s0, s1 = extract(vol_norm, half)
    ts = []
    for sl in slices:
        t = torch.from_numpy(sl.astype(np.float32, copy=False)).unsqueeze(0).unsqueeze(0)
        if t.shape[-2:] != (size, size):
            t = F.interpolate(t, size=(size, size), mode='bilinear', align_corners=False)
        ts.append(t)
    conds = torch.cat(ts, 0)
    fake = torch.from_numpy(rng.normal(0.0, 0.8, (len(slices), 1, vol.shape[0], vol.shape[1])).astype(np.float32))
    pred = ((fake + 1.0) / 2.0).clamp(0.0, 1.0).cpu().numpy()
    rebuilt = rebuild([pred[i, 0] for i in range(pred.shape[0])], vol.shape, s0, s1)
    store[f'{name}_vol'] = vol
    store[f'{name}_norm'] = np.asarray(vol_norm)
    store[f'{name}_conds'] = conds.numpy()
    store[f'{name}_p'] = np.array([half, size, s0, s1])
    store[f'{name}_fake'] = fake.numpy()
    store[f'{name}_rebuilt'] = rebuilt
np.savez_compressed(OUT, **store)
print('wrote', OUT, {k: v.shape for k, v in store.items() if k.endswith('_conds')})

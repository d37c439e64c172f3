"""Generate tests/golden/testset.npz by RUNNING THE REFERENCE's slice-test data path (build container only):

    python tests/golden/make_testset_golden.py

  * `dataset.dataset_brats.BratsDataset` (dataset/dataset_brats.py:8-92) is imported from /root/reference and
    run on a temporary BraTS-style split of `.npy` files (inputs are stored in the fixture, so the test can
    re-create the very same files on the GPU box);
  * the global-window 8-bit export block of `engine/test.py` (`sample_and_test`, the statements from
    `all_pred_array = np.concatenate(...)` to the PNG loop, :367-388) is cut out of the reference's SOURCE TEXT
    (the module itself needs skimage/matplotlib) and exec'd with a stand-in `Image` that records the arrays
    instead of writing files.
/root/reference does not exist on the GPU box: the outputs are committed fixtures.
"""
import os
import sys
import tempfile
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, REF)


class _Recorder:
    """Stands in for PIL.Image inside the exec'd block: Image.fromarray(a).save(path) records `a` under `path`."""

    def __init__(self):
        self.saved = {}

    def fromarray(self, arr):
        rec = self

        class _Img:
            def save(self, path):
                rec.saved[os.path.basename(path)] = np.array(arr)
        return _Img()


def export_block_source():
    src = open(os.path.join(REF, 'engine/test.py')).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.strip().startswith('all_pred_array = np.concatenate'))
    end = next(i for i, l in enumerate(src) if i > start and 'Successfully completed testing' in l)
    return textwrap.dedent("\n".join(src[start:end]))


def run_export(pred_slices, gt_slices):
    import logging
    rec = _Recorder()
    ns = dict(np=np, os=os, logging=logging, Image=rec, all_pred_slices=pred_slices, all_gt_slices=gt_slices,
              pred_dir='pred', gt_dir='gt')
    exec(compile(export_block_source(), 'engine/test.py[367:388]', 'exec'), ns)
    n = len(pred_slices)
    p8 = np.stack([rec.saved[f"pred_{i:05d}.png"] for i in range(n)])
    g8 = np.stack([rec.saved[f"gt_{i:05d}.png"] for i in range(n)])
    return p8, g8, np.array([ns['global_min'], ns['global_max']], dtype=np.float64)


def main():
    from dataset.dataset_brats import BratsDataset
    rng = np.random.default_rng(2024)
    n, h, w = 9, 24, 20
    out = {}
    mods = {m: rng.normal(0.0, 1.7, (n, h, w)).astype(np.float32) for m in ('FLAIR', 'T2', 'T1', 'T1CE')}
    mods['T2'][3, 5, 7] = 3.0                      # exact clamp boundaries and far outliers
    mods['T2'][3, 5, 8] = -3.0
    mods['T1'][0, 0, 0] = 41.5
    mods['T1CE'][8, 23, 19] = -17.25
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, 'test'))
        for m, a in mods.items():
            np.save(os.path.join(td, 'test', f'{m}.npy'), a)
            out[f'in_{m}'] = a
        for target in ('T1CE', 'FLAIR', 'T2', 'T1'):
            ds = BratsDataset(split='test', base_path=td, target_modality=target)
            assert len(ds) == n
            items = [ds[i] for i in range(n)]
            out[f'{target}_cond'] = torch.stack([c for c, _ in items]).numpy()      # [n, 3, h, w]
            out[f'{target}_target'] = torch.stack([t for _, t in items]).numpy()    # [n, 1, h, w]
            out[f'{target}_order'] = np.array(ds.modality_order)
    # export block on (prediction, ground truth) pairs: a tanh-range prediction vs the dataset's target, a wider
    # range (window set by the prediction), and the constant-image fallback
    gt = out['T1CE_target'][:, 0]
    pred = np.tanh(rng.normal(0.0, 1.0, gt.shape)).astype(np.float32)
    for tag, p, g in (('a', pred, gt), ('b', (pred * 1.7 - 0.2).astype(np.float32), gt),
                      ('const', np.full((2, 4, 4), 0.25, np.float32), np.full((2, 4, 4), 0.25, np.float32))):
        p8, g8, win = run_export(list(p), list(g))
        out[f'exp_{tag}_pred'], out[f'exp_{tag}_gt'] = p, g
        out[f'exp_{tag}_p8'], out[f'exp_{tag}_g8'], out[f'exp_{tag}_win'] = p8, g8, win
    np.savez_compressed(os.path.join(HERE, 'testset.npz'), **out)
    print('wrote testset.npz:', {k: v.shape for k, v in out.items() if k.startswith('exp_') or k.endswith('_cond')})


if __name__ == '__main__':
    main()

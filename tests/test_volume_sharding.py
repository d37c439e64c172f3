"""Multi-GPU path exercised on CPU with gloo, world_size 2: shard bounds,
per-slice RNG independence of the sharding, one all-gather, volume re-stacking."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_sampler(conds, x_init, latents, noises):
    """Deterministic stand-in for the 4-step sampler (a function of every input it is given)."""
    y = 0.3 * conds[0] - 0.2 * conds[1] + 0.1 * conds[2] + 0.05 * x_init
    for z, e in zip(latents, noises):
        y = y + 0.01 * e + 0.001 * z.mean(dim=1).view(-1, 1, 1, 1)
    return torch.tanh(y)


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import mudiff_b200.volume as V
    g = torch.Generator().manual_seed(5)
    conds = [torch.rand(n, 1, 8, 8, generator=g) * 2 - 1 for _ in range(3)]
    full = V.predict_slices_sharded(_fake_sampler, conds, seed=7, volume=3, nz=10, n_time=4, batch=3, device='cpu')
    np.save(os.path.join(out_dir, f'r{rank}.npy'), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [11, 8])
def test_sharded_equals_single_process(tmp_path, n):
    import mudiff_b200.volume as V
    g = torch.Generator().manual_seed(5)
    conds = [torch.rand(n, 1, 8, 8, generator=g) * 2 - 1 for _ in range(3)]
    ref = V.predict_slices_sharded(_fake_sampler, conds, seed=7, volume=3, nz=10, n_time=4, batch=4, device='cpu')
    assert tuple(ref.shape) == (n, 1, 8, 8)
    port = 29500 + (os.getpid() % 2000) + n
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / 'r0.npy'), np.load(tmp_path / 'r1.npy')
    np.testing.assert_array_equal(r0, r1)                 # every rank holds the whole volume
    np.testing.assert_array_equal(r0, ref.numpy())        # identical for any world size / batch size


def test_shard_bounds_cover_exactly():
    import mudiff_b200.volume as V
    for n in (0, 1, 7, 155, 161):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi, per = V.shard_bounds(n, world, r)
                assert hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_volume_oracle_pre_post_processing():
    """The numpy restatement of engine/test_volume.py:135-181 (oracle/volume_oracle.py) and the index helpers."""
    import mudiff_b200.volume as V
    from oracle import volume_oracle as VO
    rng = np.random.default_rng(0)
    vol = rng.random((8, 8, 21)).astype(np.float32) * 100
    vol[0, 0, :] = 0
    x = VO.robust_minmax_to_minus1_1(vol)
    assert x.min() >= -1 and x.max() <= 1 and x.dtype == np.float32
    vals = vol[vol != 0]
    lo, hi = np.percentile(vals, 1), np.percentile(vals, 99)
    np.testing.assert_allclose(x, np.clip((vol - lo) / (hi - lo), 0, 1) * 2 - 1, atol=1e-6)
    assert VO.robust_minmax_to_minus1_1(np.zeros((2, 2, 2))).sum() == 0
    assert V.center_slice_bounds(155, 80) == (0, 154)
    assert V.center_slice_bounds(200, 80) == (20, 180)
    sl, s0, s1 = VO.extract_center_slices(vol, 5)
    assert (s0, s1) == V.center_slice_bounds(21, 5) == (5, 15) and len(sl) == 11
    out = VO.reconstruct_volume_from_slices([np.ones((8, 8))] * 11, vol.shape, s0, s1)
    assert np.all(out[:, :, :5] == 0) and np.all(out[:, :, 16:] == 0) and np.all(out[:, :, 5:16] == 1)
    with pytest.raises(RuntimeError):                     # the product pre/post-processing has no CPU path
        V.predict_volume(_fake_sampler, [vol, vol, vol], slice_half_range=5, nz=10, batch=4, device='cpu')

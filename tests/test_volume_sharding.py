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


def _make_volumes(sizes):
    g = torch.Generator().manual_seed(9)
    return [[torch.rand(n, 1, 8, 8, generator=g) * 2 - 1 for _ in range(3)] for n in sizes]


def _worker_multi(rank, world, port, sizes, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import mudiff_b200.volume as V
    outs = V.predict_volumes_sharded(_fake_sampler, _make_volumes(sizes), seed=11, first_volume=4, nz=10, n_time=4, batch=3,
                                     device='cpu')
    for v, o in enumerate(outs):
        np.save(os.path.join(out_dir, f'r{rank}_v{v}.npy'), o.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_multi_volume_sharded_equals_single_process(tmp_path):
    """predict_volumes_sharded: several volumes of different depth (one shorter than the world size, so a rank owns
    nothing of it) walked as one pipelined work list, one all-gather per volume, identical to the unsharded result
    and to per-volume predict_slices_sharded calls."""
    import mudiff_b200.volume as V
    sizes = [7, 1, 10]
    vols = _make_volumes(sizes)
    ref = V.predict_volumes_sharded(_fake_sampler, vols, seed=11, first_volume=4, nz=10, n_time=4, batch=4, device='cpu')
    for v, (c, r) in enumerate(zip(vols, ref)):
        one = V.predict_slices_sharded(_fake_sampler, c, seed=11, volume=4 + v, nz=10, n_time=4, batch=2, device='cpu')
        np.testing.assert_array_equal(r.numpy(), one.numpy())
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_multi, args=(2, port, sizes, str(tmp_path)), nprocs=2, join=True)
    for v, r in enumerate(ref):
        a, b = np.load(tmp_path / f'r0_v{v}.npy'), np.load(tmp_path / f'r1_v{v}.npy')
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, r.numpy())


def test_balanced_batch():
    import mudiff_b200.volume as V
    assert V.balanced_batch(155, 64) == 52 and V.balanced_batch(78, 64) == 39 and V.balanced_batch(20, 64) == 20
    assert V.balanced_batch(64, 64) == 64 and V.balanced_batch(65, 64) == 33 and V.balanced_batch(0, 64) == 1
    for n in range(1, 400):
        b = V.balanced_batch(n, 64)
        k = -(-n // b)
        assert b <= 64 and k == -(-n // 64) and k * b - n < k       # same number of batches, < 1 padded row per batch


def test_packed_batches_fill_across_volumes():
    """pack=True: a rank's batches are filled across volume boundaries (one batch of `batch` slices instead of one short batch
    per volume), pack=False keeps one batch per (volume, range); both give identical volumes (per-slice RNG streams)."""
    import mudiff_b200.volume as V
    vols = _make_volumes([5, 5, 5, 3, 7])
    sizes = {True: [], False: []}

    def sampler_for(flag):
        def f(conds, x_init, latents, noises):
            sizes[flag].append(x_init.shape[0])
            return _fake_sampler(conds, x_init, latents, noises)
        return f

    kw = dict(seed=11, first_volume=4, nz=10, n_time=4, batch=8, device='cpu')
    packed = V.predict_volumes_sharded(sampler_for(True), vols, pack=True, **kw)
    plain = V.predict_volumes_sharded(sampler_for(False), vols, pack=False, **kw)
    assert sizes[True] == [8, 8, 8, 1] and sizes[False] == [5, 5, 5, 3, 7]
    for a, b in zip(packed, plain):
        np.testing.assert_array_equal(a.numpy(), b.numpy())

"""Full-volume prediction with axial slices sharded across the GPUs of one box
(BASELINE.json configs[2]; engine/test_volume.py:135-181,269-294 is the single-GPU, B=1
reference loop this replaces).

One process per GPU (torch.distributed, NCCL over NVLink).  A volume's N slices are split into
contiguous shards of ceil(N/G); every rank samples its shard in batches (not B=1), and ONE
all-gather of the [ceil(N/G), 1, H, W] shards rebuilds the volume on every rank.  No other
collective is on the data path (slices are independent).

RNG: the reference draws noise from one sequential stream across slices (test_volume.py:213,279),
which cannot be reproduced under sharding; here every slice has its own stream seeded by
(seed, volume, slice) so the result is identical for any world size / batch size.

NIfTI I/O (nibabel) is outside the hot path: volumes are numpy / torch arrays [H, W, Z].
"""
import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ---- host-side pre/post processing (engine/test_volume.py:135-181) ------------------
def robust_minmax_to_minus1_1(vol: np.ndarray, mask: Optional[np.ndarray] = None, pmin: float = 1.0,
                              pmax: float = 99.0) -> np.ndarray:
    """[pmin, pmax] percentile window over non-zero voxels -> [-1, 1] (test_volume.py:135-157)."""
    data = vol.astype(np.float32, copy=False)
    m = (data != 0) if mask is None else (mask.astype(bool) & (data == data))
    if not np.any(m):
        return np.zeros_like(data, dtype=np.float32)
    vals = data[m]
    lo, hi = np.percentile(vals, pmin), np.percentile(vals, pmax)
    if not np.isfinite(lo) or not np.isfinite(hi) or hi <= lo:
        lo, hi = float(vals.min()), float(vals.max())
        if hi <= lo:
            return np.zeros_like(data, dtype=np.float32)
    return np.clip((data - lo) / (hi - lo), 0.0, 1.0) * 2.0 - 1.0


def center_slice_bounds(depth: int, half_range: int) -> Tuple[int, int]:
    """Inclusive [start, end] of the centre +-half_range axial slices (test_volume.py:159-168)."""
    c = depth // 2
    return max(0, c - half_range), min(depth - 1, c + half_range)


def reconstruct_volume_from_slices(pred: np.ndarray, original_shape, start_slice: int, end_slice: int) -> np.ndarray:
    """pred [n, H, W] -> zeros(original_shape) with slices start..end filled (test_volume.py:170-181)."""
    vol = np.zeros(original_shape, dtype=np.float32)
    for i in range(pred.shape[0]):
        k = start_slice + i
        if start_slice <= k <= end_slice and k < original_shape[2]:
            vol[:, :, k] = pred[i]
    return vol


# ---- sharding --------------------------------------------------------------------------
def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous block of slices owned by `rank`: (lo, hi, shard_len) with shard_len = ceil(n/world)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per), per


def balanced_batch(n_local: int, max_batch: int = 64) -> int:
    """Batch size that splits a shard of n_local slices into equal batches of at most max_batch (a fixed-batch CUDA
    graph then wastes at most one padded row per batch instead of up to max_batch - 1)."""
    if n_local <= 0:
        return 1
    k = (n_local + max_batch - 1) // max_batch
    return (n_local + k - 1) // k


def slice_seed(seed: int, volume: int, index: int) -> int:
    return (int(seed) * 1000003 + int(volume) * 8191 + int(index) * 131 + 17) % (2 ** 63 - 1)


def draw_slice_noise(seed, volume, indices: Sequence[int], size, nz, n_time, device):
    """Per-slice RNG streams: x_init, latents[n_time], noises[n_time] for the given slice indices,
    in the reference's draw order (x_init first - test_volume.py:279 - then z, noise per step)."""
    b = len(indices)
    x_init = torch.empty(b, 1, size[0], size[1], device=device)
    latents = [torch.empty(b, nz, device=device) for _ in range(n_time)]
    noises = [torch.empty(b, 1, size[0], size[1], device=device) for _ in range(n_time)]
    g = torch.Generator(device=device)
    for j, idx in enumerate(indices):
        g.manual_seed(slice_seed(seed, volume, idx))
        x_init[j].normal_(generator=g)
        for i in reversed(range(n_time)):
            latents[i][j].normal_(generator=g)
            noises[i][j].normal_(generator=g)
    return x_init, latents, noises


def predict_volumes_sharded(sample_fn: Callable, cond_volumes: Sequence[Sequence[torch.Tensor]], *, seed: int = 0,
                            first_volume: int = 0, nz: int = 100, n_time: int = 4, batch: int = 64, device=None,
                            group=None, gather: bool = True, prefetch: bool = True) -> List[torch.Tensor]:
    """Sample every slice of several volumes across the ranks of `group`.

    cond_volumes[v] = n_cond tensors [N_v, 1, H, W] in [-1, 1] (same on every rank; only the local shard is used).
    sample_fn(conds_batch, x_init, latents, noises) -> [b, 1, H, W]: the 4-step sampler, e.g. a GraphSliceSampler or
    `lambda c, x, z, e: sample_from_model(co, g1, c[0], g2, c[1], c[2], 4, x, None, opt, latents=z, noises=e)`.
    Every volume is split into contiguous shards of ceil(N_v / G) slices; a rank walks its shards batch by batch.
    With `prefetch` the inputs of the NEXT batch (host->device copy of the conditioning slices, per-slice noise
    draws) are produced on a side stream while the current batch is being sampled, across volume boundaries too -
    with 8 GPUs a rank has one batch of ~20 slices per volume and this host/RNG work would otherwise be serial.
    Returns one [N_v, 1, H, W] tensor per volume, in [0, 1] ((x+1)/2 clamped, test_volume.py:285), on every rank if
    `gather` (ONE all-gather per volume - the only collective of the path), else the padded local shard."""
    if not cond_volumes:
        return []
    device = torch.device(device) if device is not None else cond_volumes[0][0].device
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    cuda = device.type == 'cuda'
    main = torch.cuda.current_stream(device) if cuda else None
    side = torch.cuda.Stream(device=device) if (cuda and prefetch) else None

    items = []                                   # (volume position, b0, b1) of this rank, in order
    bounds = []
    for v, conds in enumerate(cond_volumes):
        lo, hi, per = shard_bounds(conds[0].shape[0], world, rank)
        bounds.append((lo, hi, per))
        for b0 in range(lo, hi, batch):
            items.append((v, b0, min(hi, b0 + batch)))

    def prepare(item):
        v, b0, b1 = item
        conds = cond_volumes[v]
        hw = tuple(conds[0].shape[-2:])
        ctx = torch.cuda.stream(side) if side is not None else _null_ctx()
        with ctx:
            cb = [c[b0:b1].to(device, non_blocking=True) for c in conds]
            x_init, latents, noises = draw_slice_noise(seed, first_volume + v, list(range(b0, b1)), hw, nz, n_time, device)
            ev = side.record_event() if side is not None else None
        if side is not None:
            for t in cb + [x_init] + latents + noises:
                t.record_stream(main)
        return cb, x_init, latents, noises, ev

    outs: List[Optional[torch.Tensor]] = [None] * len(cond_volumes)
    locals_: List[Optional[torch.Tensor]] = [None] * len(cond_volumes)

    def finish(v):
        lo, hi, per = bounds[v]
        conds = cond_volumes[v]
        n, hw = conds[0].shape[0], tuple(conds[0].shape[-2:])
        local = locals_[v]
        if local is None:                        # this rank owns no slice of the volume
            local = torch.zeros(per, 1, hw[0], hw[1], device=device)
        if world == 1 or not gather:
            outs[v] = local[:hi - lo] if world == 1 else local
        else:
            full = torch.empty(world * per, 1, hw[0], hw[1], device=device)
            dist.all_gather_into_tensor(full, local, group=group)       # the ONLY collective of the path
            outs[v] = full[:n]
        locals_[v] = None

    nxt = prepare(items[0]) if items else None
    done_upto = 0                                # volumes [0, done_upto) are finished (gathers stay in volume order)
    for i, (v, b0, b1) in enumerate(items):
        cb, x_init, latents, noises, ev = nxt
        if ev is not None:
            main.wait_event(ev)
        fake = sample_fn(cb, x_init, latents, noises)
        lo, hi, per = bounds[v]
        if locals_[v] is None:
            locals_[v] = torch.zeros(per, 1, fake.shape[-2], fake.shape[-1], device=device)   # padded shard
        locals_[v][b0 - lo:b1 - lo] = ((fake + 1.0) / 2.0).clamp(0.0, 1.0)
        nxt = prepare(items[i + 1]) if i + 1 < len(items) else None        # overlaps the sampling just enqueued
        last_of_volume = i + 1 == len(items) or items[i + 1][0] != v
        if last_of_volume:
            while done_upto <= v:                # also volumes in between of which this rank owns nothing
                finish(done_upto)
                done_upto += 1
    while done_upto < len(cond_volumes):
        finish(done_upto)
        done_upto += 1
    return outs


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def predict_slices_sharded(sample_fn: Callable, conds: Sequence[torch.Tensor], *, seed: int = 0, volume: int = 0,
                           nz: int = 100, n_time: int = 4, batch: int = 64, device=None,
                           group=None, gather: bool = True, prefetch: bool = True) -> torch.Tensor:
    """One volume: see predict_volumes_sharded.  conds: n_cond tensors [N, 1, H, W] in [-1, 1]."""
    return predict_volumes_sharded(sample_fn, [conds], seed=seed, first_volume=volume, nz=nz, n_time=n_time, batch=batch,
                                   device=device, group=group, gather=gather, prefetch=prefetch)[0]


class GraphSliceSampler:
    """`sample_fn` for predict_slices_sharded / predict_volume backed by ONE captured CUDA graph of the whole
    n_time-step loop at a fixed batch (sampling.GraphSampler): a shard of b <= batch slices is copied into the
    graph's static buffers, the graph is replayed, the first b outputs are returned.  Rows b..batch-1 keep whatever
    the previous call left there; the kernels are batch-invariant (tests), so they cannot influence rows 0..b-1.
    Removes the ~1500 eager launches per batch from the host's critical path - what limits scaling when a volume's
    155 slices are spread over 8 GPUs (20 slices per rank)."""

    def __init__(self, coefficients, generator1, generator2, n_time, batch, size, nz, n_cond=3, device='cuda'):
        from .sampling import GraphSampler
        self.batch = int(batch)
        self.gs = GraphSampler(coefficients, generator1, generator2, n_time, self.batch, size, nz, n_cond=n_cond,
                               device=device, warmup=1)

    def __call__(self, conds, x_init, latents, noises):
        b = x_init.shape[0]
        if b > self.batch:
            raise RuntimeError(f"mu-diff_b200: GraphSliceSampler captured for batch {self.batch}, got {b}")
        gs = self.gs
        for d, s in zip(gs.conds, conds):
            d[:b].copy_(s, non_blocking=True)
        gs.x_init[:b].copy_(x_init, non_blocking=True)
        for d, s in zip(gs.latents, latents):
            d[:b].copy_(s, non_blocking=True)
        for d, s in zip(gs.noises, noises):
            d[:b].copy_(s, non_blocking=True)
        return gs.replay()[:b].clone()


def predict_volume(sample_fn: Callable, volumes: Sequence[np.ndarray], *, slice_half_range: int = 80, seed: int = 0,
                   volume_index: int = 0, nz: int = 100, n_time: int = 4, batch: int = 64, device='cuda',
                   group=None) -> np.ndarray:
    """engine/test_volume.py:209-299 without the NIfTI I/O: normalise each input modality volume
    [H, W, Z], take the centre slices, sample them (sharded + batched), rebuild the [H, W, Z] volume."""
    shape = volumes[0].shape
    for v in volumes:
        if v.shape != shape:
            raise ValueError(f"All input volumes must share shape. Got {v.shape} vs {shape}")
    s0, s1 = center_slice_bounds(shape[2], slice_half_range)
    conds = []
    for v in volumes:
        vn = robust_minmax_to_minus1_1(v)
        sl = np.ascontiguousarray(np.moveaxis(vn[:, :, s0:s1 + 1], 2, 0))[:, None]      # [n,1,H,W]
        t = torch.from_numpy(sl.astype(np.float32, copy=False))
        conds.append(t.pin_memory() if torch.device(device).type == 'cuda' else t)   # pinned: the H2D copy is truly async
    pred = predict_slices_sharded(sample_fn, conds, seed=seed, volume=volume_index, nz=nz, n_time=n_time,
                                  batch=batch, device=torch.device(device), group=group)
    return reconstruct_volume_from_slices(pred[:, 0].cpu().numpy(), shape, s0, s1)

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
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ---- pre/post processing on the GPU (engine/test_volume.py:135-181, 270-276, 285) -------------------
# The numpy restatement of these functions lives in oracle/volume_oracle.py (test infrastructure).
def center_slice_bounds(depth: int, half_range: int) -> Tuple[int, int]:
    """Inclusive [start, end] of the centre +-half_range axial slices (test_volume.py:159-168)."""
    c = depth // 2
    return max(0, c - half_range), min(depth - 1, c + half_range)


def _require_cuda_device(device) -> torch.device:
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError("mu-diff_b200: volume pre/post-processing runs on the GPU only (no CPU path; the numpy "
                           "restatement is test infrastructure in oracle/volume_oracle.py)")
    return device


def volume_window(vol_dev: torch.Tensor, pmin: float = 1.0, pmax: float = 99.0) -> torch.Tensor:
    """Robust percentile window of a device volume (fp32): returns the opaque device workspace holding (lo, hi, status)
    for volume_to_slices.  Exact order statistics by a two-level radix select (mudiff_volume_window)."""
    from . import _lib as L
    lib = L.lib()
    ws = torch.empty(lib.mudiff_volume_workspace_bytes(), dtype=torch.uint8, device=vol_dev.device)
    L.check(lib.mudiff_volume_window(vol_dev.data_ptr(), vol_dev.numel(), float(pmin), float(pmax), ws.data_ptr(),
                                     L.stream_ptr(vol_dev.device)), 'volume_window')
    return ws


def window_values(ws: torch.Tensor) -> Tuple[float, float, int]:
    """(lo, hi, status) of a volume_window workspace (synchronises; for tests / logging)."""
    from . import _lib as L
    out = torch.empty(3, dtype=torch.float32, device=ws.device)
    L.check(L.lib().mudiff_volume_window_read(ws.data_ptr(), out.data_ptr(), L.stream_ptr(ws.device)), 'volume_window_read')
    lo, hi = out[:2].tolist()
    return lo, hi, int(out[2:].view(torch.int32).item())


def volume_to_slices(volume: np.ndarray, half_range: int, image_size: int, device='cuda', pmin: float = 1.0,
                     pmax: float = 99.0) -> Tuple[torch.Tensor, int, int]:
    """load_and_preprocess_volume + per-slice tensors (test_volume.py:183-192, 270-276) as two kernels: volume [H, W, Z]
    (any float dtype; cast to fp32 like :142) -> normalised [-1, 1] conditioning slices [n, 1, S, S] on `device`."""
    from . import _lib as L
    device = _require_cuda_device(device)
    h, w, z = volume.shape
    s0, s1 = center_slice_bounds(z, half_range)
    n = s1 - s0 + 1
    host = torch.from_numpy(np.ascontiguousarray(volume, dtype=np.float32))
    vol_dev = host.to(device, non_blocking=True)
    ws = volume_window(vol_dev, pmin, pmax)
    out = torch.empty((n, 1, image_size, image_size), dtype=torch.float32, device=device)
    L.check(L.lib().mudiff_volume_to_slices(vol_dev.data_ptr(), h, w, z, s0, n, image_size, image_size, ws.data_ptr(),
                                            out.data_ptr(), L.stream_ptr(device)), 'volume_to_slices')
    return out, s0, s1


def slices_to_volume(pred: torch.Tensor, original_shape, start_slice: int, to01: bool = False) -> torch.Tensor:
    """reconstruct_volume_from_slices (test_volume.py:170-181) on the device: pred [n, 1, H, W] -> [H, W, Z] fp32 with the
    other slices zero; `to01` also applies the ((x + 1) / 2).clamp(0, 1) of :285."""
    from . import _lib as L
    _require_cuda_device(pred.device)
    h, w, z = (int(v) for v in original_shape)
    n = pred.shape[0]
    if tuple(pred.shape[-2:]) != (h, w):
        # the reference's `vol[:, :, k] = sl` raises for mismatching shapes, too
        raise ValueError(f"could not broadcast input array from shape {tuple(pred.shape[-2:])} into shape {(h, w)}")
    pred = pred.float().contiguous()
    vol = torch.empty((h, w, z), dtype=torch.float32, device=pred.device)
    L.check(L.lib().mudiff_slices_to_volume(pred.data_ptr(), h, w, z, start_slice, n, 1 if to01 else 0, vol.data_ptr(),
                                            L.stream_ptr(pred.device)), 'slices_to_volume')
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
                            group=None, gather: bool = True, prefetch: bool = True, pack: bool = True) -> List[torch.Tensor]:
    """Sample every slice of several volumes across the ranks of `group`.

    cond_volumes[v] = n_cond tensors [N_v, 1, H, W] in [-1, 1] (same on every rank; only the local shard is used).
    sample_fn(conds_batch, x_init, latents, noises) -> [b, 1, H, W]: the 4-step sampler, e.g. a GraphSliceSampler or
    `lambda c, x, z, e: sample_from_model(co, g1, c[0], g2, c[1], c[2], 4, x, None, opt, latents=z, noises=e)`.
    Every volume is split into contiguous shards of ceil(N_v / G) slices; a rank walks its shards batch by batch.
    With `pack` a batch is filled ACROSS volume boundaries (same-sized slices only): with 8 GPUs a rank owns ~20 slices
    of each 155-slice volume, and three volumes' shards make one batch of ~60 instead of three batches of 20 (the
    sampler runs ~8 % faster per slice at that batch; per-slice RNG streams + batch-invariant kernels keep every output
    bit unchanged).
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

    pieces = []                                  # (volume position, b0, b1) of this rank, in order
    bounds = []
    for v, conds in enumerate(cond_volumes):
        lo, hi, per = shard_bounds(conds[0].shape[0], world, rank)
        bounds.append((lo, hi, per))
        for b0 in range(lo, hi, batch):
            pieces.append((v, b0, min(hi, b0 + batch)))
    # batches = lists of pieces; packed: consecutive pieces (possibly of different volumes) are cut / merged into batches
    # of exactly `batch` slices (the last one may be shorter)
    items: List[List[Tuple[int, int, int]]] = []
    if pack:
        cur, room = [], batch
        for v, b0, b1 in pieces:
            hw = tuple(cond_volumes[v][0].shape[-2:])
            if cur and tuple(cond_volumes[cur[0][0]][0].shape[-2:]) != hw:
                items.append(cur); cur, room = [], batch
            while b0 < b1:
                take = min(room, b1 - b0)
                cur.append((v, b0, b0 + take))
                b0 += take; room -= take
                if room == 0:
                    items.append(cur); cur, room = [], batch
        if cur:
            items.append(cur)
    else:
        items = [[pc] for pc in pieces]
    last_item_of = {}                            # volume position -> index of the last batch holding one of its pieces
    for i, it in enumerate(items):
        for v, _, _ in it:
            last_item_of[v] = i

    def prepare(item):
        ctx = torch.cuda.stream(side) if side is not None else _null_ctx()
        with ctx:
            parts = []
            for v, b0, b1 in item:
                conds = cond_volumes[v]
                hw = tuple(conds[0].shape[-2:])
                cb = [c[b0:b1].to(device, non_blocking=True) for c in conds]
                parts.append((cb,) + draw_slice_noise(seed, first_volume + v, list(range(b0, b1)), hw, nz, n_time, device))
            if len(parts) == 1:
                cb, x_init, latents, noises = parts[0]
            else:
                cb = [torch.cat([pt[0][k] for pt in parts]) for k in range(len(parts[0][0]))]
                x_init = torch.cat([pt[1] for pt in parts])
                latents = [torch.cat([pt[2][k] for pt in parts]) for k in range(n_time)]
                noises = [torch.cat([pt[3][k] for pt in parts]) for k in range(n_time)]
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
    for i, item in enumerate(items):
        cb, x_init, latents, noises, ev = nxt
        if ev is not None:
            main.wait_event(ev)
        fake = sample_fn(cb, x_init, latents, noises)
        fake01 = ((fake + 1.0) / 2.0).clamp(0.0, 1.0)
        off = 0
        for v, b0, b1 in item:
            lo, hi, per = bounds[v]
            if locals_[v] is None:
                locals_[v] = torch.zeros(per, 1, fake.shape[-2], fake.shape[-1], device=device)   # padded shard
            locals_[v][b0 - lo:b1 - lo] = fake01[off:off + (b1 - b0)]
            off += b1 - b0
        nxt = prepare(items[i + 1]) if i + 1 < len(items) else None        # overlaps the sampling just enqueued
        # every volume whose last piece is in this batch (and the volumes before it of which this rank owns nothing).  With
        # packed batches the ranks' batch boundaries fall on different volumes, and an all-gather issued between two batches
        # on the compute stream makes a rank wait for peers that are still one batch behind (8 ranks: 1.39 s instead of
        # 0.9 s for 8 volumes): the gathers are then issued after the rank's last batch, still one per volume, in volume order.
        if not (pack and world > 1 and gather):
            v_done = max((v for v, _, _ in item if last_item_of[v] == i), default=-1)
            while done_upto <= v_done:
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
                   group=None, image_size: Optional[int] = None) -> np.ndarray:
    """engine/test_volume.py:209-299 without the NIfTI I/O: normalise each input modality volume [H, W, Z] (robust
    percentile window, on the GPU), take the centre slices (bilinear resize to `image_size` if given and different),
    sample them (sharded + batched), rebuild the [H, W, Z] volume on the GPU and return it as numpy."""
    device = _require_cuda_device(device)
    shape = volumes[0].shape
    for v in volumes:
        if v.shape != shape:
            raise ValueError(f"All input volumes must share shape. Got {v.shape} vs {shape}")
    size = int(image_size) if image_size else int(shape[0])
    conds, s0 = [], 0
    for v in volumes:
        c, s0, _ = volume_to_slices(v, slice_half_range, size, device)
        conds.append(c)
    pred = predict_slices_sharded(sample_fn, conds, seed=seed, volume=volume_index, nz=nz, n_time=n_time,
                                  batch=batch, device=device, group=group)      # already in [0, 1]
    return slices_to_volume(pred, shape, s0).cpu().numpy()

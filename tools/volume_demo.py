"""Full-volume prediction demo / check (BASELINE.json configs[2] shape, scaled by --slices / --size):
slices sharded over the ranks, batched per shard, ONE NCCL all-gather.  Prints a checksum that must
be identical for any world size (per-slice RNG streams + batch-invariant kernels).
    python tools/volume_demo.py                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/volume_demo.py
"""
import argparse
import hashlib
import os
import sys
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import mudiff_b200 as M
from mudiff_b200 import volume as V
from mudiff_b200.utils import randomize_

ap = argparse.ArgumentParser()
ap.add_argument('--slices', type=int, default=155)
ap.add_argument('--size', type=int, default=256)
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--volumes', type=int, default=1)
ap.add_argument('--eager', action='store_true', help='eager launches instead of one CUDA graph per batch')
args = ap.parse_args()

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
cfg = Namespace(num_channels=1, num_channels_dae=64, ch_mult=[1, 2, 4], num_res_blocks=2, attn_resolutions=[16],
                dropout=0.0, resamp_with_conv=True, conditional=True, fir=True, fir_kernel=[1, 3, 3, 1],
                skip_rescale=True, resblock_type='biggan', progressive='none', progressive_input='residual',
                progressive_combine='sum', embedding_type='positional', fourier_scale=16.0, not_use_tanh=False,
                image_size=args.size, nz=100, z_emb_dim=256, t_emb_dim=256, n_mlp=3, centered=True, num_timesteps=4,
                beta_min=0.1, beta_max=20.0, use_geometric=False, b200_precision='bf16')
g1 = randomize_(M.NCSNpp(cfg), 0).to(dev).eval()
g2 = randomize_(M.NCSNpp_adaptive(cfg), 1).to(dev).eval()
co = M.Posterior_Coefficients(cfg, dev)


def sample_eager(c, x, z, e):
    return M.sample_from_model(co, g1, c[0], g2, c[1], c[2], 4, x, None, cfg, latents=z, noises=e)


per_rank = (args.slices + world - 1) // world
gbatch = V.balanced_batch(per_rank, args.batch)
sample = sample_eager if args.eager else V.GraphSliceSampler(co, g1, g2, 4, gbatch, args.size, cfg.nz, n_cond=3, device=dev)


gen = torch.Generator().manual_seed(42)
conds = [(torch.randn(args.slices, 1, args.size, args.size, generator=gen).clamp(-3, 3) / 3).pin_memory() for _ in range(3)]
if world > 1:                                   # NCCL connects lazily: keep the one-off channel set-up out of the timed region
    warm = torch.zeros(world * 8, device=dev)
    dist.all_gather_into_tensor(warm, warm[:8].clone())
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
fulls = V.predict_volumes_sharded(sample, [conds] * args.volumes, seed=7, first_volume=0, nz=cfg.nz, n_time=4, batch=gbatch,
                                  device=dev)
full = fulls[-1]
torch.cuda.synchronize()
dt = time.perf_counter() - t0
h = hashlib.sha256(full.cpu().numpy().tobytes()).hexdigest()[:16]
if rank == 0:
    print(f"VOLUME world={world} slices={args.slices} size={args.size} volumes={args.volumes}: {args.volumes * args.slices / dt:.1f} slices/s "
          f"({'eager' if args.eager else 'CUDA graph'}, batch {gbatch}, incl. per-slice RNG + all-gather), checksum(last volume)={h} shape={tuple(full.shape)}", flush=True)
if world > 1:
    dist.destroy_process_group()

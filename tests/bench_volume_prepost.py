"""(Lives under tests/ because it executes the oracle: only tests/, smoke() and bench.py's CPU-baseline leg may.)
Times the volume-prediction front / back end (robust percentile normalisation + centre slices + resize, clamp +
re-stack) for one BraTS-sized case (3 modalities, 240 x 240 x 155 -> 155 slices of 256^2): numpy/ATen restatement of the
reference on the host cores vs the GPU kernels (including the H2D copy of the raw volumes and the D2H of the result)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mudiff_b200 import volume as V
from oracle import volume_oracle as VO

rng = np.random.default_rng(0)
vols = [np.where(rng.random((240, 240, 155)) < 0.55, 0.0, np.round(rng.gamma(2.0, 200.0, (240, 240, 155)))) for _ in range(3)]
fake = torch.randn(155, 1, 240, 240)

t0 = time.perf_counter()
ref = [VO.preprocess_volume(v, 80, 256)[0] for v in vols]
t_pre_cpu = time.perf_counter() - t0
t0 = time.perf_counter()
rb = VO.reconstruct_volume_from_slices(list(VO.postprocess_slices(fake)), vols[0].shape, 0, 154)
t_post_cpu = time.perf_counter() - t0

for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    got = [V.volume_to_slices(v, 80, 256, 'cuda')[0] for v in vols]
    torch.cuda.synchronize(); t_pre_gpu = time.perf_counter() - t0
    fd = fake.cuda(); torch.cuda.synchronize(); t0 = time.perf_counter()
    out = V.slices_to_volume(fd, vols[0].shape, 0, to01=True).cpu()
    torch.cuda.synchronize(); t_post_gpu = time.perf_counter() - t0
err = max((g.cpu() - r).abs().max().item() for g, r in zip(got, ref))
print(f"PREPOST 3 x 240x240x155 -> 155 x 256^2: pre  cpu {t_pre_cpu * 1e3:.0f} ms | gpu {t_pre_gpu * 1e3:.1f} ms (incl. host fp64->fp32 cast + H2D); "
      f"post cpu {t_post_cpu * 1e3:.0f} ms | gpu {t_post_gpu * 1e3:.1f} ms (incl. D2H); max|pre diff| {err:.2e}; "
      f"post equal {bool(np.array_equal(out.numpy(), rb))}; host threads {torch.get_num_threads()}")

"""Probe: does replaying two half-batch graphs on two streams (memory-bound kernels of one overlapping the tensor-core
kernels of the other) beat one full-batch graph?  Timing only (shared scratch buffers race, results are not checked)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import mudiff_b200 as M
from mudiff_b200.utils import randomize_

B = int(os.environ.get('HB', '64'))
args = bench.argparse.Namespace(nf=64, size=256, precision='bf16')
cfg = bench.build_cfg(args)
dev = torch.device('cuda', 0)
g1 = randomize_(M.NCSNpp(cfg), 0).to(dev).eval()
g2 = randomize_(M.NCSNpp_adaptive(cfg), 1).to(dev).eval()
co = M.Posterior_Coefficients(cfg, dev)


def fill(gs):
    gen = torch.Generator(device=dev).manual_seed(7)
    for t in gs.conds:
        t.normal_(generator=gen).clamp_(-3, 3).div_(3)
    gs.x_init.normal_(generator=gen)
    for t in gs.latents + gs.noises:
        t.normal_(generator=gen)


def timeit(fn, n=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


full = M.GraphSampler(co, g1, g2, 4, B, 256, cfg.nz, n_cond=3, device=dev, warmup=1)
fill(full)
t_full = timeit(full.replay)
print(f"one graph  B={B}: {t_full * 1e3:.1f} ms/step  {B / t_full:.1f} slices/s", flush=True)
del full
torch.cuda.empty_cache()
parts = int(os.environ.get('HPARTS', '2'))
hs = [M.GraphSampler(co, g1, g2, 4, B // parts, 256, cfg.nz, n_cond=3, device=dev, warmup=1) for _ in range(parts)]
for h in hs:
    fill(h)
streams = [torch.cuda.Stream(device=dev) for _ in range(parts)]


def both():
    for h, s in zip(hs, streams):
        with torch.cuda.stream(s):
            h.graph.replay()


t_seq = timeit(lambda: [h.replay() for h in hs])
print(f"{parts} graphs B={B // parts} one stream: {t_seq * 1e3:.1f} ms/step  {B / t_seq:.1f} slices/s", flush=True)
t_par = timeit(both)
print(f"{parts} graphs B={B // parts} {parts} streams: {t_par * 1e3:.1f} ms/step  {B / t_par:.1f} slices/s", flush=True)

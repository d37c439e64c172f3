"""Micro-benchmark of the fused AdaGN + SiLU + FIR resample kernel (mudiff_upfirdn2d_gn) on the four shapes of the
bench workload (B = 64, nf = 64, 256^2): achieved HBM GB/s = (read x + write FIR(h) + write FIR(x)) / time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mudiff_b200 as M
from mudiff_b200 import ops, up_or_down_sampling as U

B = int(os.environ.get('B', '64'))
dev = torch.device('cuda', 0)
shapes = [(64, 256, False), (128, 128, False), (256, 64, True), (128, 128, True)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for C, S, up in shapes:
    x = ops.as_nhwc(torch.randn(B, C, S, S, device=dev).to(torch.bfloat16))
    table = torch.stack([1.0 + 0.2 * torch.randn(B, C), 0.3 * torch.randn(B, C)], -1).contiguous().to(dev)
    for _ in range(3):
        h, xr = U.resample_2d_gn(x, table, [1, 3, 3, 1], up=up)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h, xr = U.resample_2d_gn(x, table, [1, 3, 3, 1], up=up)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    gb = (x.numel() + h.numel() + xr.numel()) * 2 / 1e9
    print(f"C={C:3d} {S}^2 {'up  ' if up else 'down'}: {t * 1e3:8.1f} us  {gb:6.3f} GB  {gb / t:6.2f} TB/s = {100 * gb / t / 6.54:5.1f}% of the 6.54 TB/s HBM copy peak")

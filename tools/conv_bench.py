"""Micro-benchmark of conv_tc variants (ablation flags) on representative generator shapes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mudiff_b200 as M
from mudiff_b200 import ops

B = int(os.environ.get('CB_BATCH', '16'))
SHAPES = [  # (H, Cin list, taps, N)
    (256, [64], [9], 64),
    (256, [256], [9], 64),
    (256, [64, 128, 64], [9, 1, 1], 64),
    (128, [128], [9], 128),
    (128, [256, 128], [9, 9], 128),
    (64, [256], [9], 256),
    (64, [512], [9], 256),
    (256, [192], [9], 384),
    (256, [320], [9], 64),
    (256, [192], [9], 64),
    (256, [128], [9], 64),
    (256, [256, 64], [9, 9], 64),
    (256, [384], [9], 64),
    (256, [448], [9], 64),
    (256, [512], [9], 64),
    (128, [320], [9], 128),
    (128, [384], [9], 128),
]
if os.environ.get('CB_VARIANTS') == 'short':
    VARIANTS = [('default', 0), ('dry', 64), ('dry+noepi', 192), ('mt1', 16), ('mt1+dry+noepi', 16 + 192), ('nohalo', 2), ('nohalo+dry+noepi', 2 + 192)]
else:
  VARIANTS = [('default', 0), ('pair+dry', 64), ('pair+noepi', 128), ('pair+dry+noepi', 192), ('single', 0x4000), ('single+nostore', 0x4000 + 2048), ('single+noldtm', 0x4000 + 4096), ('single+noepi', 0x4000 + 128), ('single+dry', 0x4000 + 64), ('single+dry+noepi', 0x4000 + 192), ('single+mt1', 0x4000 + 16)]
if os.environ.get('CB_SHAPES'):
    SHAPES = [SHAPES[int(i)] for i in os.environ['CB_SHAPES'].split(',')]


def bench(fn, iters=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for H, cins, taps, N in SHAPES:
    segs, ws = [], []
    for ci, tp in zip(cins, taps):
        x = torch.randn(B, ci, H, H, device='cuda').to(torch.bfloat16)
        k = 3 if tp == 9 else 1
        w = (torch.randn(N, ci, k, k, device='cuda') / (ci * tp) ** 0.5).to(torch.bfloat16)
        segs.append((ops.as_nhwc(x), tp)); ws.append(ops.pack_conv_weight(w, (ci,), torch.bfloat16))
    wt = torch.cat(ws, 1).contiguous()
    ktot = wt.shape[1]
    flops = 2.0 * B * H * H * N * ktot
    out = ops.empty_nhwc(B, N, H, H, torch.bfloat16, 'cuda')
    line = f"H={H:3d} Cin={cins} N={N:3d} K={ktot:5d}: "
    for name, fl in VARIANTS:
        for st in (0,):
            try:
                ms = bench(lambda: ops.conv(segs, wt, N, out=out, flags=fl, force='tc', want_stats=bool(st)))
                line += f"{name}{'+stats' if st else ''}={flops / ms / 1e9:6.0f}  "
            except RuntimeError as e:
                line += f"{name}=ERR  "
    print(line, flush=True)

"""GPU bring-up diagnostics for the tcgen05 conv kernel: every case runs in its own
subprocess (a device trap in one case must not poison the rest).  Usage on the GPU box:
    python tools/gpu_diag.py            # all cases
    python tools/gpu_diag.py CASE       # one case (internal)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # flags: 2 = no halo staging, 8 = no stationary weights, 16 = one pixel tile per unit
    'gemm_1x1_n64':        dict(B=2, H=32, W=32, C=[64], taps=[1], N=64, flags=0),
    'conv3_pertap_n64':    dict(B=2, H=32, W=32, C=[64], taps=[9], N=64, flags=2),
    'conv3_halo_stat_n64': dict(B=2, H=32, W=32, C=[64], taps=[9], N=64, flags=0),
    'conv3_halo_strm_n64': dict(B=2, H=32, W=32, C=[64], taps=[9], N=64, flags=8),
    'conv3_halo_mt1_n64':  dict(B=2, H=32, W=32, C=[64], taps=[9], N=64, flags=8 | 16),
    'conv3_halo_n128':     dict(B=3, H=40, W=24, C=[128], taps=[9], N=128, flags=0),
    'conv3_halo_n256':     dict(B=1, H=32, W=32, C=[256], taps=[9], N=256, flags=0),
    'conv3_c320_n64':      dict(B=1, H=48, W=40, C=[256, 64], taps=[9, 9], N=64, flags=0),
    'fused_shortcut':      dict(B=2, H=32, W=32, C=[64, 128, 64], taps=[9, 1, 1], N=64, flags=0, epi=True),
    'n384_sigmoid':        dict(B=1, H=32, W=32, C=[192], taps=[9], N=384, flags=0, act=2),
    'gemm_mode_h1':        dict(B=2, H=1, W=256, C=[128], taps=[1], N=256, flags=0),
    'g_w4096_n256':        dict(B=2, H=1, W=4096, C=[256], taps=[1], N=256, flags=0),
    'g_w512_n4096':        dict(B=4, H=1, W=512, C=[256], taps=[1], N=4096, flags=0),
    'g_w4096_n4096_dry':   dict(B=2, H=1, W=4096, C=[256], taps=[1], N=4096, flags=64),
    'g_w4096_n4096_noepi': dict(B=2, H=1, W=4096, C=[256], taps=[1], N=4096, flags=128),
    'g_w4096_n4096_rr':    dict(B=2, H=1, W=4096, C=[256], taps=[1], N=4096, flags=32),
    'c_mt1_multi':         dict(B=4, H=128, W=128, C=[128], taps=[9], N=128, flags=16),
    'c_n256_multi':        dict(B=8, H=64, W=64, C=[128], taps=[9], N=256, flags=0),
    'v_b1':                dict(B=1, H=1, W=4096, C=[256], taps=[1], N=4096, flags=0),
    'v_n1024':             dict(B=2, H=1, W=4096, C=[256], taps=[1], N=1024, flags=0),
    'v_n512':              dict(B=2, H=1, W=4096, C=[256], taps=[1], N=512, flags=0),
    'v_w2048':             dict(B=2, H=1, W=2048, C=[256], taps=[1], N=2048, flags=0),
    'v_k64':               dict(B=2, H=1, W=4096, C=[64], taps=[1], N=4096, flags=0),
    'v_b1_st1':            dict(B=1, H=1, W=4096, C=[256], taps=[1], N=4096, flags=256),
    'v_b1_a4':             dict(B=1, H=1, W=4096, C=[256], taps=[1], N=4096, flags=512),
    'v_b1_k128':           dict(B=1, H=1, W=4096, C=[128], taps=[1], N=4096, flags=0),
    'gemm_ntiles4':        dict(B=1, H=1, W=1024, C=[256], taps=[1], N=1024, flags=0),
    'gemm_ntiles16_b2':    dict(B=2, H=1, W=4096, C=[256], taps=[1], N=4096, flags=0),
    'conv_multiunit_bias': dict(B=4, H=128, W=128, C=[64], taps=[9], N=64, flags=0, epi=True),
    'many_tiles':          dict(B=8, H=64, W=64, C=[64], taps=[9], N=64, flags=0),
    'many_tiles_n128':     dict(B=8, H=64, W=64, C=[128], taps=[9], N=128, flags=0),
}


def run_case(name):
    import torch
    import torch.nn.functional as F
    import mudiff_b200 as M
    from mudiff_b200 import ops
    c = CASES[name]
    print('   debug channel selftest:', M._lib.lib().mudiff_debug_selftest(), flush=True)
    torch.manual_seed(0)
    dev = 'cuda'
    B, H, W, N = c['B'], c['H'], c['W'], c['N']
    segs, ws, ref = [], [], 0
    for ci, taps in zip(c['C'], c['taps']):
        x = torch.randn(B, ci, H, W, device=dev).to(torch.bfloat16)
        k = 3 if taps == 9 else 1
        w = (torch.randn(N, ci, k, k, device=dev) / (ci * taps) ** 0.5).to(torch.bfloat16)
        ref = ref + F.conv2d(x.float(), w.float(), padding=k // 2)
        segs.append((ops.as_nhwc(x), taps))
        ws.append(ops.pack_conv_weight(w, (ci,), torch.bfloat16))
    wt = torch.cat(ws, dim=1).contiguous()
    kw = {}
    if c.get('epi'):
        bias = torch.randn(N, device=dev)
        rowbias = torch.randn(B, N, device=dev)
        res = torch.randn(B, N, H, W, device=dev)
        ref = 0.7 * (ref + bias[None, :, None, None] + rowbias[:, :, None, None]) + 0.3 * res
        kw = dict(bias=bias, rowbias=rowbias, residual=ops.as_nhwc(res), alpha=0.7, beta=0.3)
    if c.get('act') == 2:
        ref = torch.sigmoid(ref)
        kw['act'] = 2
    torch.cuda.synchronize()
    print('   launching f32-out conv', flush=True)
    out = ops.conv(segs, wt, N, out_dtype=torch.float32, flags=c['flags'], force='tc', **kw)
    print('   launched', flush=True)
    torch.cuda.synchronize()
    print('   f32-out conv done', flush=True)
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    # bf16 output too
    kw2 = dict(kw)
    if 'residual' in kw2:
        kw2['residual'] = kw2['residual'].to(torch.bfloat16)
    out16 = ops.conv(segs, wt, N, out_dtype=torch.bfloat16, flags=c['flags'], force='tc', **kw2)
    torch.cuda.synchronize()
    err16 = (out16.float() - ref).abs().max().item()
    ok = err < 5e-3 * max(scale, 1.0)
    print(f"CASE {name}: max|err| f32-out {err:.3e}  bf16-out {err16:.3e}  (|ref|max {scale:.3f})  {'OK' if ok else 'MISMATCH'}")
    if not ok:
        d = (out - ref).abs()
        bad = (d > 1e-2 * max(scale, 1.0))
        print("   bad fraction", bad.float().mean().item(), " per-channel bad frac (first 8):",
              bad.float().mean(dim=(0, 2, 3))[:8].tolist())
        print("   per-row(y) bad frac:", [round(v, 2) for v in bad.float().mean(dim=(0, 1, 3))[:16].tolist()])
        print("   per-col(x) bad frac:", [round(v, 2) for v in bad.float().mean(dim=(0, 1, 2))[:16].tolist()])
    return 0 if ok else 1


if __name__ == '__main__':
    if len(sys.argv) > 1:
        try:
            rc = run_case(sys.argv[1])
        except Exception as e:
            import ctypes
            import mudiff_b200 as M
            info = (ctypes.c_int32 * 8)()
            M._lib.lib().mudiff_debug_last_timeout(info)
            print('EXCEPTION', type(e).__name__, str(e).splitlines()[0], ' timeout-info', list(info), flush=True)
            rc = 2
        sys.exit(rc)
    fails = 0
    for name in CASES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=180)
            lines = [l for l in (r.stdout + r.stderr).splitlines() if l.strip()]
            tail = [l for l in lines if l.startswith('CASE') or l.startswith('   ')]
            if r.returncode != 0 and not tail:
                tail = lines[-6:]
            print('\n'.join(tail) if tail else f'CASE {name}: no output rc={r.returncode}')
            if r.returncode != 0:
                fails += 1
                print(f'   rc={r.returncode}')
        except subprocess.TimeoutExpired:
            fails += 1
            print(f'CASE {name}: TIMEOUT')
        print(f'   ({time.time() - t0:.1f}s)', flush=True)
    print('DIAG FAILS', fails)

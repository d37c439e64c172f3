#!/usr/bin/env python
"""Reference GPU baseline (the denominator of the north star's ">= 25x"): the UNMODIFIED reference modules
(baseline/_ref, see baseline/ref_harness.py), its own sampling loop (engine/test.py:180-199) and its own CUDA
extensions, eager PyTorch on one B200.

    python tools/ref_gpu_baseline.py [--out profiles/r02_reference_gpu.json] [--quick]

Rows: fp16 autocast (what engine/test.py:191 does) at B=1 and B=64, fp32 with TF32 off at B=1 and B=16, TF32 on at B=64.
Also a parity row: this repo's fp32 and bf16 paths against the reference's fp32 GPU output on the same weights, inputs
and noise at 256^2 (the reference itself as the checker, on the box).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parity(R, size=256, batch=2):
    import torch
    from argparse import Namespace
    import mudiff_b200 as M
    from oracle import mudiff_oracle as O
    dev = torch.device('cuda:0')
    cfg = R.reference_config(64, size)
    ocfg = O.default_config(num_channels_dae=64, image_size=size)
    sd1, sd2 = O.make_state_dict(ocfg, 'g1', seed=0), O.make_state_dict(ocfg, 'g2', seed=1)
    conds, x_init, latents, noises = O.synthetic_inputs(batch, size, ocfg, seed=42)
    to = lambda ts: [t.to(dev) for t in ts]
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        g1, g2 = R.build_models(cfg, dev, state_dicts=(sd1, sd2))
        ref = R.run_loop(R.engine_symbols(), cfg, g1, g2, to(conds), x_init.to(dev), to(latents), to(noises)).float()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out = {"size": size, "batch": batch, "checker": "reference fp32 on the GPU (TF32 off), same weights / inputs / noise"}
    for prec in ('fp32', 'bf16'):
        ns = Namespace(**vars(ocfg), b200_precision=prec)
        m1, m2 = M.NCSNpp(ns).to(dev).eval(), M.NCSNpp_adaptive(ns).to(dev).eval()
        m1.load_state_dict(sd1)
        m2.load_state_dict(sd2)
        co = M.Posterior_Coefficients(ns, dev)
        c = to(conds)
        y = M.sample_from_model(co, m1, c[0], m2, c[1], c[2], ocfg.num_timesteps, x_init.to(dev), None, ns,
                                latents=to(latents), noises=to(noises))
        torch.cuda.synchronize()
        out[prec] = {"max_abs": (y - ref).abs().max().item(), "rel_l2": ((y - ref).norm() / ref.norm()).item()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'reference_gpu.json'))
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    args = ap.parse_args()
    import torch
    from baseline import ref_harness as R
    assert torch.cuda.is_available() and R.available()
    rows = []
    plan = [('fp16', 1, 10), ('fp16', 64, 3), ('fp32', 1, 10), ('fp32', 16, 3), ('tf32', 64, 3)]
    if args.quick:
        plan = [('fp16', 1, 5), ('fp16', 64, 2)]
    cfg = R.reference_config(64, 256)
    models = R.build_models(cfg, torch.device('cuda:0'))
    for mode, b, it in plan:
        try:
            r = R.time_gpu(64, 256, b, mode, iters=it, warmup=3, models=models)
        except Exception as e:                              # noqa: BLE001  (e.g. OOM at a large batch: report, go on)
            r = {"mode": mode, "batch": b, "error": str(e).splitlines()[0][:200]}
            torch.cuda.empty_cache()
        rows.append(r)
        print(json.dumps(r), flush=True)
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "rows": rows,
           "what": "reference (MarioPasc/MU-Diff) NCSNpp + NCSNpp_adaptive, engine/test.py sample_from_model, 4 steps, "
                   "256^2, nf=64, eager, its own upfirdn2d CUDA extension; CUDA events, median of iters after 3 warm-ups"}
    if not args.no_parity:
        del models
        torch.cuda.empty_cache()
        res["parity_vs_reference_gpu"] = parity(R)
        print(json.dumps(res["parity_vs_reference_gpu"]), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(res, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()

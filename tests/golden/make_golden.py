"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

    PYTHONPATH=/root/reference TORCH_EXTENSIONS_DIR=/tmp/ref_ext \
        python tests/golden/make_golden.py [main|healthy]

The two generator files of the reference cannot be imported in one process
(both register 'ncsnpp'), hence the `main` / `healthy` argument.
/root/reference does not exist on the GPU box: the outputs are committed fixtures.

What is stored (all float32, all small):
  fir.npz      upfirdn2d_native (utils/op/upfirdn2d.py:201) on the shapes/pads the
               generators use + odd cases (negative pads, up=3/down=2, 5x3 kernel);
               fused_leaky_relu CPU branch (utils/op/fused_act.py:113-120)
  posterior.npz  Posterior_Coefficients tables (engine/test.py:101-123) - the engine
               file is not importable here (skimage/matplotlib), so the class body is
               exec'd from its source text - and sample_posterior_combine outputs
  gen_<variant>.npz  reference NCSNpp / NCSNpp_adaptive forward on the oracle's
               deterministic state_dict (loaded with strict=True => key/shape parity),
               nf=64 @ 32^2 B=2, nf=16 @ 64^2 B=1; plus a full 4-step sample
"""
import ast
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'
sys.path.insert(0, REF)

from oracle import mudiff_oracle as O  # noqa: E402

torch.set_num_threads(8)


def ns(cfg):
    return Namespace(**vars(cfg))


def load_engine_symbols():
    """exec the pure-torch part of engine/test.py (:48-177) without importing the module."""
    src = open(os.path.join(REF, 'engine/test.py')).read()
    tree = ast.parse(src)
    keep = {'var_func_vp', 'var_func_geometric', 'extract', 'get_time_schedule', 'get_sigma_schedule',
            'Posterior_Coefficients', 'sample_posterior_combine', 'sample_from_model'}
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in keep]
    mod = types.ModuleType('ref_engine_test')
    from contextlib import nullcontext
    mod.__dict__.update(dict(torch=torch, np=np, autocast=nullcontext))
    exec(compile(ast.Module(body=body, type_ignores=[]), 'engine/test.py', 'exec'), mod.__dict__)
    return mod


def gen_fir():
    from utils.op.upfirdn2d import upfirdn2d_native
    from utils.op.fused_act import fused_leaky_relu
    g = torch.Generator().manual_seed(1)
    out = {}
    cases = [  # (name, shape, kernel, up, down, (px0,px1,py0,py1))
        ('down2', (2, 8, 16, 16), O.setup_kernel([1, 3, 3, 1]), 1, 2, (1, 1, 1, 1)),
        ('up2', (2, 8, 8, 8), O.setup_kernel([1, 3, 3, 1]) * 4, 2, 1, (2, 1, 2, 1)),
        ('pre', (1, 4, 16, 16), O.setup_kernel([1, 3, 3, 1]), 1, 1, (2, 2, 2, 2)),
        ('negpad', (1, 3, 12, 10), O.setup_kernel([1, 2, 1]), 1, 1, (-1, 2, 1, -2)),
        ('up3down2', (1, 2, 7, 9), O.setup_kernel([1, 4, 6, 4, 1]), 3, 2, (3, 2, 3, 2)),
        ('odd', (1, 1, 5, 5), O.setup_kernel([1, 1]), 2, 1, (0, 0, 0, 0)),
    ]
    for name, shape, k, up, down, pad in cases:
        x = torch.randn(*shape, generator=g)
        kt = torch.tensor(k, dtype=torch.float32)
        y = upfirdn2d_native(x, kt, up, up, down, down, *pad)
        out[f'{name}_x'], out[f'{name}_k'], out[f'{name}_y'] = x.numpy(), kt.numpy(), y.numpy()
        out[f'{name}_p'] = np.array([up, down, *pad], dtype=np.int64)
    # asymmetric kernel (tests the flip) with per-axis up/down through upfirdn2d_native directly
    x = torch.randn(1, 2, 9, 11, generator=g)
    kt = torch.randn(3, 5, generator=g)
    y = upfirdn2d_native(x, kt, 2, 1, 1, 2, 1, 3, 0, 2)
    out['asym_x'], out['asym_k'], out['asym_y'] = x.numpy(), kt.numpy(), y.numpy()
    out['asym_p'] = np.array([2, 1, 1, 2, 1, 3, 0, 2], dtype=np.int64)   # up_x up_y down_x down_y pads
    x = torch.randn(2, 6, 5, 7, generator=g)
    b = torch.randn(6, generator=g)
    out['lrelu_x'], out['lrelu_b'] = x.numpy(), b.numpy()
    out['lrelu_y'] = fused_leaky_relu(x, b).numpy()            # CPU branch: slope 0.2, scale sqrt2
    np.savez_compressed(os.path.join(HERE, 'fir.npz'), **out)
    print('fir.npz', len(out))


def gen_posterior():
    E = load_engine_symbols()
    cfg = O.default_config()
    pc = E.Posterior_Coefficients(ns(cfg), torch.device('cpu'))
    out = dict(betas=pc.betas.numpy(), coef1=pc.posterior_mean_coef1.numpy(),
               coef2=pc.posterior_mean_coef2.numpy(), var=pc.posterior_variance.numpy(),
               log_var=pc.posterior_log_variance_clipped.numpy())
    g = torch.Generator().manual_seed(2)
    x01, x02, xt, noise = (torch.randn(4, 1, 8, 8, generator=g) for _ in range(4))
    t = torch.tensor([3, 2, 1, 0], dtype=torch.int64)
    # the reference draws its own noise (engine/test.py:169): patch randn_like for this call
    orig = torch.randn_like
    torch.randn_like = lambda *_a, **_k: noise
    try:
        y = E.sample_posterior_combine(pc, x01, x02, xt, t)
    finally:
        torch.randn_like = orig
    out.update(x01=x01.numpy(), x02=x02.numpy(), xt=xt.numpy(), noise=noise.numpy(), t=t.numpy(), y=y.numpy())
    np.savez_compressed(os.path.join(HERE, 'posterior.npz'), **out)
    print('posterior.npz')


def gen_generators(which):
    E = load_engine_symbols()
    if which == 'main':
        from backbones.ncsnpp_generator_adagn_feat import NCSNpp, NCSNpp_adaptive
        v1, v2, ncond = 'g1', 'g2', 3
    else:
        from backbones.ncsnpp_generator_adagn_feat_healthy import NCSNpp, NCSNpp_adaptive
        v1, v2, ncond = 'g1_healthy', 'g2_healthy', 2
    out = {}
    # known-answer: parameter counts of the production config (error_logs/...out:116)
    cfg = O.default_config()
    G1, G2 = NCSNpp(ns(cfg)), NCSNpp_adaptive(ns(cfg))
    out['nparams'] = np.array([sum(p.numel() for p in G1.parameters()), sum(p.numel() for p in G2.parameters())])
    print('param counts', out['nparams'])
    del G1, G2
    for tag, nf, size, batch in (('nf64_s32', 64, 32, 2), ('nf16_s64', 16, 64, 1)):
        cfg = O.default_config(num_channels_dae=nf, image_size=size)
        sd1, sd2 = O.make_state_dict(cfg, v1, seed=0), O.make_state_dict(cfg, v2, seed=1)
        G1, G2 = NCSNpp(ns(cfg)).eval(), NCSNpp_adaptive(ns(cfg)).eval()
        G1.load_state_dict(sd1, strict=True)
        G2.load_state_dict(sd2, strict=True)
        conds, x_init, latents, noises = O.synthetic_inputs(batch, size, cfg, ncond=ncond, seed=42)
        t = torch.tensor([3, 1][:batch], dtype=torch.int64)
        with torch.no_grad():
            y1 = G1(x_init, *conds, t, latents[0])
            y2 = G2(x_init, *conds, t, latents[0], y1[:, [0], :])
        out[f'{tag}_g1'], out[f'{tag}_g2'] = y1.numpy(), y2.numpy()
        # cross-check the oracle right here
        o1 = O.generator_forward(sd1, cfg, v1, x_init, conds, t, latents[0])
        o2 = O.generator_forward(sd2, cfg, v2, x_init, conds, t, latents[0], pseudo_target=y1[:, [0], :])
        print(tag, 'oracle-vs-reference max|d|', (o1 - y1).abs().max().item(), (o2 - y2).abs().max().item(),
              ' |y| max', y1.abs().max().item(), y2.abs().max().item())
        if which == 'main':
            # full 4-step loop through the reference's own sample_from_model (engine/test.py:180-199);
            # its RNG draws are redirected to the pre-drawn tensors, in the reference's draw order.
            pc = E.Posterior_Coefficients(ns(cfg), torch.device('cpu'))
            queue = []
            for i in reversed(range(cfg.num_timesteps)):
                queue += [latents[i], noises[i]]
            o_randn, o_like = torch.randn, torch.randn_like
            torch.randn = lambda *a, **k: queue.pop(0)
            torch.randn_like = lambda *a, **k: queue.pop(0)
            try:
                xs = E.sample_from_model(pc, G1, conds[0], G2, conds[1], conds[2], cfg.num_timesteps, x_init, None, ns(cfg))
            finally:
                torch.randn, torch.randn_like = o_randn, o_like
            assert not queue
            out[f'{tag}_sample'] = xs.numpy()
            co = O.PosteriorCoefficients(cfg)
            xo = O.sample_from_model(co, sd1, sd2, cfg, conds, x_init, latents, noises)
            print(tag, 'loop oracle-vs-reference max|d|', (xo - xs).abs().max().item(), '|x| max', xs.abs().max().item())
    np.savez_compressed(os.path.join(HERE, f'gen_{which}.npz'), **out)
    print(f'gen_{which}.npz')


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'main'
    if which == 'main':
        gen_fir()
        gen_posterior()
    gen_generators(which)

"""CPU oracle for the MU-Diff reverse-sampling hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32) restatement of the reference algorithm for
the path named by BASELINE.json `north_star`.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the product
package (`mu-diff_b200/`) never does.

Parity pin: the reference has no tests or golden vectors of its own, so
this restatement is pinned against outputs of the reference itself, produced in the
build container by `tests/golden/make_golden.py` (imports /root/reference, loads the
state_dict made by `make_state_dict` below with strict=True, runs the reference
modules, stores the outputs in tests/golden/*.npz), against the parameter-count
known-answers of `error_logs/log_mudiff_T1.13967221.out:116`, and against the
posterior tables recomputed from the reference formulas.
The conv / GEMM / GroupNorm / softmax arithmetic itself lives in PyTorch ATen (pinned
torch==2.4.1 in the reference's requirement.txt:1; 2.11.0 here) - third-party, not
under /root/reference - and is called here exactly at the reference's call sites.

Reference files followed (all paths relative to /root/reference):
  engine/test.py:48-63,75-97,101-123,150-199        schedules, posterior, sampling loop
  backbones/ncsnpp_generator_adagn_feat.py:279-447   NCSNpp.forward            (G1)
  backbones/ncsnpp_generator_adagn_feat.py:694-905   NCSNpp_adaptive.forward   (G2)
  backbones/ncsnpp_generator_adagn_feat_healthy.py:279,693   2-contrast variants
  backbones/layerspp.py:37-54,98-137,176-210,261-324,394-501   blocks
  backbones/layers.py:465-479,496-505                timestep embedding, NIN
  backbones/up_or_down_sampling.py:149-262           FIR front-end
  utils/op/upfirdn2d.py:170-242, utils/op/fused_act.py:112-123   native-op CPU paths
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------
def default_config(**over) -> SimpleNamespace:
    """README.md:85 / demo notebook cell 3 inference configuration."""
    cfg = dict(num_channels=1, num_channels_dae=64, ch_mult=[1, 2, 4], num_res_blocks=2,
               attn_resolutions=[16], dropout=0.0, resamp_with_conv=True, conditional=True,
               fir=True, fir_kernel=[1, 3, 3, 1], skip_rescale=True, resblock_type='biggan',
               progressive='none', progressive_input='residual', progressive_combine='sum',
               embedding_type='positional', fourier_scale=16.0, not_use_tanh=False,
               image_size=256, nz=100, z_emb_dim=256, t_emb_dim=256, n_mlp=3, centered=True,
               num_timesteps=4, beta_min=0.1, beta_max=20.0, use_geometric=False)
    cfg.update(over)
    return SimpleNamespace(**cfg)


def _groups(c: int) -> int:
    return min(c // 4, 32)          # layerspp.py:268,281,103


# ----------------------------------------------------------------------------------
# architecture spec: the order of `all_modules` (ncsnpp_generator_adagn_feat.py:87-269)
# ----------------------------------------------------------------------------------
def build_spec(cfg, variant: str) -> List[dict]:
    """List of module descriptors in the order the reference appends them to
    `all_modules` for the default flags (positional embedding, conditional, biggan
    blocks, progressive='none', progressive_input='residual', fir, resamp_with_conv).

    variant: 'g1' | 'g2' | 'g1_healthy' | 'g2_healthy'
    """
    assert cfg.embedding_type == 'positional' and cfg.conditional
    assert cfg.resblock_type == 'biggan' and cfg.progressive == 'none'
    assert cfg.progressive_input == 'residual' and cfg.fir and cfg.resamp_with_conv
    nf, ch = cfg.num_channels_dae, cfg.num_channels
    nres = cfg.num_res_blocks
    nlev = len(cfg.ch_mult)
    res = [cfg.image_size // (2 ** i) for i in range(nlev)]
    healthy = variant.endswith('healthy')
    adaptive = variant.startswith('g2')
    ncond = 2 if healthy else 3

    spec: List[dict] = [dict(kind='linear', cin=nf, cout=4 * nf),
                        dict(kind='linear', cin=4 * nf, cout=4 * nf)]
    if not adaptive:                                    # :177-180 (healthy: :177-179)
        for _ in range(1 + ncond):
            spec.append(dict(kind='feat', cin=ch, cout=nf))
        stem_c = nf * (1 + ncond)
    else:                                               # :578-582 (healthy: :577-580)
        spec.append(dict(kind='gap', cin=ch, cout=nf))
        spec.append(dict(kind='feat', cin=ch, cout=nf))
        for _ in range(ncond):
            spec.append(dict(kind='adafeat', cin=ch, cout=nf))
        stem_c = nf * 4 if not healthy else nf * 2      # :584 / healthy :585

    hs_c = [stem_c]
    in_ch = stem_c
    pyr_ch = ch
    for lv in range(nlev):
        for _ in range(nres):
            out_ch = nf * cfg.ch_mult[lv]
            spec.append(dict(kind='res', cin=in_ch, cout=out_ch, up=False, down=False))
            in_ch = out_ch
            if res[lv] in cfg.attn_resolutions:
                spec.append(dict(kind='attn', c=in_ch))
            hs_c.append(in_ch)
        if lv != nlev - 1:
            spec.append(dict(kind='res', cin=in_ch, cout=in_ch, up=False, down=True))
            spec.append(dict(kind='pyrdown', cin=pyr_ch, cout=in_ch))
            pyr_ch = in_ch
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    spec.append(dict(kind='res', cin=in_ch, cout=in_ch, up=False, down=False))
    spec.append(dict(kind='attn', c=in_ch))
    spec.append(dict(kind='res', cin=in_ch, cout=in_ch, up=False, down=False))
    for lv in reversed(range(nlev)):
        for _ in range(nres + 1):
            out_ch = nf * cfg.ch_mult[lv]
            spec.append(dict(kind='res', cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False))
            in_ch = out_ch
        if res[lv] in cfg.attn_resolutions:
            spec.append(dict(kind='attn', c=in_ch))
        if lv != 0:
            spec.append(dict(kind='res', cin=in_ch, cout=in_ch, up=True, down=False))
    assert not hs_c
    spec.append(dict(kind='gn', c=in_ch))
    spec.append(dict(kind='conv3', cin=in_ch, cout=ch))
    return spec


# ----------------------------------------------------------------------------------
# deterministic, NON-degenerate weights (SURVEY.md §0.5, §8c hygiene)
# ----------------------------------------------------------------------------------
def make_state_dict(cfg, variant: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    """A state_dict with exactly the reference's keys/shapes.  Every weight is
    fan-avg uniform at scale 1 (including the tensors the reference initialises at
    1e-10), every bias is N(0, 0.1^2), AdaGN style biases are [1..1, 0..0] + N(0, 0.1^2).
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    nf = cfg.num_channels_dae

    def uni(*shape, fan_in, fan_out):
        a = math.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
        return (torch.rand(*shape, generator=g) * 2.0 - 1.0) * a

    def bias(n):
        return torch.randn(n, generator=g) * 0.1

    def conv(key, cin, cout, k):
        sd[key + '.weight'] = uni(cout, cin, k, k, fan_in=cin * k * k, fan_out=cout * k * k)
        sd[key + '.bias'] = bias(cout)

    def lin(key, cin, cout):
        sd[key + '.weight'] = uni(cout, cin, fan_in=cin, fan_out=cout)
        sd[key + '.bias'] = bias(cout)

    def style(key, sdim, c):
        sd[key + '.weight'] = uni(2 * c, sdim, fan_in=sdim, fan_out=2 * c)
        b = bias(2 * c)
        b[:c] += 1.0
        sd[key + '.bias'] = b

    for i, m in enumerate(build_spec(cfg, variant)):
        p = f'all_modules.{i}'
        k = m['kind']
        if k == 'linear':
            lin(p, m['cin'], m['cout'])
        elif k in ('feat', 'gap', 'adafeat'):
            conv(p + '.conv1', m['cin'], m['cout'], 3)
            if k == 'adafeat':
                style(p + '.group_norm.style', 256, m['cout'])      # layerspp.py:427,434
            conv(p + '.conv2', m['cout'], m['cout'], 3)
            if k == 'gap':
                lin(p + '.fc', m['cout'], 256)                      # layerspp.py:459,476
        elif k == 'res':
            style(p + '.GroupNorm_0.style', cfg.z_emb_dim, m['cin'])
            conv(p + '.Conv_0', m['cin'], m['cout'], 3)
            lin(p + '.Dense_0', 4 * nf, m['cout'])
            style(p + '.GroupNorm_1.style', cfg.z_emb_dim, m['cout'])
            conv(p + '.Conv_1', m['cout'], m['cout'], 3)
            if m['cin'] != m['cout'] or m['up'] or m['down']:
                conv(p + '.Conv_2', m['cin'], m['cout'], 1)
        elif k == 'attn':
            c = m['c']
            sd[p + '.GroupNorm_0.weight'] = 1.0 + bias(c)
            sd[p + '.GroupNorm_0.bias'] = bias(c)
            for j in range(4):
                sd[p + f'.NIN_{j}.W'] = uni(c, c, fan_in=c, fan_out=c)
                sd[p + f'.NIN_{j}.b'] = bias(c)
        elif k == 'pyrdown':
            conv(p + '.Conv2d_0', m['cin'], m['cout'], 3)
        elif k == 'gn':
            sd[p + '.weight'] = 1.0 + bias(m['c'])
            sd[p + '.bias'] = bias(m['c'])
        elif k == 'conv3':
            conv(p, m['cin'], m['cout'], 3)
        else:
            raise AssertionError(k)

    if variant.startswith('g2'):
        ncond = 2 if variant.endswith('healthy') else 3
        names_w = ['c1', 'c2', 'c3'] if ncond == 3 else ['c1']
        names_a = ['c12', 'c23', 'c31'] if ncond == 3 else ['c12']
        for n in names_w:
            conv(f'feat_weight_{n}', nf, nf, 3)
        for n in names_a:
            conv(f'feat_att1_{n}', ncond * nf, nf, 3)
            conv(f'feat_att2_{n}', ncond * nf, nf, 3)

    lin('z_transform.1', cfg.nz, cfg.z_emb_dim)
    for j in range(cfg.n_mlp):
        lin(f'z_transform.{3 + 2 * j}', cfg.z_emb_dim, cfg.z_emb_dim)
    return sd


def param_count(sd: Dict[str, torch.Tensor]) -> int:
    return int(sum(v.numel() for v in sd.values()))


# ----------------------------------------------------------------------------------
# native-op CPU paths
# ----------------------------------------------------------------------------------
def setup_kernel(k: Sequence[float]) -> np.ndarray:
    """up_or_down_sampling.py:186-193: separable taps -> normalised 2-D kernel."""
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k = k / np.sum(k)
    assert k.ndim == 2 and k.shape[0] == k.shape[1]
    return k


def upfirdn2d_ref(x: torch.Tensor, kernel: torch.Tensor, up_x: int, up_y: int, down_x: int,
                  down_y: int, px0: int, px1: int, py0: int, py1: int) -> torch.Tensor:
    """Restates utils/op/upfirdn2d.py:201-242 (`upfirdn2d_native`): zero-insert upsample,
    pad (negative pads crop), true convolution with `kernel` (i.e. correlation with the
    flipped kernel, :228), keep every `down`-th sample.  Written tap-by-tap instead of
    through F.conv2d so that the CUDA kernel has an independent check."""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    up = x.new_zeros(n, c, h * up_y, w * up_x)
    up[:, :, ::up_y, ::up_x] = x
    up = F.pad(up, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    up = up[:, :, max(-py0, 0): up.shape[2] - max(-py1, 0), max(-px0, 0): up.shape[3] - max(-px1, 0)]
    full_h = h * up_y + py0 + py1 - kh + 1
    full_w = w * up_x + px0 + px1 - kw + 1
    out_h = (h * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (w * up_x + px0 + px1 - kw) // down_x + 1
    acc = x.new_zeros(n, c, max(full_h, 0), max(full_w, 0))
    kf = torch.flip(kernel, [0, 1])
    for i in range(kh):
        for j in range(kw):
            acc = acc + kf[i, j] * up[:, :, i:i + full_h, j:j + full_w]
    out = acc[:, :, ::down_y, ::down_x]
    assert out.shape[2] == out_h and out.shape[3] == out_w
    return out.contiguous()


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """utils/op/upfirdn2d.py:170-181 (scalar up/down, 2-pad applied to both axes)."""
    return upfirdn2d_ref(x, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def fused_leaky_relu_ref(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """utils/op/fused_act.py:112-120 CPU branch (NOTE it hard-codes slope 0.2, :117;
    the CUDA branch honours the argument, fused_bias_act_kernel.cu:38-47 - we restate
    the CUDA semantics: lrelu(x + b, slope) * scale)."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    return F.leaky_relu(x + bias.view(*shape), negative_slope=negative_slope) * scale


def upsample_2d(x, k=(1, 3, 3, 1), factor=2, gain=1):
    """up_or_down_sampling.py:200-229."""
    kk = setup_kernel(k) * (gain * factor ** 2)
    p = kk.shape[0] - factor
    return upfirdn2d(x, torch.tensor(kk, dtype=x.dtype), up=factor,
                     pad=((p + 1) // 2 + factor - 1, p // 2))


def downsample_2d(x, k=(1, 3, 3, 1), factor=2, gain=1):
    """up_or_down_sampling.py:232-262."""
    kk = setup_kernel(k) * gain
    p = kk.shape[0] - factor
    return upfirdn2d(x, torch.tensor(kk, dtype=x.dtype), down=factor, pad=((p + 1) // 2, p // 2))


def conv_downsample_2d(x, w, k=(1, 3, 3, 1), factor=2, gain=1):
    """up_or_down_sampling.py:149-183: FIR pre-filter then stride-`factor` conv, no pad."""
    kk = setup_kernel(k) * gain
    p = (kk.shape[0] - factor) + (w.shape[-1] - 1)
    x = upfirdn2d(x, torch.tensor(kk, dtype=x.dtype), pad=((p + 1) // 2, p // 2))
    return F.conv2d(x, w, stride=factor, padding=0)


# ----------------------------------------------------------------------------------
# diffusion schedule / posterior (engine/test.py:48-177)
# ----------------------------------------------------------------------------------
def sigma_schedule(cfg):
    """engine/test.py:75-97 (float64 time grid, fp32 betas with a leading 1e-8)."""
    n = cfg.num_timesteps
    t = torch.from_numpy(np.arange(0, n + 1, dtype=np.float64) / n) * (1.0 - 1e-3) + 1e-3
    if cfg.use_geometric:
        var = cfg.beta_min * ((cfg.beta_max / cfg.beta_min) ** t)
    else:
        var = 1.0 - torch.exp(2.0 * (-0.25 * t ** 2 * (cfg.beta_max - cfg.beta_min) - 0.5 * t * cfg.beta_min))
    abar = 1.0 - var
    betas = torch.cat((torch.tensor(1e-8)[None], 1 - abar[1:] / abar[:-1])).type(torch.float32)
    return betas ** 0.5, torch.sqrt(1 - betas), betas


class PosteriorCoefficients:
    """engine/test.py:101-123."""

    def __init__(self, cfg):
        betas = sigma_schedule(cfg)[2].type(torch.float32)[1:]
        self.betas = betas
        alphas = 1 - betas
        acp = torch.cumprod(alphas, 0)
        acp_prev = torch.cat((torch.tensor([1.0], dtype=torch.float32), acp[:-1]), 0)
        self.posterior_variance = betas * (1 - acp_prev) / (1 - acp)
        self.posterior_mean_coef1 = betas * torch.sqrt(acp_prev) / (1 - acp)
        self.posterior_mean_coef2 = (1 - acp_prev) * torch.sqrt(alphas) / (1 - acp)
        self.posterior_log_variance_clipped = torch.log(self.posterior_variance.clamp(min=1e-20))


def _extract(tab, t, ndim):
    return torch.gather(tab, 0, t).reshape(-1, *([1] * (ndim - 1)))      # engine/test.py:58-63


def sample_posterior_combine(co: PosteriorCoefficients, x01, x02, xt, t, noise):
    """engine/test.py:150-177 with the noise passed in (the reference draws
    `torch.randn_like(x_t)` at :169 - also at t == 0, where it is masked out)."""
    c1 = _extract(co.posterior_mean_coef1, t, xt.ndim)
    c2 = _extract(co.posterior_mean_coef2, t, xt.ndim)
    mean = ((c1 * x01 + c2 * xt) + (c1 * x02 + c2 * xt)) / 2
    lv = _extract(co.posterior_log_variance_clipped, t, xt.ndim)
    mask = 1 - (t == 0).type(torch.float32)
    return mean + mask[:, None, None, None] * torch.exp(0.5 * lv) * noise


# ----------------------------------------------------------------------------------
# generator forward (functional, on a state_dict)
# ----------------------------------------------------------------------------------
def timestep_embedding(t, dim, max_positions=10000):
    """layers.py:465-479."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
    e = t.float()[:, None] * freq[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if dim % 2 == 1:
        e = F.pad(e, (0, 1))
    return e


class _G:
    """Functional evaluator of one generator over a state_dict."""

    def __init__(self, sd, cfg, variant, hook=None):
        self.sd, self.cfg, self.variant = sd, cfg, variant
        self.spec = build_spec(cfg, variant)
        self.hook = hook

    # --- leaf ops -----------------------------------------------------------------
    def conv(self, key, x, pad=1):
        return F.conv2d(x, self.sd[key + '.weight'], self.sd[key + '.bias'], padding=pad)

    def lin(self, key, x):
        return F.linear(x, self.sd[key + '.weight'], self.sd[key + '.bias'])

    def nin(self, key, x):                                  # layers.py:502-505
        y = torch.einsum('bhwc,cd->bhwd', x.permute(0, 2, 3, 1), self.sd[key + '.W']) + self.sd[key + '.b']
        return y.permute(0, 3, 1, 2)

    def adagn(self, key, x, style):                         # layerspp.py:47-54
        c = x.shape[1]
        s = self.lin(key + '.style', style)[:, :, None, None]
        gamma, beta = s[:, :c], s[:, c:]
        return gamma * F.group_norm(x, _groups(c), eps=1e-6) + beta

    # --- blocks -------------------------------------------------------------------
    def feat(self, p, x, style=None, kind='feat'):          # layerspp.py:410-423,442-455,478-501
        h = self.conv(p + '.conv1', x)
        if kind == 'adafeat':
            h = self.adagn(p + '.group_norm', h, style)
        else:
            h = F.group_norm(h, _groups(h.shape[1]), eps=1e-6)
        h = self.conv(p + '.conv2', F.silu(h))
        if kind == 'gap':
            h = self.lin(p + '.fc', h.mean(dim=(2, 3)))
        return h

    def res(self, p, m, x, temb, zemb):                     # layerspp.py:292-324
        h = F.silu(self.adagn(p + '.GroupNorm_0', x, zemb))
        if m['up']:
            h, x = upsample_2d(h), upsample_2d(x)
        elif m['down']:
            h, x = downsample_2d(h), downsample_2d(x)
        h = self.conv(p + '.Conv_0', h)
        h = h + self.lin(p + '.Dense_0', F.silu(temb))[:, :, None, None]
        h = F.silu(self.adagn(p + '.GroupNorm_1', h, zemb))
        h = self.conv(p + '.Conv_1', h)
        if m['cin'] != m['cout'] or m['up'] or m['down']:
            x = self.conv(p + '.Conv_2', x, pad=0)
        return (x + h) / SQRT2

    def attn(self, p, x):                                   # layerspp.py:111-137
        b, c, hh, ww = x.shape
        h = F.group_norm(x, _groups(c), self.sd[p + '.GroupNorm_0.weight'], self.sd[p + '.GroupNorm_0.bias'], eps=1e-6)
        q, k, v = (self.nin(p + f'.NIN_{j}', h) for j in range(3))
        w = torch.einsum('bchw,bcij->bhwij', q, k) * (int(c) ** (-0.5))
        w = F.softmax(w.reshape(b, hh, ww, hh * ww), dim=-1).reshape(b, hh, ww, hh, ww)
        h = self.nin(p + '.NIN_3', torch.einsum('bhwij,bcij->bchw', w, v))
        return (x + h) / SQRT2                              # skip_rescale=True (generator :112-114)

    def pyrdown(self, p, x):                                # up_or_down_sampling.py:49-61
        y = conv_downsample_2d(x, self.sd[p + '.Conv2d_0.weight'])
        return y + self.sd[p + '.Conv2d_0.bias'].reshape(1, -1, 1, 1)

    # --- whole network --------------------------------------------------------------
    def forward(self, x, conds, t, z, pseudo_target=None):
        cfg, sd = self.cfg, self.sd
        adaptive = self.variant.startswith('g2')
        ncond = len(conds)
        # z mapping network: PixelNorm + (n_mlp+1) x (dense, SiLU)   (:44-49, :271-277)
        ze = z / torch.sqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
        ze = F.silu(self.lin('z_transform.1', ze))
        for j in range(cfg.n_mlp):
            ze = F.silu(self.lin(f'z_transform.{3 + 2 * j}', ze))
        temb = timestep_embedding(t, cfg.num_channels_dae)              # :296
        temb = self.lin('all_modules.0', temb)
        temb = self.lin('all_modules.1', F.silu(temb))                  # :302-305
        if not cfg.centered:
            x = 2 * x - 1.0
        pyr = x
        i = 2
        if not adaptive:                                                # :318-330
            feats = [self.feat(f'all_modules.{i}', x)]
            for j, c in enumerate(conds):
                feats.append(self.feat(f'all_modules.{i + 1 + j}', c))
            i += 1 + ncond
            h0 = torch.cat(feats, dim=1)
        else:                                                           # :733-791
            pw = self.feat(f'all_modules.{i}', pseudo_target, kind='gap')
            xf = self.feat(f'all_modules.{i + 1}', x)
            cf = [self.feat(f'all_modules.{i + 2 + j}', c, style=pw, kind='adafeat') for j, c in enumerate(conds)]
            i += 2 + ncond
            allc = torch.cat(cf, dim=1)
            pairs = [('c12', 'c1', 0, 1), ('c23', 'c2', 1, 2), ('c31', 'c3', 2, 0)] if ncond == 3 else [('c12', 'c1', 0, 1)]
            fused = []
            for an, wn, a, b in pairs:
                g1 = torch.sigmoid(self.conv(f'feat_att1_{an}', allc))
                g2 = torch.sigmoid(self.conv(f'feat_att2_{an}', allc))
                att = self.conv(f'feat_weight_{wn}', g1 * cf[a])
                fused.append(g2 * att + (1 - g2) * cf[b])
            h0 = torch.cat([xf] + fused, dim=1)
        if self.hook:
            self.hook('stem', h0)
        hs = [h0]
        nlev = len(cfg.ch_mult)
        spec = self.spec

        def run_res(inp):
            nonlocal i
            out = self.res(f'all_modules.{i}', spec[i], inp, temb, ze)
            if self.hook:
                self.hook(f'res{i}', out)
            i += 1
            return out

        for lv in range(nlev):                                          # :335-368
            for _ in range(cfg.num_res_blocks):
                h = run_res(hs[-1])
                if h.shape[-1] in cfg.attn_resolutions:
                    h = self.attn(f'all_modules.{i}', h)
                    i += 1
                hs.append(h)
            if lv != nlev - 1:
                h = run_res(hs[-1])
                pyr = self.pyrdown(f'all_modules.{i}', pyr)
                i += 1
                pyr = (pyr + h) / SQRT2
                h = pyr
                hs.append(h)
        h = run_res(hs[-1])                                             # :370-376
        h = self.attn(f'all_modules.{i}', h)
        if self.hook:
            self.hook(f'attn{i}', h)
        i += 1
        h = run_res(h)
        for lv in reversed(range(nlev)):                                # :381-429
            for _ in range(cfg.num_res_blocks + 1):
                h = run_res(torch.cat([h, hs.pop()], dim=1))
            if h.shape[-1] in cfg.attn_resolutions:
                h = self.attn(f'all_modules.{i}', h)
                i += 1
            if lv != 0:
                h = run_res(h)
        assert not hs
        p = f'all_modules.{i}'                                          # :436-445
        h = F.silu(F.group_norm(h, _groups(h.shape[1]), sd[p + '.weight'], sd[p + '.bias'], eps=1e-6))
        h = self.conv(f'all_modules.{i + 1}', h)
        assert i + 2 == len(spec)
        return h if cfg.not_use_tanh else torch.tanh(h)


def generator_forward(sd, cfg, variant, x, conds, t, z, pseudo_target=None, hook=None):
    """G1: NCSNpp.forward(x, cond1, cond2[, cond3], time_cond, z)  (:279 / healthy :279)
    G2: NCSNpp_adaptive.forward(..., z, pseudo_target)             (:694 / healthy :693)"""
    with torch.no_grad():
        return _G(sd, cfg, variant, hook).forward(x, list(conds), t, z, pseudo_target)


def sample_from_model(co, sd1, sd2, cfg, conds, x_init, latents, noises, healthy=False, n_time=None):
    """engine/test.py:180-199 with pre-drawn `latents[i]` (=`latent_z`, :188) and
    `noises[i]` (=`randn_like`, :169), both indexed by the step index i in 3,2,1,0.
    No autocast: this is the fp32 oracle."""
    n_time = cfg.num_timesteps if n_time is None else n_time
    v1, v2 = ('g1_healthy', 'g2_healthy') if healthy else ('g1', 'g2')
    x = x_init
    for i in reversed(range(n_time)):
        t = torch.full((x.size(0),), i, dtype=torch.int64)
        x01 = generator_forward(sd1, cfg, v1, x, conds, t, latents[i])
        x02 = generator_forward(sd2, cfg, v2, x, conds, t, latents[i], pseudo_target=x01[:, [0], :])
        x = sample_posterior_combine(co, x01[:, [0], :], x02[:, [0], :], x, t, noises[i])
    return x


# ----------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------
def synthetic_inputs(batch, size, cfg, ncond=3, seed=42):
    g = torch.Generator().manual_seed(seed)
    conds = [torch.randn(batch, 1, size, size, generator=g).clamp(-3, 3) / 3 for _ in range(ncond)]
    x_init = torch.randn(batch, 1, size, size, generator=g)
    latents = [torch.randn(batch, cfg.nz, generator=g) for _ in range(cfg.num_timesteps)]
    noises = [torch.randn(batch, 1, size, size, generator=g) for _ in range(cfg.num_timesteps)]
    return conds, x_init, latents, noises


def psnr(a, b, data_range=1.0):
    """tools/metric_calc.py:40 convention (skimage PSNR, data_range=1.0) on [0,1] images."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float('inf') if mse == 0 else 10.0 * math.log10(data_range ** 2 / mse)

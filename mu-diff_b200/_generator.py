"""Shared implementation of the four MU-Diff generators (main / healthy x G1 / G2).

Module construction order, attribute names and therefore state_dict keys follow
backbones/ncsnpp_generator_adagn_feat.py:56-277 (NCSNpp) and :454-692 (NCSNpp_adaptive)
(+ the healthy file's 2-contrast stem, ncsnpp_generator_adagn_feat_healthy.py:177-184,
577-631).  The forward walks `all_modules` by running index exactly like :279-447 /
:694-905, but
  * the input is converted once to the compute dtype (bf16 tensor-core path or fp32
    CUDA-core path), channels-last, and never leaves it until the final tanh;
  * the four stem feature maps are written by their producing convs straight into the
    concatenated [B, 4nf, H, W] buffer (torch.cat of :330/:791 disappears);
  * skip concats `torch.cat([h, hs.pop()], 1)` (:383) are passed as (h, skip) tuples;
  * all AdaGN style Linear layers and all Dense_0(act(temb)) of a forward are evaluated
    by TWO batched GEMM launches (they only depend on z and t);
  * the one-shot debug prints of :735-753 are not reproduced (they break graph capture).
"""
import functools

import torch
from torch import nn

from . import _lib as L
from . import dense_layer, layers, layerspp, ops

ResnetBlockBigGAN = layerspp.ResnetBlockBigGANpp_Adagn
conv3x3 = layerspp.conv3x3
default_initializer = layers.default_init
dense = dense_layer.dense


class PixelNorm(nn.Module):
    """ncsnpp_generator_adagn_feat.py:44-49"""

    def forward(self, input):
        return ops.pixelnorm(input)


def _get(config, name, default):
    return getattr(config, name, default)


class _NCSNppBase(nn.Module, layers.PackCache):
    adaptive = False      # G2 (pseudo-target GAP style + cross-contrast gates)
    n_cond = 3            # conditioning contrasts (2 for the 'healthy' variant)

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.not_use_tanh = config.not_use_tanh
        self.act = act = nn.SiLU()
        self.z_emb_dim = z_emb_dim = config.z_emb_dim
        self.nf = nf = config.num_channels_dae
        ch_mult = config.ch_mult
        self.num_res_blocks = num_res_blocks = config.num_res_blocks
        self.attn_resolutions = attn_resolutions = config.attn_resolutions
        dropout = config.dropout
        resamp_with_conv = config.resamp_with_conv
        self.num_resolutions = num_resolutions = len(ch_mult)
        self.all_resolutions = all_resolutions = [config.image_size // (2 ** i) for i in range(num_resolutions)]
        self.conditional = conditional = config.conditional
        fir, fir_kernel = config.fir, config.fir_kernel
        self.skip_rescale = skip_rescale = config.skip_rescale
        self.resblock_type = resblock_type = config.resblock_type.lower()
        self.progressive = progressive = config.progressive.lower()
        self.progressive_input = progressive_input = config.progressive_input.lower()
        self.embedding_type = embedding_type = config.embedding_type.lower()
        init_scale = 0.
        assert progressive in ['none', 'output_skip', 'residual']
        assert progressive_input in ['none', 'input_skip', 'residual']
        assert embedding_type in ['fourier', 'positional']
        # B200 path: the configuration every MU-Diff experiment uses (README.md:85, experiments/cfg/local.yaml)
        if (embedding_type != 'positional' or not conditional or resblock_type != 'biggan' or progressive != 'none'
                or progressive_input != 'residual' or not fir or not resamp_with_conv):
            raise NotImplementedError(
                "mu-diff_b200 implements the MU-Diff sampling configuration: embedding_type='positional', "
                "conditional, resblock_type='biggan', progressive='none', progressive_input='residual', fir, "
                "resamp_with_conv")
        self.precision = _get(config, 'b200_precision', 'bf16')     # 'bf16' | 'fp32'

        modules = []
        embed_dim = nf
        modules.append(nn.Linear(embed_dim, nf * 4))
        modules[-1].weight.data = default_initializer()(modules[-1].weight.shape)
        nn.init.zeros_(modules[-1].bias)
        modules.append(nn.Linear(nf * 4, nf * 4))
        modules[-1].weight.data = default_initializer()(modules[-1].weight.shape)
        nn.init.zeros_(modules[-1].bias)

        AttnBlock = functools.partial(layerspp.AttnBlockpp, init_scale=init_scale, skip_rescale=skip_rescale)
        pyramid_downsample = functools.partial(layerspp.Downsample, fir=fir, fir_kernel=fir_kernel, with_conv=True)
        ResnetBlock = functools.partial(ResnetBlockBigGAN, act=act, dropout=dropout, fir=fir, fir_kernel=fir_kernel,
                                        init_scale=init_scale, skip_rescale=skip_rescale, temb_dim=nf * 4,
                                        zemb_dim=z_emb_dim)
        channels = config.num_channels
        input_pyramid_ch = channels
        if not self.adaptive:
            for _ in range(1 + self.n_cond):
                modules.append(layerspp.ConvFeatBlock(act=act, in_ch=channels, out_ch=nf))
            stem_c = nf * (1 + self.n_cond)
        else:
            modules.append(layerspp.ConvBlock_GAP(act=act, in_ch=channels, out_ch=nf))
            modules.append(layerspp.ConvFeatBlock(act=act, in_ch=channels, out_ch=nf))
            for _ in range(self.n_cond):
                modules.append(layerspp.ConvBlock(act=act, in_ch=channels, out_ch=nf))
            stem_c = nf * 4 if self.n_cond == 3 else nf * 2
        self.stem_c = stem_c
        hs_c = [stem_c]
        in_ch = stem_c
        for i_level in range(num_resolutions):
            for _ in range(num_res_blocks):
                out_ch = nf * ch_mult[i_level]
                modules.append(ResnetBlock(in_ch=in_ch, out_ch=out_ch))
                in_ch = out_ch
                if all_resolutions[i_level] in attn_resolutions:
                    modules.append(AttnBlock(channels=in_ch))
                hs_c.append(in_ch)
            if i_level != num_resolutions - 1:
                modules.append(ResnetBlock(down=True, in_ch=in_ch))
                modules.append(pyramid_downsample(in_ch=input_pyramid_ch, out_ch=in_ch))
                input_pyramid_ch = in_ch
                hs_c.append(in_ch)
        in_ch = hs_c[-1]
        modules.append(ResnetBlock(in_ch=in_ch))
        modules.append(AttnBlock(channels=in_ch))
        modules.append(ResnetBlock(in_ch=in_ch))

        if self.adaptive:
            if self.n_cond == 3:
                self.feat_weight_c1 = conv3x3(nf, nf)
                self.feat_weight_c2 = conv3x3(nf, nf)
                self.feat_weight_c3 = conv3x3(nf, nf)
                self.feat_att1_c12 = conv3x3(3 * nf, nf)
                self.feat_att2_c12 = conv3x3(3 * nf, nf)
                self.feat_att1_c23 = conv3x3(3 * nf, nf)
                self.feat_att2_c23 = conv3x3(3 * nf, nf)
                self.feat_att1_c31 = conv3x3(3 * nf, nf)
                self.feat_att2_c31 = conv3x3(3 * nf, nf)
            else:
                self.feat_weight_c1 = conv3x3(nf, nf)
                self.feat_att1_c12 = conv3x3(2 * nf, nf)
                self.feat_att2_c12 = conv3x3(2 * nf, nf)

        for i_level in reversed(range(num_resolutions)):
            for _ in range(num_res_blocks + 1):
                out_ch = nf * ch_mult[i_level]
                modules.append(ResnetBlock(in_ch=in_ch + hs_c.pop(), out_ch=out_ch))
                in_ch = out_ch
            if all_resolutions[i_level] in attn_resolutions:
                modules.append(AttnBlock(channels=in_ch))
            if i_level != 0:
                modules.append(ResnetBlock(in_ch=in_ch, up=True))
        assert not hs_c
        modules.append(layerspp.GroupNorm(num_groups=min(in_ch // 4, 32), num_channels=in_ch, eps=1e-6))
        modules.append(conv3x3(in_ch, channels, init_scale=init_scale))
        self.all_modules = nn.ModuleList(modules)

        mapping_layers = [PixelNorm(), dense(config.nz, z_emb_dim), self.act]
        for _ in range(config.n_mlp):
            mapping_layers.append(dense(z_emb_dim, z_emb_dim))
            mapping_layers.append(self.act)
        self.z_transform = nn.Sequential(*mapping_layers)

    # ------------------------------------------------------------------------------
    def compute_dtype(self):
        p = self.precision
        if p in ('bf16', torch.bfloat16):
            return torch.bfloat16
        if p in ('fp32', 'f32', torch.float32):
            return torch.float32
        raise ValueError(f"b200_precision must be 'bf16' or 'fp32', got {p!r}")

    def _embeddings(self, time_cond, z):
        """zemb (:282), temb (:296-305)."""
        ze = ops.pixelnorm(z)
        for m in self.z_transform:
            if isinstance(m, nn.Linear):
                ze = ops.linear(ze, m.weight, m.bias, act_out=L.ACT_SILU)
        modules = self.all_modules
        temb = layers.get_timestep_embedding(time_cond, self.nf)
        temb = ops.linear(temb, modules[0].weight, modules[0].bias)
        temb = ops.linear(temb, modules[1].weight, modules[1].bias, act_in=L.ACT_SILU)
        return ze, temb

    def _batched_styles(self, zemb, temb):
        """One GEMM for every AdaGN style layer fed by zemb, one for every Dense_0(act(temb)).
        Returns {id(block): (gb0, gb1, tbias)} of views into the two result matrices."""
        blocks = [m for m in self.all_modules if isinstance(m, ResnetBlockBigGAN)]
        params = []
        for b in blocks:
            params += [b.GroupNorm_0.style.weight, b.GroupNorm_0.style.bias, b.GroupNorm_1.style.weight,
                       b.GroupNorm_1.style.bias, b.Dense_0.weight, b.Dense_0.bias]

        def build():
            ws = torch.cat([torch.cat([b.GroupNorm_0.style.weight, b.GroupNorm_1.style.weight], 0) for b in blocks], 0)
            bs = torch.cat([torch.cat([b.GroupNorm_0.style.bias, b.GroupNorm_1.style.bias], 0) for b in blocks], 0)
            wt = torch.cat([b.Dense_0.weight for b in blocks], 0)
            bt = torch.cat([b.Dense_0.bias for b in blocks], 0)
            return ws.float().contiguous(), bs.float().contiguous(), wt.float().contiguous(), bt.float().contiguous()

        ws, bs, wt, bt = self._packed(('styles',), params, build)
        sty = ops.linear(zemb, ws, bs)
        tb = ops.linear(temb, wt, bt, act_in=L.ACT_SILU)
        out, so, to = {}, 0, 0
        for b in blocks:
            c0, c1 = 2 * b.in_ch, 2 * b.out_ch
            out[id(b)] = (sty[:, so:so + c0], sty[:, so + c0:so + c0 + c1], tb[:, to:to + b.out_ch])
            so += c0 + c1
            to += b.out_ch
        return out

    def _stem(self, x, conds, pseudo_target, dt):
        modules = self.all_modules
        nf = self.nf
        b, _, h, w = x.shape
        m_idx = 2
        if not self.adaptive:
            # The ConvFeatBlock stems of G1 (conv3x3 -> GroupNorm -> SiLU -> conv3x3) depend on neither t nor z: inside one pass of
            # the sampling loop (ops.stem_moments_scope) the features of the CONDITIONING contrasts are the same in all diffusion
            # steps.  h0 (the stem concat) and its statistics are kept for the sample and only x_t's slice is recomputed:
            # 3 x (stem + 64 -> 64 conv + statistics pass) per step instead of 12 per sample for the three later steps.  Nothing
            # writes h0 after the stem (it is read as a ResBlock input and as a skip connection).
            scope = ops.loop_scope()
            key = ('g1_stem', id(self), dt) + tuple((c.data_ptr(), tuple(c.shape), tuple(c.stride()), c._version) for c in conds)
            ent = scope.get(key) if scope is not None else None
            if ent is not None and ent[0].shape[0] == b:
                h0, cs0 = ent[0], ent[1]
                modules[m_idx](x, compute_dtype=dt, out=h0, out_coff=0, stats_out=(cs0, 0))
                return h0, m_idx + 1 + len(conds)
            h0 = ops.empty_nhwc(b, self.stem_c, h, w, dt, x.device)
            cs0 = torch.empty((b, self.stem_c, 2), dtype=torch.float64, device=x.device)   # per-channel GN statistics of h0
            ops.set_chstats(h0, cs0)
            for j, inp in enumerate([x] + list(conds)):
                modules[m_idx](inp, compute_dtype=dt, out=h0, out_coff=j * nf, stats_out=(cs0, j * nf))
                m_idx += 1
            if scope is not None:
                scope[key] = (h0, cs0, list(conds))        # the conds stay alive: their addresses cannot be reused in the scope
            return h0, m_idx
        h0 = ops.empty_nhwc(b, self.stem_c, h, w, dt, x.device)
        cs0 = torch.empty((b, self.stem_c, 2), dtype=torch.float64, device=x.device)   # per-channel GN statistics of h0
        ops.set_chstats(h0, cs0)
        # ---- adaptive (G2): :733-791 -------------------------------------------------
        pseudo_weight = modules[m_idx](pseudo_target, compute_dtype=dt)
        m_idx += 1
        modules[m_idx](x, compute_dtype=dt, out=h0, out_coff=0, stats_out=(cs0, 0))
        m_idx += 1
        nc = self.n_cond
        cf = ops.empty_nhwc(b, nc * nf, h, w, dt, x.device)          # cat(cond1_feat, cond2_feat[, cond3_feat])
        for j, c in enumerate(conds):
            modules[m_idx](c, pseudo_weight, compute_dtype=dt, out=cf, out_coff=j * nf)
            m_idx += 1
        feats = [cf[:, j * nf:(j + 1) * nf] for j in range(nc)]
        if nc == 3:
            gates = [self.feat_att1_c12, self.feat_att2_c12, self.feat_att1_c23, self.feat_att2_c23,
                     self.feat_att1_c31, self.feat_att2_c31]
            pairs = [(0, 1, self.feat_weight_c1), (1, 2, self.feat_weight_c2), (2, 0, self.feat_weight_c3)]
        else:
            gates = [self.feat_att1_c12, self.feat_att2_c12]
            pairs = [(0, 1, self.feat_weight_c1)]
        # all sigmoid gates share their input: ONE conv with N = len(gates)*nf and a sigmoid epilogue
        wg = self._packed(('gates', dt), [g.weight for g in gates],
                          lambda: torch.cat([g.packed_weight(dt) for g in gates], 0).contiguous())
        bg = self._packed(('gates_b',), [g.bias for g in gates],
                          lambda: torch.cat([g.bias for g in gates]).detach().float().contiguous())
        g_all = ops.conv([(cf, 9)], wg, len(gates) * nf, bias=bg, act=L.ACT_SIGMOID)
        for p, (a, bb, wconv) in enumerate(pairs):
            g1 = g_all[:, (2 * p) * nf:(2 * p + 1) * nf]
            g2 = g_all[:, (2 * p + 1) * nf:(2 * p + 2) * nf]
            att = wconv(ops.gate_mul(g1, feats[a]), compute_dtype=dt)
            ops.gate_blend(g2, att, feats[bb], out=h0[:, (1 + p) * nf:(2 + p) * nf])
        ops.gn_stats(h0[:, nf:], out=(cs0, nf))          # gated features: stand-alone statistics pass
        return h0, m_idx

    def _forward(self, x, conds, time_cond, z, pseudo_target=None):
        L.require_cuda(x, z, time_cond)
        if self.training:
            self.eval()          # inference-only path (dropout=0 at sampling time, engine/test.py:279-283)
        dt = self.compute_dtype()
        modules = self.all_modules
        zemb, temb = self._embeddings(time_cond, z)
        pre = self._batched_styles(zemb, temb)
        if not self.config.centered:
            x = 2 * x - 1.
        x = x.float()
        conds = [c.float() for c in conds]
        input_pyramid = x
        h0, m_idx = self._stem(x, conds, pseudo_target.float() if pseudo_target is not None else None, dt)
        sc = ops.SQRT2_INV if self.skip_rescale else 1.0

        def res(inp):
            nonlocal m_idx
            blk = modules[m_idx]
            g0, g1, tb = pre[id(blk)]
            out = blk(inp, temb, zemb, gb0=g0, gb1=g1, tbias=tb)
            m_idx += 1
            return out

        hs = [h0]
        for i_level in range(self.num_resolutions):
            for _ in range(self.num_res_blocks):
                h = res(hs[-1])
                if h.shape[-1] in self.attn_resolutions:
                    h = modules[m_idx](h)
                    m_idx += 1
                hs.append(h)
            if i_level != self.num_resolutions - 1:
                h = res(hs[-1])
                input_pyramid = modules[m_idx](input_pyramid, compute_dtype=dt)
                m_idx += 1
                input_pyramid = ops.add_scale(input_pyramid, h, sc)
                h = input_pyramid
                hs.append(h)
        h = res(hs[-1])
        h = modules[m_idx](h)
        m_idx += 1
        h = res(h)
        for i_level in reversed(range(self.num_resolutions)):
            for _ in range(self.num_res_blocks + 1):
                h = res((h, hs.pop()))
            if h.shape[-1] in self.attn_resolutions:
                h = modules[m_idx](h)
                m_idx += 1
            if i_level != 0:
                h = res(h)
        assert not hs
        h = modules[m_idx](h, act=L.ACT_SILU)
        m_idx += 1
        head = modules[m_idx]
        m_idx += 1
        assert m_idx == len(modules)
        wt = head.packed_weight(dt)
        out = ops.conv([(h, 9)], wt, head.out_channels, bias=head.bias_f32(), out_dtype=torch.float32,
                       act=L.ACT_NONE if self.not_use_tanh else L.ACT_TANH, force='simt')
        return out

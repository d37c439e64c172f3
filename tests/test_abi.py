"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads without a GPU and
exports every symbol include/mudiff_b200.h declares; the Python host mirrors the reference's
module/state_dict interface; the product never imports the oracle."""
import ctypes
import os
import re
from argparse import Namespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def M():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, 'mu-diff_b200', 'libmudiff_b200.so')):
        g.build()
    import mudiff_b200
    return mudiff_b200


def test_library_exports_every_declared_symbol(M):
    hdr = open(os.path.join(ROOT, 'include', 'mudiff_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(mudiff_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(M._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert set(M._lib.PROTOTYPES) == declared
    assert lib.mudiff_abi_version() == 1
    assert lib.mudiff_conv_desc_size() == ctypes.sizeof(M._lib.ConvDesc)


def test_no_cpu_fallback(M):
    x = torch.zeros(1, 2, 8, 8)
    k = torch.ones(2, 2)
    with pytest.raises(RuntimeError):
        M.upfirdn2d(x, k)
    with pytest.raises(RuntimeError):
        M.fused_leaky_relu(x, torch.zeros(2))


def test_operator_signatures_match_reference(M):
    import inspect
    assert list(inspect.signature(M.upfirdn2d).parameters) == ['input', 'kernel', 'up', 'down', 'pad']
    assert list(inspect.signature(M.fused_leaky_relu).parameters) == ['input', 'bias', 'negative_slope', 'scale']
    assert list(inspect.signature(M.FusedLeakyReLU.__init__).parameters) == ['self', 'channel', 'negative_slope', 'scale']
    lp = M.layerspp
    assert list(inspect.signature(lp.ResnetBlockBigGANpp_Adagn.__init__).parameters) == [
        'self', 'act', 'in_ch', 'out_ch', 'temb_dim', 'zemb_dim', 'up', 'down', 'dropout', 'fir', 'fir_kernel',
        'skip_rescale', 'init_scale']
    assert list(inspect.signature(lp.ResnetBlockBigGANpp_Adagn.forward).parameters)[:4] == ['self', 'x', 'temb', 'zemb']
    assert list(inspect.signature(lp.AttnBlockpp.__init__).parameters) == ['self', 'channels', 'skip_rescale', 'init_scale']
    assert list(inspect.signature(lp.Combine.__init__).parameters) == ['self', 'dim1', 'dim2', 'method']
    assert list(inspect.signature(lp.AdaptiveGroupNorm.__init__).parameters) == ['self', 'num_groups', 'in_channel', 'style_dim']
    assert list(inspect.signature(M.NCSNpp.forward).parameters) == ['self', 'x', 'cond1', 'cond2', 'cond3', 'time_cond', 'z']
    assert list(inspect.signature(M.NCSNpp_adaptive.forward).parameters) == [
        'self', 'x', 'cond1', 'cond2', 'cond3', 'time_cond', 'z', 'pseudo_target']
    hh = M.ncsnpp_generator_adagn_feat_healthy
    assert list(inspect.signature(hh.NCSNpp.forward).parameters) == ['self', 'x', 'cond1', 'cond2', 'time_cond', 'z']
    assert list(inspect.signature(M.sample_from_model).parameters)[:10] == [
        'coefficients', 'generator1', 'cond1', 'generator2', 'cond2', 'cond3', 'n_time', 'x_init', 'T', 'opt']


def test_state_dict_keys_and_param_counts(M):
    """Keys/shapes equal the reference's (the oracle state_dict was loaded strict=True into the
    reference modules by tests/golden/make_golden.py); counts equal error_logs/...out:116."""
    from oracle import mudiff_oracle as O
    cfg = O.default_config()
    ns = Namespace(**vars(cfg))
    g1, g2 = M.NCSNpp(ns), M.NCSNpp_adaptive(ns)
    assert sum(p.numel() for p in g1.parameters()) == 20472065
    assert sum(p.numel() for p in g2.parameters()) == 21399681
    g1.load_state_dict(O.make_state_dict(cfg, 'g1'), strict=True)
    g2.load_state_dict(O.make_state_dict(cfg, 'g2'), strict=True)
    hh = M.ncsnpp_generator_adagn_feat_healthy
    h1, h2 = hh.NCSNpp(ns), hh.NCSNpp_adaptive(ns)
    h1.load_state_dict(O.make_state_dict(cfg, 'g1_healthy'), strict=True)
    h2.load_state_dict(O.make_state_dict(cfg, 'g2_healthy'), strict=True)
    assert len(list(g1.buffers())) == 0


def test_posterior_tables_bit_identical(M, golden_dir):
    import numpy as np
    from oracle import mudiff_oracle as O
    g = np.load(os.path.join(golden_dir, 'posterior.npz'))
    co = M.Posterior_Coefficients(Namespace(**vars(O.default_config())), 'cpu')
    np.testing.assert_array_equal(co.posterior_mean_coef1.numpy(), g['coef1'])
    np.testing.assert_array_equal(co.posterior_mean_coef2.numpy(), g['coef2'])
    np.testing.assert_array_equal(co.posterior_log_variance_clipped.numpy(), g['log_var'])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'mu-diff_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f


def test_fir_setup_kernel(M):
    import numpy as np
    from oracle import mudiff_oracle as O
    np.testing.assert_array_equal(M.up_or_down_sampling._setup_kernel([1, 3, 3, 1]), O.setup_kernel([1, 3, 3, 1]))


def test_validation_share_weights_aliases_training_parameters():
    """the fast generators alias the training modules' parameters (no copy), also through a DDP-like
    `.module` wrapper; an in-place optimiser-style update is visible and bumps the version the pack cache keys on."""
    import torch
    from torch import nn
    import mudiff_b200 as M
    from mudiff_b200 import validation as VAL
    from oracle import mudiff_oracle as O
    from argparse import Namespace
    cfg = O.default_config(num_channels_dae=16, image_size=32)
    ns = Namespace(**vars(cfg), b200_precision='bf16')
    train_g = M.NCSNpp(ns)                        # stands in for the reference module: same state_dict keys (tested above)
    fast_g = M.NCSNpp(ns)

    class Wrapper(nn.Module):                     # what DistributedDataParallel looks like from outside
        def __init__(self, m):
            super().__init__()
            self.module = m

    VAL.share_weights(fast_g, Wrapper(train_g))
    tp, fp = dict(train_g.named_parameters()), dict(fast_g.named_parameters())
    assert tp.keys() == fp.keys()
    assert all(tp[k].data_ptr() == fp[k].data_ptr() for k in tp)
    assert not any(p.requires_grad for p in fast_g.parameters()) and not fast_g.training
    sig0 = VAL.params_signature((fast_g,))
    with torch.no_grad():
        next(iter(tp.values())).add_(1.0)         # optimiser step on the training module
    k0 = next(iter(tp))
    assert torch.equal(tp[k0], fp[k0])
    assert VAL.params_signature((fast_g,)) != sig0


def test_pack_cache_refreshes_in_place():
    """layers.PackCache: a packed copy never aliases its parameter, and after an in-place weight update (or a
    load_state_dict) the copy is rebuilt INTO the same tensor, so a captured CUDA graph that embeds its address stays
    valid (validation.ValidationSampler relies on this)."""
    import torch
    import mudiff_b200 as M
    conv = M.layers.ddpm_conv3x3(8, 16)
    w0 = conv.packed_weight(torch.bfloat16)
    b0 = conv.bias_f32()
    assert b0.data_ptr() != conv.bias.data_ptr()                   # fp32 -> fp32 "cast" must still be a copy
    ptr_w, ptr_b, snap = w0.data_ptr(), b0.data_ptr(), w0.clone()
    with torch.no_grad():
        conv.weight.mul_(2.0)
        conv.bias.add_(1.0)
    assert M.layers.refresh_packs(conv) == 2
    w1, b1 = conv.packed_weight(torch.bfloat16), conv.bias_f32()
    assert w1.data_ptr() == ptr_w and b1.data_ptr() == ptr_b
    assert torch.equal(w1.float(), snap.float() * 2.0) and torch.equal(b1, conv.bias.detach())
    conv.load_state_dict({'weight': torch.ones_like(conv.weight), 'bias': torch.zeros_like(conv.bias)})
    assert conv.packed_weight(torch.bfloat16).data_ptr() == ptr_w
    assert float(conv.packed_weight(torch.bfloat16).float().min()) == 1.0


def test_loop_scope_nesting_and_cleanup():
    """ops.stem_moments_scope (the per-sample scope of the sampling loop): nothing is cached outside it, nested scopes share
    one dict, and the scope is closed again when the loop raises."""
    from mudiff_b200 import ops
    assert ops.loop_scope() is None
    with ops.stem_moments_scope():
        d = ops.loop_scope()
        assert d == {}
        d['k'] = 1
        with ops.stem_moments_scope():
            assert ops.loop_scope() is d
        assert ops.loop_scope() is d
    assert ops.loop_scope() is None
    with pytest.raises(ValueError):
        with ops.stem_moments_scope():
            raise ValueError('loop failed')
    assert ops.loop_scope() is None
    with ops.stem_moments_scope():
        assert ops.loop_scope() == {}          # a new sample starts with an empty scope

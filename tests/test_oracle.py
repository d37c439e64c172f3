"""The oracle (oracle/mudiff_oracle.py) against the fixtures produced by the reference
itself (tests/golden/make_golden.py) and the reference's known-answers."""
import os

import numpy as np
import pytest
import torch

from oracle import mudiff_oracle as O


def _npz(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_param_count_known_answer(golden_dir):
    # error_logs/log_mudiff_T1.13967221.out:116 : G1=20,472,065  G2=21,399,681
    cfg = O.default_config()
    assert O.param_count(O.make_state_dict(cfg, 'g1')) == 20472065
    assert O.param_count(O.make_state_dict(cfg, 'g2')) == 21399681
    g = _npz(golden_dir, 'gen_main.npz')
    assert list(g['nparams']) == [20472065, 21399681]
    h = _npz(golden_dir, 'gen_healthy.npz')     # SURVEY.md §8 a6 [probe]
    assert O.param_count(O.make_state_dict(cfg, 'g1_healthy')) == int(h['nparams'][0]) == 20286785
    assert O.param_count(O.make_state_dict(cfg, 'g2_healthy')) == int(h['nparams'][1]) == 20443585


def test_posterior_tables(golden_dir):
    g = _npz(golden_dir, 'posterior.npz')
    co = O.PosteriorCoefficients(O.default_config())
    np.testing.assert_array_equal(co.betas.numpy(), g['betas'])
    np.testing.assert_array_equal(co.posterior_mean_coef1.numpy(), g['coef1'])
    np.testing.assert_array_equal(co.posterior_mean_coef2.numpy(), g['coef2'])
    np.testing.assert_array_equal(co.posterior_log_variance_clipped.numpy(), g['log_var'])
    # SURVEY.md §8 a2 goldens (beta_min=0.1, beta_max=20, n=4)
    np.testing.assert_allclose(g['betas'], [0.478255302, 0.849206030, 0.956417680, 0.987403929], rtol=1e-6)
    np.testing.assert_allclose(g['coef1'], [1.000000119, 0.665778458, 0.269190848, 0.057821561], rtol=1e-6)
    np.testing.assert_allclose(g['coef2'], [0, 0.201576263, 0.193000868, 0.111852221], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(g['log_var'], [-46.0517006, -0.819120586, -0.123069257, -0.0160676315], rtol=1e-6)


def test_posterior_update(golden_dir):
    g = _npz(golden_dir, 'posterior.npz')
    co = O.PosteriorCoefficients(O.default_config())
    T = torch.from_numpy
    y = O.sample_posterior_combine(co, T(g['x01']), T(g['x02']), T(g['xt']), T(g['t']), T(g['noise']))
    np.testing.assert_array_equal(y.numpy(), g['y'])          # same ops, same order: bit-exact


@pytest.mark.parametrize('name', ['down2', 'up2', 'pre', 'negpad', 'up3down2', 'odd'])
def test_upfirdn2d(golden_dir, name):
    g = _npz(golden_dir, 'fir.npz')
    up, down, px0, px1, py0, py1 = (int(v) for v in g[f'{name}_p'])
    y = O.upfirdn2d_ref(torch.from_numpy(g[f'{name}_x']), torch.from_numpy(g[f'{name}_k']),
                        up, up, down, down, px0, px1, py0, py1)
    assert y.shape == g[f'{name}_y'].shape
    np.testing.assert_allclose(y.numpy(), g[f'{name}_y'], rtol=0, atol=2e-6)


def test_upfirdn2d_asymmetric(golden_dir):
    g = _npz(golden_dir, 'fir.npz')
    p = [int(v) for v in g['asym_p']]
    y = O.upfirdn2d_ref(torch.from_numpy(g['asym_x']), torch.from_numpy(g['asym_k']), *p)
    assert y.shape == g['asym_y'].shape
    np.testing.assert_allclose(y.numpy(), g['asym_y'], rtol=0, atol=2e-6)


def test_upfirdn2d_empty():
    x = torch.zeros(0, 3, 8, 8)
    y = O.upfirdn2d(x, torch.ones(2, 2), down=2)
    assert y.shape == (0, 3, 4, 4)


def test_fused_leaky_relu(golden_dir):
    g = _npz(golden_dir, 'fir.npz')
    y = O.fused_leaky_relu_ref(torch.from_numpy(g['lrelu_x']), torch.from_numpy(g['lrelu_b']))
    np.testing.assert_allclose(y.numpy(), g['lrelu_y'], rtol=0, atol=1e-6)


@pytest.mark.parametrize('which,tag', [('main', 'nf64_s32'), ('main', 'nf16_s64'),
                                       ('healthy', 'nf64_s32'), ('healthy', 'nf16_s64')])
def test_generators_match_reference(golden_dir, which, tag):
    g = _npz(golden_dir, f'gen_{which}.npz')
    nf, size = (64, 32) if tag == 'nf64_s32' else (16, 64)
    batch = 2 if tag == 'nf64_s32' else 1
    v1, v2, ncond = ('g1', 'g2', 3) if which == 'main' else ('g1_healthy', 'g2_healthy', 2)
    cfg = O.default_config(num_channels_dae=nf, image_size=size)
    sd1, sd2 = O.make_state_dict(cfg, v1, seed=0), O.make_state_dict(cfg, v2, seed=1)
    conds, x_init, latents, noises = O.synthetic_inputs(batch, size, cfg, ncond=ncond, seed=42)
    t = torch.tensor([3, 1][:batch], dtype=torch.int64)
    y1 = O.generator_forward(sd1, cfg, v1, x_init, conds, t, latents[0])
    y2 = O.generator_forward(sd2, cfg, v2, x_init, conds, t, latents[0], pseudo_target=y1[:, [0], :])
    # fp32, same ATen kernels, slightly different op grouping: 1e-5 absolute on tanh outputs
    np.testing.assert_allclose(y1.numpy(), g[f'{tag}_g1'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(y2.numpy(), g[f'{tag}_g2'], rtol=0, atol=1e-5)
    assert np.abs(g[f'{tag}_g1']).mean() > 0.1      # non-degenerate


def test_sampling_loop_matches_reference(golden_dir):
    g = _npz(golden_dir, 'gen_main.npz')
    cfg = O.default_config(num_channels_dae=64, image_size=32)
    sd1, sd2 = O.make_state_dict(cfg, 'g1', seed=0), O.make_state_dict(cfg, 'g2', seed=1)
    conds, x_init, latents, noises = O.synthetic_inputs(2, 32, cfg, seed=42)
    x = O.sample_from_model(O.PosteriorCoefficients(cfg), sd1, sd2, cfg, conds, x_init, latents, noises)
    np.testing.assert_allclose(x.numpy(), g['nf64_s32_sample'], rtol=0, atol=1e-5)


@pytest.mark.parametrize('name', ['mri', 'smooth', 'flat', 'zeros', 'two_values'])
def test_volume_oracle_matches_reference_golden(golden_dir, name):
    """oracle/volume_oracle.py == outputs of the reference's own engine/test_volume.py functions
    (tests/golden/make_volume_golden.py ran them), bit for bit."""
    from oracle import volume_oracle as VO
    g = np.load(os.path.join(golden_dir, 'volume.npz'))
    vol = g[f'{name}_vol']
    half, size, s0, s1 = (int(v) for v in g[f'{name}_p'])
    np.testing.assert_array_equal(VO.robust_minmax_to_minus1_1(vol), g[f'{name}_norm'])
    conds, a, b = VO.preprocess_volume(vol, half, size)
    assert (a, b) == (s0, s1)
    np.testing.assert_array_equal(conds.numpy(), g[f'{name}_conds'])
    pred = VO.postprocess_slices(torch.from_numpy(g[f'{name}_fake']))
    np.testing.assert_array_equal(VO.reconstruct_volume_from_slices(list(pred), vol.shape, s0, s1), g[f'{name}_rebuilt'])


def test_testset_oracle_vs_reference_fixture(golden_dir):
    """oracle/testset_oracle.py against tests/golden/testset.npz, which make_testset_golden.py produced by running the
    reference's BratsDataset.__getitem__ (dataset/dataset_brats.py:73-92) and the export block of engine/test.py:367-388."""
    from oracle import testset_oracle as TO
    g = _npz(golden_dir, 'testset.npz')
    for target in ('T1CE', 'FLAIR', 'T2', 'T1'):
        order = list(g[f'{target}_order'])
        for j, m in enumerate(order[:-1]):
            np.testing.assert_array_equal(TO.zscore_to_unit(g[f'in_{m}']).numpy(), g[f'{target}_cond'][:, j])
        np.testing.assert_array_equal(TO.zscore_to_unit(g[f'in_{order[-1]}']).numpy(), g[f'{target}_target'][:, 0])
    for tag in ('a', 'b', 'const'):
        p8, g8, (lo, hi) = TO.export_uint8(list(g[f'exp_{tag}_pred']), list(g[f'exp_{tag}_gt']))
        np.testing.assert_array_equal(p8, g[f'exp_{tag}_p8'])
        np.testing.assert_array_equal(g8, g[f'exp_{tag}_g8'])
        assert (lo, hi) == tuple(g[f'exp_{tag}_win'])


def test_discriminator_oracle_vs_reference_fixture(golden_dir):
    """oracle/disc_oracle.py against tests/golden/disc.npz (the reference's Discriminator_large run on CPU)."""
    from oracle import disc_oracle as DO
    g = _npz(golden_dir, 'disc.npz')
    for tag, ngf, temb, size, batch in (('ngf16_s128_b8', 16, 128, 128, 8), ('ngf64_s64_b4', 64, 256, 64, 4),
                                        ('ngf16_s64_b2', 16, 64, 64, 2)):
        sd = DO.make_state_dict(nc=2, ngf=ngf, t_emb_dim=temb, seed=3)
        x, x_t, t = disc_inputs(g, tag, size, batch)
        logits, mid = DO.discriminator_forward(sd, x, t, x_t)
        np.testing.assert_allclose(logits.numpy(), g[f'{tag}_logits'], rtol=0, atol=2e-6)
        np.testing.assert_allclose(mid.numpy(), g[f'{tag}_mid'], rtol=0, atol=2e-6)


def disc_inputs(g, tag, size, batch):
    """Re-draw the inputs make_disc_golden.py used (same CPU generator seed), guarded by the stored sums."""
    gen = torch.Generator().manual_seed(17)
    x, x_t = torch.randn(batch, 1, size, size, generator=gen), torch.randn(batch, 1, size, size, generator=gen)
    np.testing.assert_allclose([x.double().sum().item(), x_t.double().sum().item()], g[f'{tag}_xsum'], rtol=0, atol=1e-9)
    return x, x_t, torch.from_numpy(g[f'{tag}_t'])

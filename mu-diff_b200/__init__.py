"""mu-diff_b200: B200-native (sm_100a) implementation of MU-Diff's reverse-sampling hot path.

Import as `importlib.import_module('mu-diff_b200')` or through the `mudiff_b200` alias
package at the repository root.  Layout mirrors the reference for the path only:

    op/                      utils/op            upfirdn2d, fused_leaky_relu, FusedLeakyReLU
    up_or_down_sampling.py   backbones/up_or_down_sampling.py
    layers.py, dense_layer.py, layerspp.py       backbones/*
    ncsnpp_generator_adagn_feat[_healthy].py     NCSNpp, NCSNpp_adaptive
    sampling.py              engine/test.py:48-199 (+ GraphSampler, StreamingSampler)
    volume.py                engine/test_volume.py:135-181,269-294 (sharded, batched, GPU pre/post)
    testset.py               engine/test.py:265-400 (batched slice-test driver, uint8 export)
    validation.py            engine/train.py:1148-1175 (validation sampling on weights shared with training)
    discriminator.py         backbones/discriminator.py (Discriminator_large forward)
    csrc/, libmudiff_b200.so C ABI (include/mudiff_b200.h)
"""
from . import _lib, ops  # noqa: F401
from . import op  # noqa: F401
from . import layers, dense_layer, up_or_down_sampling, layerspp  # noqa: F401
from . import ncsnpp_generator_adagn_feat, ncsnpp_generator_adagn_feat_healthy  # noqa: F401
from . import discriminator  # noqa: F401
from .ncsnpp_generator_adagn_feat import NCSNpp, NCSNpp_adaptive  # noqa: F401
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d, upfirdn2d_ada  # noqa: F401
from .sampling import (GraphSampler, StreamingSampler, Posterior_Coefficients, get_sigma_schedule, get_time_schedule,  # noqa: F401
                       sample_from_model, sample_posterior_combine)

__version__ = '0.1.0'

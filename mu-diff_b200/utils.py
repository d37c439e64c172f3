"""Model registry decorators (backbones/utils.py:10-30).  Unlike the reference, registering
the same name twice replaces the entry instead of raising, so the main and the 'healthy'
generator files can live in one process."""
_MODELS = {}


def register_model(cls=None, *, name=None):
    def _register(cls):
        _MODELS[name if name is not None else cls.__name__] = cls
        return cls
    return _register if cls is None else _register(cls)


def get_model(name):
    return _MODELS[name]


def randomize_(module, seed=0):
    """Deterministic NON-degenerate weights for benchmarking / smoke runs: the reference's default
    init gives ~1e-5 outputs and all-zero biases.  Every >=2-D parameter is
    fan-avg uniform at scale 1, biases N(0, 0.1^2), AdaGN style biases [1..1, 0..0] + noise."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.ndim >= 2:
                rf = 1
                for d in p.shape[2:]:
                    rf *= d
                fan_in, fan_out = p.shape[1] * rf, p.shape[0] * rf
                if name.endswith('.W'):            # NIN: [in, out]
                    fan_in, fan_out = p.shape[0], p.shape[1]
                a = math.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
                p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) * a).to(p.device))
            else:
                v = torch.randn(p.shape, generator=g) * 0.1
                if name.endswith('style.bias'):
                    v[:p.shape[0] // 2] += 1.0
                elif name.endswith('GroupNorm_0.weight') or (name.endswith('.weight') and p.ndim == 1):
                    v += 1.0
                p.copy_(v.to(p.device))
    return module

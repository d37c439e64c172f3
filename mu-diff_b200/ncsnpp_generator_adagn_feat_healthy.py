"""2-contrast ('healthy') generators with the reference's signatures
(backbones/ncsnpp_generator_adagn_feat_healthy.py:279 and :693)."""
from . import utils
from ._generator import PixelNorm, _NCSNppBase  # noqa: F401


@utils.register_model(name='ncsnpp')
class NCSNpp(_NCSNppBase):
    adaptive, n_cond = False, 2

    def forward(self, x, cond1, cond2, time_cond, z):
        return self._forward(x, (cond1, cond2), time_cond, z)


@utils.register_model(name='ncsnpp_adaptive')
class NCSNpp_adaptive(_NCSNppBase):
    adaptive, n_cond = True, 2

    def forward(self, x, cond1, cond2, time_cond, z, pseudo_target):
        return self._forward(x, (cond1, cond2), time_cond, z, pseudo_target)

"""`NCSNpp` (G1) and `NCSNpp_adaptive` (G2) with the reference's constructor and forward
signatures (backbones/ncsnpp_generator_adagn_feat.py:53,279 and :451,694)."""
from . import utils
from ._generator import PixelNorm, _NCSNppBase  # noqa: F401


@utils.register_model(name='ncsnpp')
class NCSNpp(_NCSNppBase):
    """NCSN++ model (3 conditioning contrasts)."""
    adaptive, n_cond = False, 3

    def forward(self, x, cond1, cond2, cond3, time_cond, z):
        return self._forward(x, (cond1, cond2, cond3), time_cond, z)


@utils.register_model(name='ncsnpp_adaptive')
class NCSNpp_adaptive(_NCSNppBase):
    """NCSN++ model with pseudo-target adaptive stem and cross-contrast gating."""
    adaptive, n_cond = True, 3

    def forward(self, x, cond1, cond2, cond3, time_cond, z, pseudo_target):
        return self._forward(x, (cond1, cond2, cond3), time_cond, z, pseudo_target)

"""Generate tests/golden/disc.npz by RUNNING THE REFERENCE's Discriminator_large (backbones/discriminator.py:175-263) on
CPU (build container only; /root/reference does not exist on the GPU box):

    python tests/golden/make_disc_golden.py

Weights: oracle/disc_oracle.make_state_dict (deterministic, non-degenerate), loaded with strict=True => key/shape parity.
Stored: inputs (x, x_t, t), logits and mid_feat for ngf=16 @ 128^2 B=8 (two sub-batches of the minibatch-stddev group of
4), ngf=64 @ 64^2 B=4, and ngf=16 @ 64^2 B=2 (group = batch < 4).
Run with TORCH_EXTENSIONS_DIR=/tmp/ref_ext (NOT baseline/_ref/torch_ext: importing /root/reference's utils.op would
re-target that directory's ninja files).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
from oracle import disc_oracle as DO  # noqa: E402

torch.set_num_threads(8)


def main():
    from backbones.discriminator import Discriminator_large
    out = {}
    for tag, ngf, temb, size, batch in (('ngf16_s128_b8', 16, 128, 128, 8), ('ngf64_s64_b4', 64, 256, 64, 4),
                                        ('ngf16_s64_b2', 16, 64, 64, 2)):
        sd = DO.make_state_dict(nc=2, ngf=ngf, t_emb_dim=temb, seed=3)
        D = Discriminator_large(nc=2, ngf=ngf, t_emb_dim=temb, act=torch.nn.LeakyReLU(0.2)).eval()
        D.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(17)
        x, x_t = torch.randn(batch, 1, size, size, generator=g), torch.randn(batch, 1, size, size, generator=g)
        t = torch.randint(0, 4, (batch,), generator=g)
        with torch.no_grad():
            logits, mid = D(x, t, x_t)
        o_logits, o_mid = DO.discriminator_forward(sd, x, t, x_t)
        print(tag, 'oracle-vs-reference max|d|', (o_logits - logits).abs().max().item(), (o_mid - mid).abs().max().item(),
              '|logits| max', logits.abs().max().item(), '|mid| max', mid.abs().max().item())
        out[f'{tag}_t'] = t.numpy()                 # x / x_t are re-drawn by the tests from the same CPU generator seed
        out[f'{tag}_xsum'] = np.array([x.double().sum().item(), x_t.double().sum().item()])   # guards that re-draw
        out[f'{tag}_logits'], out[f'{tag}_mid'] = logits.numpy(), mid.numpy()
    np.savez_compressed(os.path.join(HERE, 'disc.npz'), **out)
    print('disc.npz')


if __name__ == '__main__':
    main()

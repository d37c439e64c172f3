"""Validation-time sampling inside training through the fast path (engine/train.py:1148-1175 runs
the 4-step sampler over the whole validation split every epoch, on every rank, with the TRAINING modules at B = 1).

The fast generators are inference-only modules with the reference's state_dict keys, so they can ALIAS the parameters of
the (DDP-wrapped) training generators instead of copying them: after `share_weights` an optimiser step on the training
module is what the fast module sees.  Kernel-ready copies of the weights (packed bf16 matrices, layers.PackCache) are
keyed by the parameters' version counters and rebuilt lazily after each update; a captured CUDA graph embeds those
copies, so `ValidationSampler` re-captures its graph when any parameter changed since the capture.
"""
from typing import Iterable, Optional

import torch
from torch import nn

from .sampling import GraphSampler, Posterior_Coefficients


def _unwrap(module: nn.Module) -> nn.Module:
    return module.module if hasattr(module, 'module') and isinstance(module.module, nn.Module) else module


def share_weights(fast: nn.Module, training: nn.Module) -> nn.Module:
    """Make every parameter / buffer of `fast` alias the tensor of the same name in `training` (a plain module or a
    DistributedDataParallel wrapper; train.py:1135 saves `module.`-less keys the same way).  No copy, strict keys."""
    src = _unwrap(training)
    state = {k: v.detach() for k, v in src.state_dict(keep_vars=True).items()}
    fast.load_state_dict(state, strict=True, assign=True)
    for p in fast.parameters():
        p.requires_grad_(False)
    return fast.eval()


def params_signature(modules: Iterable[nn.Module]):
    return tuple((p.data_ptr(), p._version) for m in modules for p in m.parameters())


class ValidationSampler:
    """Batched 4-step sampler for the validation loop: `sample(conds, x_init, latents, noises)` replays one CUDA graph
    at a fixed batch and re-captures it when the shared weights changed (once per epoch in train.py's schedule)."""

    def __init__(self, args, fast_g1: nn.Module, fast_g2: nn.Module, batch: int, size: int, n_cond: int = 3,
                 device='cuda', coefficients: Optional[Posterior_Coefficients] = None):
        self.args, self.g1, self.g2 = args, fast_g1, fast_g2
        self.batch, self.size, self.n_cond, self.device = int(batch), int(size), n_cond, torch.device(device)
        self.co = coefficients if coefficients is not None else Posterior_Coefficients(args, self.device)
        self._gs: Optional[GraphSampler] = None
        self._sig = None

    def _ensure(self):
        sig = params_signature((self.g1, self.g2))
        if self._gs is None or sig != self._sig:
            self._gs = GraphSampler(self.co, self.g1, self.g2, self.args.num_timesteps, self.batch, self.size, self.args.nz,
                                    n_cond=self.n_cond, device=self.device, warmup=1)
            self._sig = sig
        return self._gs

    def sample(self, conds, x_init, latents, noises) -> torch.Tensor:
        gs = self._ensure()
        b = x_init.shape[0]
        if b > self.batch:
            raise RuntimeError(f"mu-diff_b200: ValidationSampler captured for batch {self.batch}, got {b}")
        for d, s in zip(gs.conds, conds):
            d[:b].copy_(s, non_blocking=True)
        gs.x_init[:b].copy_(x_init, non_blocking=True)
        for d, s in zip(gs.latents, latents):
            d[:b].copy_(s, non_blocking=True)
        for d, s in zip(gs.noises, noises):
            d[:b].copy_(s, non_blocking=True)
        return gs.replay()[:b].clone()

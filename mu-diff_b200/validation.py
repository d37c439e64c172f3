"""Validation-time sampling inside training through the fast path (engine/train.py:1148-1175 runs
the 4-step sampler over the whole validation split every epoch, on every rank, with the TRAINING modules at B = 1).

The fast generators are inference-only modules with the reference's state_dict keys.  Two ways to feed them the
training weights:

  mirror_weights(fast, training)   (default of ValidationSampler)  the fast module OWNS its parameter storage and the
        training values are copied into it (`copy_`, ~170 MB device-to-device for G1 + G2).  Addresses never change, so
        ONE captured CUDA graph stays valid for the whole training run: after a weight update the sampler copies the
        new values, refreshes the kernel-ready packed copies in place (layers.refresh_packs) and replays.  This is
        safe under everything the reference's loop does to its parameters: in-place optimizer steps, DDP, and the EMA
        wrapper's `p.data = ema` REBINDING (utils/EMA.py:86-90, called at train.py:1126 and :1139).
  share_weights(fast, training)    the fast module ALIASES the training tensors (no copy, no extra memory).  Only
        in-place updates are followed for free; if a training parameter is rebound to other storage (EMA swap) the
        sampler notices (it compares the SOURCE parameters' data_ptr / _version, not just its own), re-aliases,
        drops the packed copies and re-captures its graph.
"""
from typing import Optional, Sequence

import torch
from torch import nn

from . import layers
from .sampling import GraphSampler, Posterior_Coefficients


def _unwrap(module: nn.Module) -> nn.Module:
    return module.module if hasattr(module, 'module') and isinstance(module.module, nn.Module) else module


def _source_state(training: nn.Module):
    """`module.`-less state of a plain module or a DistributedDataParallel wrapper (train.py:1135 saves the same keys)."""
    return {k: v.detach() for k, v in _unwrap(training).state_dict(keep_vars=True).items()}


def share_weights(fast: nn.Module, training: nn.Module) -> nn.Module:
    """Make every parameter / buffer of `fast` alias the tensor of the same name in `training`.  No copy, strict keys.
    Constraint: follows in-place updates only - see the module docstring (ValidationSampler(mode='alias') handles
    rebinding by re-aliasing)."""
    fast.load_state_dict(_source_state(training), strict=True, assign=True)
    for m in fast.modules():                       # packed copies were built from the previous storage
        m.__dict__.pop('_pack_cache', None)
    for p in fast.parameters():
        p.requires_grad_(False)
    return fast.eval()


def mirror_weights(fast: nn.Module, training: nn.Module) -> nn.Module:
    """Copy the training values into `fast`'s own storage (strict keys, dtype / device of `fast` kept) and refresh its
    packed copies in place.  Never changes an address a captured graph may hold."""
    src = _source_state(training)
    dst = fast.state_dict(keep_vars=True)
    if set(src) != set(dst):
        missing, extra = sorted(set(dst) - set(src)), sorted(set(src) - set(dst))
        raise RuntimeError(f"mu-diff_b200: state_dict keys differ (missing {missing[:3]}..., unexpected {extra[:3]}...)")
    with torch.no_grad():
        for k, d in dst.items():
            if d.data_ptr() != src[k].data_ptr():
                d.copy_(src[k], non_blocking=True)
    for p in fast.parameters():
        p.requires_grad_(False)
    layers.refresh_packs(fast)
    return fast.eval()


def params_signature(modules: Sequence[nn.Module]):
    return tuple((p.data_ptr(), p._version) for m in modules for p in _unwrap(m).parameters())


class ValidationSampler:
    """Batched 4-step sampler for the validation loop: `sample(conds, x_init, latents, noises)` replays one CUDA graph at
    a fixed batch.  With `sources=(training_g1, training_g2)` it tracks the TRAINING modules: whenever one of their
    parameters was updated in place or rebound since the last call, the fast modules are brought up to date first
    (mode 'mirror': copy + in-place pack refresh, the graph is kept; mode 'alias': see share_weights).  Without
    `sources` the fast modules' own parameters are watched (weights loaded / modified by the caller)."""

    def __init__(self, args, fast_g1: nn.Module, fast_g2: nn.Module, batch: int, size: int, n_cond: int = 3,
                 device='cuda', coefficients: Optional[Posterior_Coefficients] = None,
                 sources: Optional[Sequence[nn.Module]] = None, mode: str = 'mirror'):
        if mode not in ('mirror', 'alias'):
            raise ValueError("mode must be 'mirror' or 'alias'")
        self.args, self.g1, self.g2 = args, fast_g1, fast_g2
        self.batch, self.size, self.n_cond, self.device = int(batch), int(size), n_cond, torch.device(device)
        self.co = coefficients if coefficients is not None else Posterior_Coefficients(args, self.device)
        self.sources, self.mode = (tuple(sources) if sources is not None else None), mode
        self._gs: Optional[GraphSampler] = None
        self._sig = None            # signature of the fast modules' parameters the graph / packs were built for
        self._src_sig = None
        self.captures = 0           # how many times a graph was captured (tests, logging)

    def _follow_sources(self):
        if self.sources is None:
            return
        src_sig = params_signature(self.sources)
        if src_sig == self._src_sig:
            return
        for fast, src in zip((self.g1, self.g2), self.sources):
            if self.mode == 'mirror':
                mirror_weights(fast, src)
            else:
                fs, ss = fast.state_dict(keep_vars=True), _source_state(src)
                if any(fs[k].data_ptr() != ss[k].data_ptr() for k in fs):      # rebound (EMA swap) or never shared
                    share_weights(fast, src)
                    self._gs = None                                            # addresses changed: capture again
        self._src_sig = src_sig

    def _ensure(self):
        self._follow_sources()
        sig = params_signature((self.g1, self.g2))
        if self._gs is not None and sig != self._sig:
            if tuple(p for p, _ in sig) == tuple(p for p, _ in self._sig):
                layers.refresh_packs(self.g1)      # same storage, new values: packed copies rebuilt in place,
                layers.refresh_packs(self.g2)      # the captured graph stays valid
            else:
                self._gs = None
        if self._gs is None:
            self._gs = GraphSampler(self.co, self.g1, self.g2, self.args.num_timesteps, self.batch, self.size, self.args.nz,
                                    n_cond=self.n_cond, device=self.device, warmup=1)
            self.captures += 1
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

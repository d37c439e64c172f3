"""The reverse-sampling loop of MU-Diff with the reference's names and argument meaning
(engine/test.py:48-199; identical copies in engine/train.py:363-375 and
engine/test_volume.py:109-129):

    get_time_schedule, get_sigma_schedule, Posterior_Coefficients,
    sample_posterior_combine(coefficients, x_0_1, x_0_2, x_t, t),
    sample_from_model(coefficients, generator1, cond1, generator2, cond2, cond3, n_time, x_init, T, opt)

plus `GraphSampler`, which captures the whole n_time-step loop (8 generator forwards + 4
posterior kernels for n_time = 4) in ONE CUDA graph and replays it on static buffers.
"""
import numpy as np
import torch

from . import _lib as L
from . import ops


def var_func_vp(t, beta_min, beta_max):
    log_mean_coeff = -0.25 * t ** 2 * (beta_max - beta_min) - 0.5 * t * beta_min
    return 1. - torch.exp(2. * log_mean_coeff)


def var_func_geometric(t, beta_min, beta_max):
    return beta_min * ((beta_max / beta_min) ** t)


def extract(input, t, shape):
    out = torch.gather(input, 0, t)
    return out.reshape(*([shape[0]] + [1] * (len(shape) - 1)))


def _time_grid(n_timestep):
    eps_small = 1e-3
    t = np.arange(0, n_timestep + 1, dtype=np.float64) / n_timestep
    return torch.from_numpy(t) * (1. - eps_small) + eps_small


def get_time_schedule(args, device):
    return _time_grid(args.num_timesteps).to(device)


def get_sigma_schedule(args, device):
    """engine/test.py:75-97 (host-side, float64 grid -> fp32 betas with a leading 1e-8)."""
    t = _time_grid(args.num_timesteps)
    if getattr(args, 'use_geometric', False):
        var = var_func_geometric(t, args.beta_min, args.beta_max)
    else:
        var = var_func_vp(t, args.beta_min, args.beta_max)
    alpha_bars = 1.0 - var
    betas = 1 - alpha_bars[1:] / alpha_bars[:-1]
    betas = torch.cat((torch.tensor(1e-8)[None], betas)).to(device).type(torch.float32)
    return betas ** 0.5, torch.sqrt(1 - betas), betas


class Posterior_Coefficients():
    """engine/test.py:101-123.  The tables are computed on the host in fp32 exactly as the
    reference does (so they are bit-identical), then kept on `device` for the fused kernel."""

    def __init__(self, args, device):
        _, _, betas = get_sigma_schedule(args, device='cpu')
        self.betas = betas.type(torch.float32)[1:]
        self.alphas = 1 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, 0)
        self.alphas_cumprod_prev = torch.cat((torch.tensor([1.], dtype=torch.float32), self.alphas_cumprod[:-1]), 0)
        self.posterior_variance = self.betas * (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.rsqrt(self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1 / self.alphas_cumprod - 1)
        self.posterior_mean_coef1 = self.betas * torch.sqrt(self.alphas_cumprod_prev) / (1 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1 - self.alphas_cumprod_prev) * torch.sqrt(self.alphas) / (1 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = torch.log(self.posterior_variance.clamp(min=1e-20))
        for k, v in list(vars(self).items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device).contiguous())


def sample_posterior_combine(coefficients, x_0_1, x_0_2, x_t, t, noise=None):
    """engine/test.py:150-177 as one fused kernel.  `noise` defaults to torch.randn_like(x_t)
    (drawn even at t == 0 and masked, like the reference, so the RNG stream stays aligned)."""
    if noise is None:
        noise = torch.randn_like(x_t)
    return ops.posterior_update(x_0_1, x_0_2, x_t, noise, t, coefficients.posterior_mean_coef1,
                                coefficients.posterior_mean_coef2, coefficients.posterior_log_variance_clipped)


def _ch0(x):
    """`x[:, [0], :]` of engine/test.py:193-195 as a view (no gather kernel; identical values)."""
    return x[:, 0:1]


def sample_from_model(coefficients, generator1, cond1, generator2, cond2, cond3, n_time, x_init, T, opt,
                      latents=None, noises=None):
    """engine/test.py:180-199.  `T` is accepted and unused, as in the reference.  For the
    2-contrast generators pass cond3=None.  `latents[i]` / `noises[i]` (indexed by the step
    index i) replace the reference's torch.randn draws when given (oracle parity, graph capture).
    The reference's autocast() context is replaced by the generators' own precision switch."""
    x = x_init
    conds = [c for c in (cond1, cond2, cond3) if c is not None]
    with torch.no_grad(), ops.stem_moments_scope():
        for i in reversed(range(n_time)):
            t = torch.full((x.size(0),), i, dtype=torch.int64, device=x.device)
            latent_z = latents[i] if latents is not None else torch.randn(x.size(0), opt.nz, device=x.device)
            x_0_1 = generator1(x, *conds, t, latent_z)
            x_0_2 = generator2(x, *conds, t, latent_z, _ch0(x_0_1))
            x_new = sample_posterior_combine(coefficients, _ch0(x_0_1), _ch0(x_0_2), x, t,
                                             noise=noises[i] if noises is not None else None)
            x = x_new.detach()
    return x


class GraphSampler:
    """Whole-loop CUDA graph.  Static buffers: conds, x_init, latents[n_time], noises[n_time];
    call `run(...)` with new contents (copied in) or fill the buffers yourself and `replay()`."""

    def __init__(self, coefficients, generator1, generator2, n_time, batch, size, nz, n_cond=3, device='cuda',
                 warmup=2):
        dev = torch.device(device)
        self.co, self.g1, self.g2, self.n_time = coefficients, generator1, generator2, n_time
        self.conds = [torch.zeros(batch, 1, size, size, device=dev) for _ in range(n_cond)]
        self.x_init = torch.zeros(batch, 1, size, size, device=dev)
        self.latents = [torch.zeros(batch, nz, device=dev) for _ in range(n_time)]
        self.noises = [torch.zeros(batch, 1, size, size, device=dev) for _ in range(n_time)]
        self.ts = [torch.full((batch,), i, dtype=torch.int64, device=dev) for i in range(n_time)]
        self.out = None
        self.graph = None
        self.launches_per_replay = 0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # packs weights, warms allocator
                self._loop()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = self._loop()
        self.launches_per_replay = L.launch_count() - n0

    def _loop(self):
        x = self.x_init
        with torch.no_grad(), ops.stem_moments_scope():
            for i in reversed(range(self.n_time)):
                x01 = self.g1(x, *self.conds, self.ts[i], self.latents[i])
                x02 = self.g2(x, *self.conds, self.ts[i], self.latents[i], _ch0(x01))      # engine/test.py:193
                x = sample_posterior_combine(self.co, _ch0(x01), _ch0(x02), x, self.ts[i], noise=self.noises[i])
        return x

    def load(self, conds, x_init, latents, noises, non_blocking=True):
        for d, s in zip(self.conds, conds):
            d.copy_(s, non_blocking=non_blocking)
        self.x_init.copy_(x_init, non_blocking=non_blocking)
        for d, s in zip(self.latents, latents):
            d.copy_(s, non_blocking=non_blocking)
        for d, s in zip(self.noises, noises):
            d.copy_(s, non_blocking=non_blocking)

    def replay(self):
        """Replay the graph; returns the graph's STATIC output buffer (overwritten by the next replay - copy it if it
        has to outlive that, as `run` does)."""
        self.graph.replay()
        return self.out

    def run(self, conds, x_init, latents, noises):
        """Copy new inputs in, replay, return a fresh tensor owned by the caller (like the reference's loop)."""
        self.load(conds, x_init, latents, noises)
        return self.replay().clone()


class StreamingSampler:
    """Host pipeline around a `GraphSampler` for the call pattern of engine/test.py:294-331 (a loader hands over one batch of
    conditioning slices after the other, every result goes back to the host): the H2D copy of batch i+1 and the D2H copy of
    batch i-1 run on two copy streams WHILE the graph of batch i replays.  Two device staging slots per direction; the
    graph's static buffers are filled / drained by device-to-device copies on the compute stream (tens of microseconds).

        ss = StreamingSampler(gs)
        for conds_host, out_host in batches:          # pinned host tensors
            ss.submit(conds_host, out_host)           # returns at once; out_host is valid after ss.synchronize()
        ss.synchronize()

    x_init / z / noise are drawn on the device before every replay (engine/test.py:188,331) with `generator`, in the same
    order as a sequential `gs.load`-free loop would draw them, so the pipelined outputs equal the sequential ones bit for bit.
    """

    def __init__(self, gs, generator=None):
        dev = gs.x_init.device
        self.gs, self.generator, self.dev = gs, generator, dev
        self.s_in, self.s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.c_stage = [[torch.empty_like(c) for c in gs.conds] for _ in range(2)]
        self.o_stage = [torch.empty_like(gs.out) for _ in range(2)]
        self.c_free = [torch.cuda.Event() for _ in range(2)]
        self.c_full = [torch.cuda.Event() for _ in range(2)]
        self.o_free = [torch.cuda.Event() for _ in range(2)]
        self.o_full = [torch.cuda.Event() for _ in range(2)]
        self.i = 0

    def draw_noise(self):
        gs, g = self.gs, self.generator
        gs.x_init.normal_(generator=g)
        for t in gs.latents:
            t.normal_(generator=g)
        for t in gs.noises:
            t.normal_(generator=g)

    def submit(self, conds_host, out_host, draw=True):
        gs, s = self.gs, self.i & 1
        n = conds_host[0].shape[0]
        comp = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):                    # H2D of this batch: overlaps the previous batch's replay
            if self.i >= 2:
                self.s_in.wait_event(self.c_free[s])
            for d, h in zip(self.c_stage[s], conds_host):
                d[:n].copy_(h, non_blocking=True)
            self.c_full[s].record(self.s_in)
        comp.wait_event(self.c_full[s])
        for d, st in zip(gs.conds, self.c_stage[s]):
            d[:n].copy_(st[:n], non_blocking=True)
        self.c_free[s].record(comp)
        if draw:
            self.draw_noise()
        y = gs.replay()
        if self.i >= 2:
            comp.wait_event(self.o_free[s])
        self.o_stage[s][:n].copy_(y[:n], non_blocking=True)
        self.o_full[s].record(comp)
        with torch.cuda.stream(self.s_out):                   # D2H of this batch: overlaps the next batch's replay
            self.s_out.wait_event(self.o_full[s])
            out_host.copy_(self.o_stage[s][:n], non_blocking=True)
            self.o_free[s].record(self.s_out)
        self.i += 1

    def synchronize(self):
        self.s_out.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()

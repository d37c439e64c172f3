"""Harness around the UNMODIFIED reference (MarioPasc/MU-Diff) for the benchmark's baselines.  Nothing in the product
package imports this file; `bench.py` (reference arm, `reference_gpu` record), `tools/ref_gpu_baseline.py` and the
tests that compare against the reference on the GPU do.

`install()` (called by `__graft_entry__.build()` in the build container, where /root/reference exists) copies the
reference's Python packages that the sampling path needs - backbones/, utils/, dataset/, engine/test.py - into
`baseline/_ref/` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so) and pre-builds the
reference's two JIT CUDA extensions (utils/op/upfirdn2d.py:21-30, utils/op/fused_act.py:23-32) for sm_100a into
`baseline/_ref/torch_ext/`.  The reference has no setup.py / pyproject.toml, so `pip install --target` does not apply;
a copy of the script tree is the reference's own "installation" (README.md: clone and run).

The loop `engine/test.py:180-199` (`sample_from_model`) is exec'd from the reference's SOURCE TEXT (the module itself
imports skimage / matplotlib, which this image does not have) - the same technique tests/golden/make_golden.py uses.
"""
import ast
import os
import shutil
import sys
import time
import types
from argparse import Namespace
from contextlib import nullcontext

HERE = os.path.dirname(os.path.abspath(__file__))
_CANON = '/root/repo/baseline/_ref'          # on the GPU box /root/repo is a symlink to the snapshot: keep ONE path so
REF_DIR = os.path.join(HERE, '_ref')         # that the pre-built extensions' ninja files stay valid
if os.path.isdir(_CANON) and os.path.isdir(REF_DIR) and os.path.samefile(_CANON, REF_DIR):
    REF_DIR = _CANON
EXT_DIR = os.path.join(REF_DIR, 'torch_ext')
SRC = '/root/reference'
_COPY = ['backbones', 'utils', 'dataset', os.path.join('engine', 'test.py')]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, 'backbones', 'ncsnpp_generator_adagn_feat.py'))


def install(prebuild: bool = True) -> bool:
    """Copy the reference's packages into baseline/_ref (build container only) and pre-build its CUDA extensions."""
    if not os.path.isdir(SRC):
        return available()
    for rel in _COPY:
        s, d = os.path.join(SRC, rel), os.path.join(REF_DIR, rel)
        if os.path.isdir(s):
            if not os.path.isdir(d):
                shutil.copytree(s, d, ignore=shutil.ignore_patterns('__pycache__'))
        elif not os.path.isfile(d):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    for dp, dn, fn in os.walk(REF_DIR):            # /root/reference is read-only; the copy must not be
        for n in dn + fn:
            q = os.path.join(dp, n)
            os.chmod(q, os.stat(q).st_mode | 0o200)
    os.makedirs(EXT_DIR, exist_ok=True)
    if prebuild and not (os.path.exists(os.path.join(EXT_DIR, 'upfirdn2d.so')) and os.path.exists(os.path.join(EXT_DIR, 'fused.so'))):
        import subprocess
        env = dict(os.environ, TORCH_EXTENSIONS_DIR=EXT_DIR, TORCH_CUDA_ARCH_LIST='10.0a', PYTHONPATH=REF_DIR)
        subprocess.run([sys.executable, '-c', 'import utils.op'], cwd=REF_DIR, env=env, check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return available()


def import_reference(healthy: bool = False):
    """Import the reference's generator module from baseline/_ref (its JIT extensions load from the pre-built
    baseline/_ref/torch_ext, or re-build there with nvcc if the box invalidated them).  The two generator files register
    the same model names: one variant per process."""
    if not available():
        raise RuntimeError("baseline/_ref is not installed (run __graft_entry__.build() in the build container)")
    os.environ.setdefault('TORCH_EXTENSIONS_DIR', EXT_DIR)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    name = 'backbones.ncsnpp_generator_adagn_feat' + ('_healthy' if healthy else '')
    return importlib.import_module(name)


def engine_symbols(autocast_ctx=None):
    """The pure-torch part of engine/test.py (:48-199) exec'd from the source text.  `autocast_ctx` is a zero-argument
    callable returning the context manager that stands for the reference's `autocast()` (:191)."""
    import numpy as np
    import torch
    src = open(os.path.join(REF_DIR, 'engine', 'test.py')).read()
    tree = ast.parse(src)
    keep = {'var_func_vp', 'var_func_geometric', 'extract', 'get_time_schedule', 'get_sigma_schedule',
            'Posterior_Coefficients', 'sample_posterior_combine', 'sample_from_model'}
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in keep]
    mod = types.ModuleType('ref_engine_test')
    mod.__dict__.update(dict(torch=torch, np=np, autocast=autocast_ctx or nullcontext))
    exec(compile(ast.Module(body=body, type_ignores=[]), 'engine/test.py', 'exec'), mod.__dict__)
    return mod


def reference_config(nf=64, size=256, **over):
    """README.md:85 / demo notebook cell 3."""
    d = dict(num_channels=1, num_channels_dae=nf, ch_mult=[1, 2, 4], num_res_blocks=2, attn_resolutions=[16], dropout=0.,
             resamp_with_conv=True, conditional=True, fir=True, fir_kernel=[1, 3, 3, 1], skip_rescale=True,
             resblock_type='biggan', progressive='none', progressive_input='residual', progressive_combine='sum',
             embedding_type='positional', fourier_scale=16., not_use_tanh=False, image_size=size, nz=100, z_emb_dim=256,
             t_emb_dim=256, n_mlp=3, centered=True, num_timesteps=4, beta_min=0.1, beta_max=20., use_geometric=False)
    d.update(over)
    return Namespace(**d)


def build_models(cfg, device, healthy=False, state_dicts=None):
    """The reference's NCSNpp / NCSNpp_adaptive on `device`, eval mode, with the oracle's deterministic non-degenerate
    weights unless `state_dicts` is given."""
    mod = import_reference(healthy)
    g1, g2 = mod.NCSNpp(cfg).eval(), mod.NCSNpp_adaptive(cfg).eval()
    if state_dicts is None:
        root = os.path.dirname(HERE)
        if root not in sys.path:
            sys.path.insert(0, root)
        from oracle import mudiff_oracle as O
        v = ('g1_healthy', 'g2_healthy') if healthy else ('g1', 'g2')
        state_dicts = (O.make_state_dict(cfg, v[0], seed=0), O.make_state_dict(cfg, v[1], seed=1))
    g1.load_state_dict(state_dicts[0], strict=True)
    g2.load_state_dict(state_dicts[1], strict=True)
    return g1.to(device), g2.to(device)


def run_loop(E, cfg, g1, g2, conds, x_init, latents=None, noises=None):
    """The reference's sample_from_model; with `latents` / `noises` its RNG draws are redirected to the pre-drawn
    tensors in the reference's own draw order (z of step i, then the posterior noise of step i)."""
    import torch
    dev = x_init.device
    pc = E.Posterior_Coefficients(cfg, dev)
    c3 = conds[2] if len(conds) > 2 else None
    if latents is None:
        return E.sample_from_model(pc, g1, conds[0], g2, conds[1], c3, cfg.num_timesteps, x_init, None, cfg)
    queue = []
    for i in reversed(range(cfg.num_timesteps)):
        queue += [latents[i], noises[i]]
    o_randn, o_like = torch.randn, torch.randn_like
    torch.randn = lambda *a, **k: queue.pop(0)
    torch.randn_like = lambda *a, **k: queue.pop(0)
    try:
        return E.sample_from_model(pc, g1, conds[0], g2, conds[1], c3, cfg.num_timesteps, x_init, None, cfg)
    finally:
        torch.randn, torch.randn_like = o_randn, o_like


def time_gpu(nf=64, size=256, batch=1, mode='fp16', iters=10, warmup=3, device='cuda:0', models=None):
    """Reference GPU baseline: the reference modules + its own loop + its own CUDA extensions, eager.
    mode: 'fp16' = torch.autocast(float16) exactly like engine/test.py:191, 'fp32' = no autocast, TF32 off,
    'tf32' = no autocast, TF32 on.  Returns dict(ms_best, ms_median, slices_per_s, ...)."""
    import statistics
    import torch
    dev = torch.device(device)
    cfg = reference_config(nf, size)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = (mode == 'tf32')
    try:
        g1, g2 = models if models is not None else build_models(cfg, dev)
        ctx = (lambda: torch.autocast('cuda', dtype=torch.float16)) if mode == 'fp16' else nullcontext
        E = engine_symbols(ctx)
        gen = torch.Generator(device=dev).manual_seed(42)
        conds = [torch.randn(batch, 1, size, size, device=dev, generator=gen).clamp(-3, 3) / 3 for _ in range(3)]
        x_init = torch.randn(batch, 1, size, size, device=dev, generator=gen)
        for _ in range(warmup):
            run_loop(E, cfg, g1, g2, conds, x_init)
        torch.cuda.synchronize(dev)
        times = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = run_loop(E, cfg, g1, g2, conds, x_init)
            b.record()
            torch.cuda.synchronize(dev)
            times.append(a.elapsed_time(b))
        ok = bool(torch.isfinite(out).all().item())
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    med = statistics.median(times)
    U = sys.modules['utils.op.upfirdn2d']         # (`utils.op.upfirdn2d` the attribute is the function, not the module)
    return {"mode": mode, "batch": batch, "size": size, "nf": nf, "iters": iters, "warmup": warmup,
            "ms_best": min(times), "ms_median": med, "slices_per_s": batch / (med * 1e-3),
            "slices_per_s_best": batch / (min(times) * 1e-3), "finite": ok,
            "cuda_extension": U.upfirdn2d_op is not None,
            "tf32": mode == 'tf32', "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}


def time_cpu(nf=64, size=256, batch=1, steps=1, warmup=0, threads=None, budget_s=None):
    """The reference's own CPU path (CPU tensors -> upfirdn2d_native, utils/op/upfirdn2d.py:201) through its own loop,
    fp32, on `threads` host threads (default: all cores).  Returns (seconds per step list, threads)."""
    import torch
    if threads is None:
        threads = os.cpu_count() or 1
    torch.set_num_threads(int(threads))
    cfg = reference_config(nf, size)
    g1, g2 = build_models(cfg, 'cpu')
    E = engine_symbols(nullcontext)
    gen = torch.Generator().manual_seed(42)
    conds = [torch.randn(batch, 1, size, size, generator=gen).clamp(-3, 3) / 3 for _ in range(3)]
    x_init = torch.randn(batch, 1, size, size, generator=gen)
    for _ in range(warmup):
        run_loop(E, cfg, g1, g2, conds, x_init)
    times = []
    start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        run_loop(E, cfg, g1, g2, conds, x_init)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - start > budget_s:
            break                                  # bounded sample: report the steps that were timed
    return times, torch.get_num_threads()

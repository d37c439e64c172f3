"""ctypes binding of libmudiff_b200.so (C ABI declared in include/mudiff_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `csrc/Makefile`.  There is no
CPU or PyTorch fallback: if the shared object is missing or a launch is rejected the
caller gets a RuntimeError (the reference raises RuntimeError("CUDA upfirdn2d extension
not available"), utils/op/upfirdn2d.py:114-115).
"""
import ctypes as C
import os

import torch

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_SIGMOID, ACT_TANH, ACT_LRELU = 0, 1, 2, 3, 4
EINVAL, EUNSUPPORTED = -22, -95

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmudiff_b200.so')

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class ConvDesc(C.Structure):
    """struct mudiff_conv_desc (include/mudiff_b200.h)."""
    _fields_ = [
        ('a', C.c_void_p * 3), ('a_c', C.c_int32 * 3), ('a_ld', C.c_int32 * 3), ('a_taps', C.c_int32 * 3),
        ('nseg', C.c_int32), ('a_batched', C.c_int32),
        ('batch', C.c_int32), ('h', C.c_int32), ('w', C.c_int32),
        ('stride', C.c_int32), ('pad', C.c_int32),
        ('wt', C.c_void_p), ('w_bstride', C.c_int64), ('w_ld', C.c_int32), ('n', C.c_int32),
        ('bias', C.c_void_p), ('rowbias', C.c_void_p), ('rowbias_ld', C.c_int32),
        ('residual', C.c_void_p), ('res_ld', C.c_int32),
        ('alpha', C.c_float), ('beta', C.c_float), ('act', C.c_int32),
        ('out', C.c_void_p), ('out_ld', C.c_int32), ('out_coff', C.c_int32), ('out_dtype', C.c_int32),
        ('stats', C.c_void_p), ('stats_groups', C.c_int32), ('flags', C.c_int32),
        ('a_xform', C.c_void_p * 3), ('a_xform_ld', C.c_int32 * 3), ('a_xform_act', C.c_int32),
    ]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes  (restype is int unless listed in _RESTYPES)
PROTOTYPES = {
    'mudiff_abi_version': [],
    'mudiff_build_info': [],
    'mudiff_launch_count': [],
    'mudiff_conv_desc_size': [],
    'mudiff_upfirdn2d': [_P, _P, _P, _I, _L, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'mudiff_upfirdn2d_gn': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'mudiff_split3_bf16': [_P, _I, _P, _L, _I, _I, _P],
    'mudiff_fused_bias_act': [_P, _P, _P, _P, _I, _L, _I, _L, _I, _I, _F, _F, _P],
    'mudiff_posterior_update': [_P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _I, _P, _I, _L, _P],
    'mudiff_gn_stats': [_P, _I, _I, _I, _I, _L, _P, _I, _I, _P],
    'mudiff_gn_fused': [_P, _P, _I, _I, _L, _I, _P, _P, _L, _F, _I, _P, _I, _I, _P],
    'mudiff_gn_stats_table': [_P, _I, _I, _I, _P, _I, _I, _P, _I, _P, _P, _L, _I, _L, _I, _F, _P, _P],
    'mudiff_gn_scale_shift': [_P, _I, _I, _P, _I, _I, _P, _P, _L, _I, _L, _I, _F, _P, _P],
    'mudiff_stats_finalize': [_P, _I, _I, _P, _I, _I, _I, _P],
    'mudiff_gn_stats_apply': [_P, _I, _I, _P, _I, _P, _I, _I, _P, _I, _I, _P, _P, _L, _P, _I, _I, _L, _I, _F, _I, _P],
    'mudiff_gn_apply': [_P, _I, _I, _P, _I, _P, _I, _I, _P, _I, _I, _P, _P, _L, _P, _I, _I, _I, _L, _I, _F, _I, _P],
    'mudiff_zero': [_P, _L, _P],
    'mudiff_conv_tc': [C.POINTER(ConvDesc), _P],
    'mudiff_conv_tc_query': [C.POINTER(ConvDesc), C.POINTER(C.c_int32)],
    'mudiff_attention_tc': [_P, _P, _P, _I, _I, _I, _F, _P],
    'mudiff_volume_workspace_bytes': [],
    'mudiff_volume_window': [_P, _L, _F, _F, _P, _P],
    'mudiff_volume_window_read': [_P, _P, _P],
    'mudiff_volume_to_slices': [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    'mudiff_slices_to_volume': [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    'mudiff_zscore_to_unit': [_P, _P, _L, _P],
    'mudiff_minmax_keys': [_P, _L, _I, _P, _P],
    'mudiff_minmax_read': [_P, _P, _P],
    'mudiff_scale_to_u8': [_P, _L, _P, _P, _P],
    'mudiff_minibatch_stddev': [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'mudiff_debug_last_timeout': [C.POINTER(C.c_int32)],
    'mudiff_debug_dump': [C.POINTER(C.c_int32), _I],
    'mudiff_set_wait_timeout': [C.c_longlong],
    'mudiff_debug_selftest': [],
    'mudiff_conv_simt': [C.POINTER(ConvDesc), _I, _P],
    'mudiff_stem_conv_tc': [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'mudiff_stem_moments': [_P, _I, _I, _I, _I, _P, _P],
    'mudiff_stem_conv_gn_act': [_P, _I, _P, _P, _P, _P, _P, _L, _I, _F, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'mudiff_softmax_rows': [_P, _P, _I, _L, _I, _F, _P],
    'mudiff_linear': [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    'mudiff_timestep_embedding': [_P, _P, _I, _I, _F, _P],
    'mudiff_pixelnorm': [_P, _P, _I, _I, _P],
    'mudiff_gate_mul': [_P, _I, _P, _I, _P, _I, _I, _L, _I, _P],
    'mudiff_gate_blend': [_P, _I, _P, _I, _P, _I, _P, _I, _I, _L, _I, _P],
    'mudiff_add_scale': [_P, _P, _P, _I, _L, _F, _P],
    'mudiff_copy_channels': [_P, _I, _I, _P, _I, _I, _L, _I, _P],
    'mudiff_gap': [_P, _I, _I, _P, _I, _L, _I, _P],
    'mudiff_tanh': [_P, _P, _I, _I, _L, _P],
}
_RESTYPES = {'mudiff_build_info': C.c_char_p, 'mudiff_launch_count': C.c_int64}

_lib = None
_CALL_PROFILER = None       # callable(name) -> context manager; used by bench.py for per-kernel timing


def set_call_profiler(fn):
    global _CALL_PROFILER
    _CALL_PROFILER = fn


class _Proxy:
    """Thin attribute proxy over the CDLL so that every entry point can be bracketed by CUDA events
    when a profiler is installed (zero extra work otherwise)."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name in PROTOTYPES:
            if hasattr(cdll, name):
                setattr(self, name, self._wrap(name, getattr(cdll, name)))
            else:                                   # only possible with MUDIFF_LIB (older debug builds)
                setattr(self, name, self._missing(name))

    @staticmethod
    def _missing(name):
        def call(*args):
            raise RuntimeError(f"mu-diff_b200: the library selected with MUDIFF_LIB does not export {name}")
        return call

    @staticmethod
    def _wrap(name, fn):
        sync_ops = os.environ.get('MUDIFF_SYNC', '')       # debug: "all" or comma-separated entry points
        do_sync = sync_ops == 'all' or name in sync_ops.split(',')
        counter = [0]

        def call(*args):
            prof = _CALL_PROFILER
            if prof is None:
                rc = fn(*args)
            else:
                with prof(name):
                    rc = fn(*args)
            if do_sync and name not in ('mudiff_debug_last_timeout', 'mudiff_debug_dump', 'mudiff_debug_selftest', 'mudiff_conv_tc_query'):
                counter[0] += 1
                try:
                    torch.cuda.synchronize()
                except Exception as e:
                    raise RuntimeError(f"mu-diff_b200: device error detected right after {name} call #{counter[0]}: "
                                       f"{str(e).splitlines()[0]}") from None
            return rc
        return call


def lib():
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"mu-diff_b200: {LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C mu-diff_b200/csrc`). There is no CPU fallback.")
        alt = os.environ.get('MUDIFF_LIB')          # debug only: bisecting older builds of the library
        l = C.CDLL(os.path.abspath(alt) if alt else LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            if alt and not hasattr(l, name):
                continue
            fn = getattr(l, name)           # AttributeError if the .so does not export a declared symbol
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        if not alt and l.mudiff_conv_desc_size() != C.sizeof(ConvDesc):
            raise RuntimeError('mu-diff_b200: ConvDesc layout mismatch between _lib.py and the shared library')
        _lib = _Proxy(l)
        w = os.environ.get('MUDIFF_WAIT_CYCLES')    # bound of the kernels' mbarrier waits; 0 = unbounded (ncu, MPS)
        if w is not None and torch.cuda.is_available() and hasattr(l, 'mudiff_set_wait_timeout'):
            set_wait_timeout(int(w))
    return _lib


def set_wait_timeout(cycles: int):
    """Bound (clock64 cycles, 0 = none) of the mbarrier waits in the tcgen05 kernels on the CURRENT device."""
    check(lib().mudiff_set_wait_timeout(int(cycles)), 'set_wait_timeout')


def dtype_code(t: torch.dtype) -> int:
    try:
        return _DTYPES[t]
    except KeyError:
        raise RuntimeError(f"mu-diff_b200: unsupported dtype {t}") from None


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc == EINVAL:
        raise RuntimeError(f"mu-diff_b200: {what}: invalid argument")
    if rc == EUNSUPPORTED:
        raise RuntimeError(f"mu-diff_b200: {what}: unsupported shape/dtype for this kernel")
    info = (C.c_int32 * 8)()
    try:
        lib().mudiff_debug_last_timeout(info)
    except Exception:
        pass
    extra = f" [mbarrier wait timed out: block {info[1]} warp {info[2]} lane {info[3]} bar@{info[4]:#x} parity {info[5]} grid {info[6]}]" if info[0] else ""
    raise RuntimeError(f"mu-diff_b200: {what}: CUDA error {rc}{extra}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mu-diff_b200: expected CUDA tensors (this package has no CPU path)")


def launch_count() -> int:
    return int(lib().mudiff_launch_count())

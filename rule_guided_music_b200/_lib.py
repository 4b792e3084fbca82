"""ctypes binding of librgm_b200.so (C ABI declared in include/rgm_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not sm_100, every call
raises.  The library is built in-tree by ``__graft_entry__.build()`` (``make -C rule_guided_music_b200/csrc``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RGM_LIB: an experiment build of the same library (csrc/Makefile EXTRA / OUT), for A/B measurements by the tools
LIB_PATH = os.environ.get("RGM_LIB") or os.path.join(_HERE, "librgm_b200.so")

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_double = ctypes.c_double
c_ll = ctypes.c_longlong


class RgmError(RuntimeError):
    pass


def _declare(lib):
    lib.rgm_last_error.restype = ctypes.c_char_p
    lib.rgm_last_error.argtypes = []
    lib.rgm_version.restype = c_int
    lib.rgm_launch_count.restype = ctypes.c_ulonglong
    lib.rgm_check_device.restype = c_int
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = args


# name -> argtypes; every function returns int (0 = ok)
_HP = ctypes.POINTER(ctypes.c_void_p)
c_char_p = ctypes.c_char_p
class RuleSpec(ctypes.Structure):
    """rgm_rule_spec of include/rgm_b200.h."""
    _fields_ = [("kind", c_int), ("interval", c_int), ("horizontal_scale", c_float), ("loss_kind", c_int),
                ("weight", c_float), ("target", c_void_p)]


COEF_ROWS = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
             "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
             "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2", "fixed_large_variance",
             "fixed_large_log_variance", "log_betas"]  # RGM_COEF_* order

_SIGNATURES = {
    "rgm_prof_enable": [c_int],
    "rgm_prof_summary": [c_char_p, c_int],
    "rgm_dit_create": [_HP, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int],
    "rgm_dit_destroy": [c_void_p],
    "rgm_dit_reserve": [c_void_p, c_int, c_int],
    "rgm_vae_reserve": [c_void_p, c_int],
    "rgm_coeff_tables": [ctypes.POINTER(c_double), c_int, c_void_p, c_void_p],
    "rgm_coeff_tables_host": [ctypes.POINTER(c_double), c_int, ctypes.POINTER(c_float)],
    "rgm_ddim_mean": [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p, c_int,
                      c_ll, c_void_p],
    "rgm_scg_create": [_HP, c_void_p, c_void_p],
    "rgm_scg_reserve": [c_void_p, c_int, c_int, c_int, c_int, c_int],
    "rgm_scg_destroy": [c_void_p],
    "rgm_scg_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                     ctypes.POINTER(RuleSpec), c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                     c_void_p],
    "rgm_dit_set_lanes": [c_void_p, c_int],
    "rgm_dit_load": [c_void_p, c_char_p, c_void_p, c_ll, c_void_p],
    "rgm_dit_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "rgm_vae_create": [_HP, c_int, ctypes.POINTER(c_int), c_int, c_int, c_int, c_int],
    "rgm_vae_destroy": [c_void_p],
    "rgm_vae_gn_timeouts": [c_void_p],
    "rgm_vae_set_lanes": [c_void_p, c_int],
    "rgm_vae_load": [c_void_p, c_char_p, c_void_p, c_ll, c_void_p],
    "rgm_vae_encode": [c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "rgm_vae_decode_latents": [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_void_p],
    "rgm_rule_pitch_hist": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "rgm_rule_note_density": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "rgm_rule_loss_accum": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "rgm_scg_fanout": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p],
    "rgm_x0_from_eps": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p],
    "rgm_scg_select": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p],
    "rgm_gemm_f16": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "rgm_conv_f16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                     c_int, c_void_p, c_void_p],
    "rgm_gn_apply_f16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "rgm_conv_gn_f16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                        c_void_p, c_void_p],
    "rgm_conv_norm_f16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                          c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "rgm_pack_conv_weight": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rgm_attention_f16": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p],
}


def lib():
    """Load the shared library once; raise loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RgmError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the B200 path has no CPU or PyTorch fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def exported_symbols():
    """Names every include/rgm_b200.h entry point must resolve to (used by the CPU-side ABI test)."""
    return ["rgm_last_error", "rgm_version", "rgm_launch_count", "rgm_check_device"] + list(_SIGNATURES)


def check(rc):
    if rc != 0:
        raise RgmError(lib().rgm_last_error().decode("utf-8", "replace"))


def call(name, *args):
    check(getattr(lib(), name)(*args))


def ptr(t):
    """Raw device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)


def prof_enable(on):
    call("rgm_prof_enable", 1 if on else 0)


def prof_summary():
    """Per kernel family: dict(name -> dict(launches, ms, flops_alg, flops_exec, bytes))."""
    buf = ctypes.create_string_buffer(1 << 20)
    call("rgm_prof_summary", buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms, fa, fe, by = line.split("\t")
        out[name] = dict(launches=int(n), ms=float(ms), flops_alg=float(fa), flops_exec=float(fe), bytes=float(by))
    return out


def launch_count():
    return int(lib().rgm_launch_count())

"""ctypes binding of the C ABI in include/mgf.h.  Loads the in-tree libmgf_sm100a.so and fails loudly if it is
missing -- there is no Python/CPU fallback for any op."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmgf_sm100a.so")

c_void_p, c_int, c_float, c_int64, c_int32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int64, ctypes.c_int32

F32, BF16, F16, F64 = 0, 1, 2, 3


class ConvShape(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("N", "IC", "H", "W", "OC", "HO", "WO", "KH", "KW", "stride_h", "stride_w",
                                       "pad_h", "pad_w", "dil_h", "dil_w", "groups")]


# name -> (restype, argtypes); the single list the "exports every symbol" test checks against include/mgf.h
SIGNATURES = {
    "mgf_last_error": (ctypes.c_char_p, []),
    "mgf_version": (c_int, []),
    "mgf_launch_count": (c_int64, []),
    "mgf_set_forward_dtype": (c_int, [c_int]),
    "mgf_get_forward_dtype": (c_int, []),
    "mgf_fp16_overflow_read": (c_int, [ctypes.POINTER(c_int), c_int, c_void_p]),
    "mgf_bias_act": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_float, c_float, c_float, c_int64, c_int64, c_int64, c_void_p]),
    "mgf_upfirdn2d": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "mgf_conv2d_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(ConvShape), c_void_p]),
    "mgf_conv2d_dgrad_f32": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(ConvShape), c_void_p]),
    "mgf_conv2d_wgrad_f32": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(ConvShape), c_void_p]),
    "mgf_style_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_style_bwd": (c_int, [c_void_p] * 6 + [c_float, c_float, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_modulate_weights": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_int64, c_void_p]),
    "mgf_scale_channels": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p]),
    "mgf_small_gemm": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_torgb_fwd": (c_int, [c_void_p] * 5 + [c_int, c_int64, c_int, c_void_p]),
    "mgf_torgb_bwd": (c_int, [c_void_p] * 7 + [c_int, c_int64, c_int, c_void_p]),
    "mgf_act_bwd": (c_int, [c_void_p] * 7 + [c_float, c_float, c_int, c_int, c_int64, c_int, c_int64, c_void_p]),
    "mgf_fir4": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int] + [c_int] * 8 + [c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int,
                         c_float, c_float, c_void_p]),
    "mgf_fir4_pad": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_upfir2_add": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_pointwise": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "mgf_pointwise_supported": (c_int, [c_int, c_int]),
    "mgf_upfir2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_fir_set_mode": (c_int, [c_int]),
    "mgf_fir_get_mode": (c_int, []),
    "mgf_attn_fwd": (c_int, [c_void_p] * 9 + [c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int64, c_void_p]),
    "mgf_attn_bwd": (c_int, [c_void_p] * 10 + [c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int64, c_void_p]),
    "mgf_attn_table_bytes": (c_int64, [c_int, c_int]),
    "mgf_attn_tables": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mgf_attn_set_split": (c_int, [c_int]),
    "mgf_lpips_prep": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mgf_mapping_param_floats": (c_int, []),
    "mgf_mapping_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mgf_mapping_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mgf_vgg_conv1_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mgf_vgg_conv1_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p]),
    "mgf_lpips_prep_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p]),
    "mgf_maxpool2_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_maxpool2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_lpips_head": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p]),
    "mgf_lpips_tap_pool_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_lpips_tap_pool_bwd": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_int, c_void_p]),
    "mgf_adam_noise_step": (c_int, [c_void_p] * 5 + [c_int, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_int64, c_void_p]),
    "mgf_fma": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
}

_lib = None

# modules that add their own entries (with ctypes struct types) to SIGNATURES when imported
_REGISTRARS = ["morphganformer_b200.tc"]


def register_all():
    import importlib
    for m in _REGISTRARS:
        importlib.import_module(m)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "morphganformer_b200: CUDA library %s is missing. Build it with `python -m morphganformer_b200.build` "
                "(or __graft_entry__.build()). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        register_all()
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class MgfError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib().mgf_last_error()
        raise MgfError("%s failed (status %d): %s" % (what or "mgf call", rc, msg.decode() if msg else ""))


def dtype_code(dt):
    import torch
    try:
        return {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16, torch.float64: F64}[dt]
    except KeyError:
        raise MgfError("unsupported dtype %s" % dt)


def stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return t.data_ptr() if t is not None else None


def require_cuda(t, who):
    if t.device.type != "cuda":
        raise MgfError("%s: tensor is on %s; this build has CUDA kernels only (no CPU fallback). "
                       "Tests compare against oracle/ on the CPU." % (who, t.device))


def i64x(*v):
    return (c_int64 * len(v))(*v)


def i32x(*v):
    return (c_int32 * len(v))(*v)


DEFAULT_FORWARD_DTYPE = "fp16"      # what a freshly loaded library is in (runtime.cu)


def set_forward_dtype(name):
    """'fp16' (default) or 'bf16': element type of the engine's forward activations / forward GEMM operands (gradients stay bf16).
    fp16 has the same tensor-core rate and 8x finer rounding: it is the mode that meets the parity bars (images within 1e-2 max-abs,
    per-step loss within 1e-3 of the fp32 reference, tests/test_fullsize_parity_gpu.py); stores saturate at +-65504 and raise a device
    flag (check_fp16_overflow).  bf16 has the fp32 exponent range and ~4e-2 image error: an explicit opt-in for checkpoints whose
    activations overflow fp16."""
    code = {"bf16": BF16, "fp16": F16}[name]
    check(lib().mgf_set_forward_dtype(code), "mgf_set_forward_dtype")


def forward_torch_dtype():
    import torch
    return torch.float16 if lib().mgf_get_forward_dtype() == F16 else torch.bfloat16


def fp16_overflow(device=None, reset=True):
    """True if an fp16-forward store saturated since the last reset (synchronises the current stream)."""
    v = c_int(0)
    import torch
    with torch.cuda.device(device):
        check(lib().mgf_fp16_overflow_read(ctypes.byref(v), int(reset), stream_ptr(device)), "mgf_fp16_overflow_read")
    return bool(v.value)


def check_fp16_overflow(device=None, who="tc engine"):
    if fp16_overflow(device):
        raise MgfError("%s: forward activations left the fp16 range (+-65504) and were clipped; results are not trustworthy. "
                       "Select bf16 forward storage for this checkpoint: _lib.set_forward_dtype('bf16') / Projector(forward_dtype='bf16')." % who)

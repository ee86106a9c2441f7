"""ctypes-level wrapper of mgf_conv_tc (include/mgf.h): the tcgen05 implicit-GEMM convolution used by the bf16 engine.
Tensors are NHWC bf16 (`[N, H, W, C]` torch tensors, channel stride 1)."""
import ctypes
import torch
from . import _lib

c_void_p, c_int32, c_int64, c_float, c_int8 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_int8


class TcAct(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("C", c_int64), ("W", c_int64), ("H", c_int64), ("N", c_int64),
                ("sW", c_int64), ("sH", c_int64), ("sN", c_int64)]


class TcTap(ctypes.Structure):
    _fields_ = [("amap", c_int8), ("dy", c_int8), ("dx", c_int8), ("_pad", c_int8), ("wz", c_int32)]


class ConvTcDesc(ctypes.Structure):
    _fields_ = [("a", TcAct * 4), ("n_a", c_int32),
                ("w", c_void_p), ("w_G", c_int64), ("w_T", c_int64), ("w_NT", c_int64), ("w_K", c_int64),
                ("taps", TcTap * 40), ("ntaps", c_int32),
                ("GW", c_int32), ("GH", c_int32), ("NB", c_int32),
                ("phases", c_int32), ("Cout", c_int32),
                ("out", c_void_p), ("OH", c_int64), ("OW", c_int64), ("OC", c_int64), ("osy", c_int32), ("osx", c_int32),
                ("ofy", c_int32 * 4), ("ofx", c_int32 * 4),
                ("scale_n", c_void_p), ("reduce_out", c_void_p), ("X", c_void_p),
                ("noise", c_void_p), ("noise_strength", c_void_p), ("bias", c_void_p),
                ("act", c_int32), ("alpha", c_float), ("gain", c_float),
                ("add", c_void_p),
                ("actgrad", c_int32), ("ag_alpha", c_float), ("ag_gain", c_float),
                ("bn", c_int32), ("reduce_per_sample", c_int32),
                ("ab_fwd", c_int32), ("out_fwd", c_int32), ("x_fwd", c_int32), ("add_fwd", c_int32), ("superpix", c_int32), ("noise_bstride", c_int64),
                ("phase_ntaps", c_int32 * 4)]


_lib.SIGNATURES["mgf_conv_tc"] = (ctypes.c_int, [ctypes.POINTER(ConvTcDesc), c_void_p])
_lib.SIGNATURES["mgf_conv_tc_set_halo"] = (ctypes.c_int, [ctypes.c_int])


def nhwc_view(t):
    """(ptr, C, W, H, N, sW, sH, sN) of a [N, H, W, C] bf16 tensor (any strides with channel stride 1)."""
    assert t.dtype in (torch.bfloat16, torch.float16) and t.ndim == 4 and t.stride(3) == 1, (t.dtype, t.shape, t.stride())
    n, h, w, c = t.shape
    return (t.data_ptr(), c, w, h, n, t.stride(2), t.stride(1), t.stride(0))


def phase_view(t, py, px):
    """strided view t[:, py::2, px::2, :] of an NHWC tensor as an activation descriptor."""
    return nhwc_view(t[:, py::2, px::2, :])


# When set to a list, every launch appends (tag, start_event, end_event, algorithmic_flops, executed_flops); used by bench.py to
# measure the dominant kernel live with CUDA events on the launching stream.
PROFILE = None


def conv_tc(acts, w, taps, grid, phases, cout, out, osy=1, osx=1, ofy=(0, 0, 0, 0), ofx=(0, 0, 0, 0), scale_n=None,
            reduce_out=None, X=None, noise=None, noise_strength=None, bias=None, act=0, alpha=0.2, gain=1.0, add=None,
            actgrad=False, ag_alpha=0.2, ag_gain=1.0, bn=0, reduce_per_sample=False, alg_scale=1.0, tag="", fwd=True, superpix=False, noise_bstride=0,
            phase_ntaps=None):
    # fwd=True: a forward launch -- operands, output and `add` are forward-dtype tensors (bf16 or fp16, _lib.set_forward_dtype);
    # fwd=False: a gradient launch -- operands/output/add are bf16 gradients; X (saved activation) is a forward tensor either way.
    """acts: list of NHWC bf16 tensors or descriptor tuples; w: [G, T, NT, K] bf16 contiguous; taps: [(amap, dy, dx, wz)];
    grid: (NB, GH, GW); out: [NB, OH, OW, OC] bf16 contiguous."""
    d = ConvTcDesc()
    d.n_a = len(acts)
    for i, a in enumerate(acts):
        v = a if isinstance(a, tuple) else nhwc_view(a)
        d.a[i] = TcAct(*v)
    h16 = (torch.bfloat16, torch.float16)
    want = _lib.forward_torch_dtype() if fwd else torch.bfloat16
    assert w.dtype == want and w.is_contiguous() and w.ndim == 4, (w.dtype, want)
    d.w, d.w_G, d.w_T, d.w_NT, d.w_K = w.data_ptr(), w.shape[0], w.shape[1], w.shape[2], w.shape[3]
    d.ntaps = len(taps)
    for i, (am, dy, dx, wz) in enumerate(taps):
        d.taps[i] = TcTap(am, dy, dx, 0, wz)
    d.NB, d.GH, d.GW = grid
    d.phases, d.Cout = phases, cout
    assert out.dtype == want and out.is_contiguous() and out.ndim == 4, (out.dtype, want)
    d.ab_fwd = d.out_fwd = d.add_fwd = int(bool(fwd))
    d.x_fwd = 1
    d.out, d.OH, d.OW, d.OC = out.data_ptr(), out.shape[1], out.shape[2], out.shape[3]
    d.osy, d.osx = osy, osx
    for i in range(4):
        d.ofy[i], d.ofx[i] = ofy[i], ofx[i]
    for name, t, dt in (("scale_n", scale_n, torch.float32), ("reduce_out", reduce_out, torch.float32), ("X", X, _lib.forward_torch_dtype()),
                        ("noise", noise, torch.float32), ("noise_strength", noise_strength, torch.float32),
                        ("bias", bias, torch.float32), ("add", add, want)):
        if t is not None:
            assert t.dtype == dt and t.is_contiguous(), name
            setattr(d, name, t.data_ptr())
    d.act, d.alpha, d.gain = act, alpha, gain
    d.actgrad, d.ag_alpha, d.ag_gain = int(bool(actgrad)), ag_alpha, ag_gain
    d.bn, d.reduce_per_sample = bn, int(bool(reduce_per_sample))
    d.superpix = int(bool(superpix))
    d.noise_bstride = int(noise_bstride)
    if phase_ntaps is not None:          # per-phase tap lists: `taps` is their concatenation, all phases share the [T][Cout][K] weights
        assert len(phase_ntaps) == phases and sum(phase_ntaps) == len(taps)
        for i, n in enumerate(phase_ntaps):
            d.phase_ntaps[i] = n
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().mgf_conv_tc(ctypes.byref(d), _lib.stream_ptr(out.device)), "mgf_conv_tc")
    if prof is not None:
        e1.record()
        ex = 2.0 * d.NB * d.GH * d.GW * (cout if phase_ntaps is not None else phases * cout) * w.shape[3] * len(taps)
        prof.append((tag, e0, e1, ex * alg_scale, ex, "B%d %dx%d C%d->%d x%d taps%d %s" % (d.NB, d.GH, d.GW, w.shape[3], cout, phases, len(taps),
                     "+".join(k for k, v in (("red", reduce_out), ("X", X), ("add", add), ("noise", noise), ("bias", bias)) if v is not None))))
    return out


TAPS_3X3 = [(0, ky - 1, kx - 1, ky * 3 + kx) for ky in range(3) for kx in range(3)]


def pack_w3x3(w):
    """[Cout, Cin, 3, 3] (correlation weights, as F.conv2d) -> [1, 9, Cout, Cin] bf16."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(1, kh * kw, co, ci).to(_lib.forward_torch_dtype()).contiguous()

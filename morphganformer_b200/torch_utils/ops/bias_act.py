"""bias_act: fused bias + activation + gain + clamp.  Mirrors reference torch_utils/ops/bias_act.py (public
names: activation_funcs, bias_act) with the body routed to mgf_bias_act (include/mgf.h).  First- and
second-order gradients go through the same kernel with grad=1/2 (reference bias_act.py:121-202)."""
import math
import torch
from ... import _lib


class _Spec(dict):
    __getattr__ = dict.__getitem__


def _spec(def_alpha, def_gain, cuda_idx, ref, has_2nd_grad):
    return _Spec(def_alpha=def_alpha, def_gain=def_gain, cuda_idx=cuda_idx, ref=ref, has_2nd_grad=has_2nd_grad)


# same table as reference bias_act.py:15-25 (the `func` entry is the oracle's business, not the product's)
activation_funcs = {
    "linear":   _spec(0,   1,            1, "",  False),
    "relu":     _spec(0,   math.sqrt(2), 2, "y", False),
    "lrelu":    _spec(0.2, math.sqrt(2), 3, "y", False),
    "tanh":     _spec(0,   1,            4, "y", True),
    "sigmoid":  _spec(0,   1,            5, "y", True),
    "elu":      _spec(0,   1,            6, "y", True),
    "selu":     _spec(0,   1,            7, "y", True),
    "softplus": _spec(0,   1,            8, "y", True),
    "swish":    _spec(0,   math.sqrt(2), 9, "x", True),
}


def _launch(x, b, xref, yref, dy, grad, dim, idx, alpha, gain, clamp):
    """One kernel call.  x dense (contiguous or channels_last); b 1-D or None."""
    _lib.require_cuda(x, "bias_act")
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    size_b, step_b = 1, 1
    if b is not None:
        size_b = b.numel()
        step_b = x.stride(dim)
    with torch.cuda.device(x.device):
        rc = _lib.lib().mgf_bias_act(_lib.ptr(x), _lib.ptr(b), _lib.ptr(xref), _lib.ptr(yref), _lib.ptr(dy), _lib.ptr(y),
                                     _lib.dtype_code(x.dtype), grad, idx, alpha, gain, clamp,
                                     x.numel(), size_b, step_b, _lib.stream_ptr(x.device))
    _lib.check(rc, "mgf_bias_act")
    return y


def _dense(t, fmt):
    return t.contiguous(memory_format=fmt)


def _fmt_of(x):
    return torch.channels_last if x.ndim == 4 and x.stride(1) == 1 and x.shape[1] > 1 else torch.contiguous_format


_cache = {}


def _make(dim, act, alpha, gain, clamp):
    key = (dim, act, alpha, gain, clamp)
    if key in _cache:
        return _cache[key]
    spec = activation_funcs[act]
    idx = spec.cuda_idx
    keep_x = ("x" in spec.ref) or spec.has_2nd_grad
    keep_y = ("y" in spec.ref) or (clamp >= 0 and "x" not in spec.ref)   # clamp masks the gradient by the saved output
    trivial = act == "linear" and gain == 1 and clamp < 0

    class BiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            ctx.fmt = _fmt_of(x)
            x = _dense(x, ctx.fmt)
            b = b.contiguous() if b is not None else None
            y = x if (trivial and b is None) else _launch(x, b, None, None, None, 0, dim, idx, alpha, gain, clamp)
            ctx.save_for_backward(x if keep_x else None, b if keep_x else None, y if keep_y else None)
            return y

        @staticmethod
        def backward(ctx, dy):
            dy = _dense(dy, ctx.fmt)
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = dy if trivial else BiasActGrad.apply(dy, x, b, y)
            if ctx.needs_input_grad[1]:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    class BiasActGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            ctx.fmt = _fmt_of(dy)
            dx = _launch(dy, b, x, y, None, 1, dim, idx, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            d_dx = _dense(d_dx, ctx.fmt)
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _launch(d_dx, b, x, y, dy, 2, dim, idx, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    _cache[key] = BiasAct
    return BiasAct


def bias_act(x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None, impl="cuda"):
    """Same contract as reference bias_act.py:47-81.  impl='ref' is not available in the product (the pure-PyTorch
    restatement lives in oracle/ops.py for tests); CPU tensors are rejected: there is no fallback."""
    assert isinstance(x, torch.Tensor)
    assert impl in ["ref", "cuda"]
    if impl == "ref":
        raise NotImplementedError("bias_act(impl='ref'): the reference implementation is test infrastructure (oracle/ops.py)")
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim
        assert b.shape[0] == x.shape[dim]
        assert b.dtype == x.dtype, "bias must have the dtype of x"
    return _make(dim, act, alpha, gain, clamp).apply(x, b)

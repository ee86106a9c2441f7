"""fma(a, b, c) = a * b + c with broadcast-reducing gradients.  Mirrors reference torch_utils/ops/fma.py:7-50.
The two broadcast patterns the hot path uses (networks.py:322: b = [N,C,1,1] demodulation, c = noise [H,W] /
[N,1,H,W] / full) run the mgf_fma kernel; any other broadcast is expanded first."""
import torch
from ... import _lib


def fma(a, b, c):  # => a * b + c
    return _FusedMultiplyAdd.apply(a, b, c)


def _launch(a, b, c):
    _lib.require_cuda(a, "fma")
    shape = torch.broadcast_shapes(a.shape, b.shape, c.shape)
    a = a.expand(shape).contiguous()
    n, ch = (shape[0], shape[1]) if len(shape) >= 2 else (1, 1)
    hw = a.numel() // max(n * ch, 1)
    bmode, cmode = 0, 1
    if len(shape) == 4 and tuple(b.shape) == (shape[0], shape[1], 1, 1):
        b = b.contiguous(); bmode = 1
    else:
        b = b.expand(shape).contiguous()
    if len(shape) == 4 and c.numel() == hw and tuple(c.shape[-2:]) == tuple(shape[-2:]):
        c = c.contiguous(); cmode = 2
    else:
        c = c.expand(shape).contiguous()
    b, c = b.to(a.dtype), c.to(a.dtype)
    out = torch.empty_like(a)
    if out.numel():
        with torch.cuda.device(a.device):
            _lib.check(_lib.lib().mgf_fma(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(out), _lib.dtype_code(a.dtype),
                                          n, ch, hw, bmode, cmode, _lib.stream_ptr(a.device)), "mgf_fma")
    return out


class _FusedMultiplyAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        out = _launch(a, b, c)
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        da = db = dc = None
        if ctx.needs_input_grad[0]:
            da = _unbroadcast(dout * b, a.shape)
        if ctx.needs_input_grad[1]:
            db = _unbroadcast(dout * a, b.shape)
        if ctx.needs_input_grad[2]:
            dc = _unbroadcast(dout, ctx.c_shape)
        return da, db, dc


def _unbroadcast(x, shape):
    extra = x.ndim - len(shape)
    assert extra >= 0
    dims = [i for i in range(x.ndim) if x.shape[i] > 1 and (i < extra or shape[i - extra] == 1)]
    if dims:
        x = x.sum(dim=dims, keepdim=True)
    if extra:
        x = x.reshape(-1, *x.shape[extra + 1:])
    assert x.shape == shape
    return x

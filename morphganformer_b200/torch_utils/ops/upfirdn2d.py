"""upfirdn2d: pad, zero-insert upsample, FIR filter, decimate.  Mirrors reference torch_utils/ops/upfirdn2d.py
(setup_filter, upfirdn2d, filter2d, upsample2d, downsample2d, _parse_padding, _get_filter_size) with the body
routed to mgf_upfirdn2d.  Backward = the same op with up/down swapped and the filter flipped (:238-256)."""
import numpy as np
import torch
from ... import _lib
from .. import misc


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and all(isinstance(x, int) for x in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(x, int) for x in padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    fw, fh = int(f.shape[-1]), int(f.shape[0])
    misc.assert_shape(f, [fh, fw][:f.ndim])
    assert fw >= 1 and fh >= 1
    return fw, fh


def setup_filter(f, device=torch.device("cpu"), normalize=True, flip_filter=False, gain=1, separable=None):
    """Same contract as reference upfirdn2d.py:64-108."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2] and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _launch(x, f2d, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    _lib.require_cuda(x, "upfirdn2d")
    assert x.ndim == 4
    n, c, h, w = x.shape
    fh, fw = f2d.shape
    ow = (w * upx + padx0 + padx1 - fw + downx) // downx
    oh = (h * upy + pady0 + pady1 - fh + downy) // downy
    if ow < 1 or oh < 1:
        raise _lib.MgfError("upfirdn2d: output size must be >= 1 (got %d x %d)" % (ow, oh))
    fmt = torch.channels_last if (x.stride(1) == 1 and c > 1) else torch.contiguous_format
    x = x.contiguous(memory_format=fmt)
    y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device, memory_format=fmt)
    if y.numel() == 0:
        return y
    f2d = f2d.to(device=x.device, dtype=torch.float32)
    xs, ys, fs = x.stride(), y.stride(), f2d.stride()
    with torch.cuda.device(x.device):
        rc = _lib.lib().mgf_upfirdn2d(
            _lib.ptr(x), _lib.ptr(f2d), _lib.ptr(y), _lib.dtype_code(x.dtype),
            _lib.i64x(w, h, c, n), _lib.i64x(xs[3], xs[2], xs[1], xs[0]),
            _lib.i32x(fw, fh), _lib.i64x(fs[1], fs[0]),
            _lib.i64x(ow, oh, c, n), _lib.i64x(ys[3], ys[2], ys[1], ys[0]),
            _lib.i32x(upx, upy), _lib.i32x(downx, downy), _lib.i32x(padx0, pady0),
            int(bool(flip)), float(gain), _lib.stream_ptr(x.device))
    _lib.check(rc, "mgf_upfirdn2d")
    return y


_cache = {}


def _make(up, down, padding, flip_filter, gain):
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    key = (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    if key in _cache:
        return _cache[key]

    class Upfirdn2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            if f is None:
                f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
            assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
            if f.ndim == 2:
                y = _launch(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
            else:  # separable: one pass per axis (reference :227-230)
                y = _launch(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, np.sqrt(gain))
                y = _launch(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, np.sqrt(gain))
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            f, = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _get_filter_size(f)
            p = [fw - padx0 - 1, iw * upx - ow * downx + padx0 - upx + 1,
                 fh - pady0 - 1, ih * upy - oh * downy + pady0 - upy + 1]
            dx = None
            if ctx.needs_input_grad[0]:
                dx = _make(up=[downx, downy], down=[upx, upy], padding=p, flip_filter=(not flip_filter), gain=gain).apply(dy, f)
            assert not ctx.needs_input_grad[1]
            return dx, None

    _cache[key] = Upfirdn2d
    return Upfirdn2d


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """Same contract as reference upfirdn2d.py:112-156; CUDA only (no fallback)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ["ref", "cuda"]
    if impl == "ref":
        raise NotImplementedError("upfirdn2d(impl='ref'): the reference implementation is test infrastructure (oracle/ops.py)")
    return _make(up, down, padding, flip_filter, gain).apply(x, f)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl="cuda"):
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2, pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)

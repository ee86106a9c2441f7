"""conv2d_gradfix: conv2d / conv_transpose2d with arbitrarily high order gradients.  Mirrors the public surface of
reference torch_utils/ops/conv2d_gradfix.py (conv2d :27, conv_transpose2d :32, flags `enabled`,
`weight_gradients_disabled`, context manager no_weight_gradients :17-23).  The reference forwards to ATen/cuDNN;
here fp32 tensors run the exact-fp32 direct-convolution kernels of libmgf_sm100a.so (mgf_conv2d_{fwd,dgrad,wgrad}_f32)
and every gradient is again one of those three kernels, so double backward works like the reference's custom op
(:96-157).  CUDA only: CPU tensors raise."""
import contextlib
import ctypes
import torch
from ... import _lib

enabled = True                      # kept for API compatibility; the custom op is always used
weight_gradients_disabled = False   # forcefully disable computation of gradients with respect to the weights


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def _pair(v):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    assert len(v) == 2 and all(isinstance(i, int) for i in v)
    return v


def _shape(n, ic, h, w, oc, ho, wo, kh, kw, stride, padding, dilation, groups):
    return _lib.ConvShape(n, ic, h, w, oc, ho, wo, kh, kw, stride[0], stride[1], padding[0], padding[1],
                          dilation[0], dilation[1], groups)


def _prep(t):
    _lib.require_cuda(t, "conv2d_gradfix")
    if t.dtype != torch.float32:
        raise _lib.MgfError("conv2d_gradfix: only float32 runs on the exact direct-convolution kernels "
                            "(bf16 goes through the tcgen05 engine in morphganformer_b200.engine); got %s" % t.dtype)
    return t.contiguous()


def _k_fwd(x, w, b, stride, padding, dilation, groups):
    x, w = _prep(x), _prep(w)
    n, ic, h, wd = x.shape
    oc, icg, kh, kw = w.shape
    assert icg * groups == ic, "weight/input channel mismatch"
    ho = (h + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    wo = (wd + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    y = torch.empty([n, oc, ho, wo], dtype=x.dtype, device=x.device)
    s = _shape(n, ic, h, wd, oc, ho, wo, kh, kw, stride, padding, dilation, groups)
    b = _prep(b) if b is not None else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mgf_conv2d_fwd_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), ctypes.byref(s),
                                                 _lib.stream_ptr(x.device)), "mgf_conv2d_fwd_f32")
    return y


def _k_dgrad(dy, w, in_hw, stride, padding, dilation, groups):
    """dx[N, IC, H, W] from dy[N, OC, HO, WO] and w[OC, IC/g, KH, KW]  (== conv_transpose2d forward)."""
    dy, w = _prep(dy), _prep(w)
    n, oc, ho, wo = dy.shape
    oc_w, icg, kh, kw = w.shape
    assert oc_w == oc
    ic = icg * groups
    h, wd = in_hw
    dx = torch.empty([n, ic, h, wd], dtype=dy.dtype, device=dy.device)
    s = _shape(n, ic, h, wd, oc, ho, wo, kh, kw, stride, padding, dilation, groups)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.lib().mgf_conv2d_dgrad_f32(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(dx), ctypes.byref(s),
                                                   _lib.stream_ptr(dy.device)), "mgf_conv2d_dgrad_f32")
    return dx


def _k_wgrad(dy, x, w_shape, stride, padding, dilation, groups):
    dy, x = _prep(dy), _prep(x)
    n, oc, ho, wo = dy.shape
    _, ic, h, wd = x.shape
    oc_w, icg, kh, kw = w_shape
    dw = torch.zeros(list(w_shape), dtype=x.dtype, device=x.device)
    s = _shape(n, ic, h, wd, oc, ho, wo, kh, kw, stride, padding, dilation, groups)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mgf_conv2d_wgrad_f32(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), ctypes.byref(s),
                                                   _lib.stream_ptr(x.device)), "mgf_conv2d_wgrad_f32")
    return dw


_cache = {}


def _op(transpose, weight_shape, stride, padding, output_padding, dilation, groups):
    """Returns the autograd Function for one (direction, geometry).  For transpose=False it maps
    x[N,IC,H,W], w[OC,IC/g,kh,kw] -> y; for transpose=True it maps x[N,OC,HO,WO], w[OC,IC/g,kh,kw] -> y[N,IC,H,W]."""
    weight_shape = tuple(int(v) for v in weight_shape)
    key = (transpose, weight_shape, stride, padding, output_padding, dilation, groups)
    if key in _cache:
        return _cache[key]
    kh, kw = weight_shape[2:]
    assert groups >= 1 and all(s >= 1 for s in stride) and all(p >= 0 for p in padding) and all(d >= 1 for d in dilation)
    if transpose:
        assert all(0 <= output_padding[i] < max(stride[i], dilation[i]) for i in range(2))
    else:
        assert all(op == 0 for op in output_padding)

    def out_hw_of_transpose(in_hw):
        return tuple((in_hw[i] - 1) * stride[i] - 2 * padding[i] + dilation[i] * (weight_shape[2 + i] - 1) + output_padding[i] + 1
                     for i in range(2))

    def calc_output_padding(input_hw, output_hw):
        # for the gradient of a non-transposed conv: the transposed conv must reproduce the input size
        if transpose:
            return (0, 0)
        return tuple(input_hw[i] - (output_hw[i] - 1) * stride[i] - (1 - 2 * padding[i]) - dilation[i] * (weight_shape[2 + i] - 1)
                     for i in range(2))

    class Conv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias):
            assert tuple(weight.shape) == weight_shape
            if not transpose:
                out = _k_fwd(input, weight, bias, stride, padding, dilation, groups)
            else:
                out = _k_dgrad(input, weight, out_hw_of_transpose(input.shape[2:]), stride, padding, dilation, groups)
                if bias is not None:
                    out = out + bias.reshape(1, -1, 1, 1)
            ctx.save_for_backward(input, weight)
            ctx.has_bias = bias is not None
            return out

        @staticmethod
        def backward(ctx, grad_output):
            input, weight = ctx.saved_tensors
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                p = calc_output_padding(input.shape[2:], grad_output.shape[2:])
                gi = _op(not transpose, weight_shape, stride, padding, p, dilation, groups).apply(grad_output, weight, None)
                assert gi.shape == input.shape
            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                gw = Conv2dGradWeight.apply(grad_output, input)
                assert tuple(gw.shape) == weight_shape
            if ctx.needs_input_grad[2]:
                gb = grad_output.sum([0, 2, 3])
            return gi, gw, gb

    class Conv2dGradWeight(torch.autograd.Function):
        @staticmethod
        def forward(ctx, grad_output, input):
            if not transpose:
                gw = _k_wgrad(grad_output, input, weight_shape, stride, padding, dilation, groups)
            else:  # roles swap: the transposed conv's input plays dy, its output gradient plays x
                gw = _k_wgrad(input, grad_output, weight_shape, stride, padding, dilation, groups)
            ctx.save_for_backward(grad_output, input)
            return gw

        @staticmethod
        def backward(ctx, grad2_grad_weight):
            grad_output, input = ctx.saved_tensors
            g2_go = g2_in = None
            if ctx.needs_input_grad[0]:
                g2_go = Conv2d.apply(input, grad2_grad_weight, None)
                assert g2_go.shape == grad_output.shape
            if ctx.needs_input_grad[1]:
                p = calc_output_padding(input.shape[2:], grad_output.shape[2:])
                g2_in = _op(not transpose, weight_shape, stride, padding, p, dilation, groups).apply(grad_output, grad2_grad_weight, None)
                assert g2_in.shape == input.shape
            return g2_go, g2_in

    _cache[key] = Conv2d
    return Conv2d


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """Same contract as reference conv2d_gradfix.py:27-30 (== torch.nn.functional.conv2d)."""
    return _op(False, weight.shape, _pair(stride), _pair(padding), (0, 0), _pair(dilation), groups).apply(input, weight, bias)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """Same contract as reference conv2d_gradfix.py:32-35 (== torch.nn.functional.conv_transpose2d);
    weight is [in_channels, out_channels/groups, kh, kw]."""
    return _op(True, weight.shape, _pair(stride), _pair(padding), _pair(output_padding), _pair(dilation), groups).apply(input, weight, bias)

"""conv2d_resample: 2-D convolution with optional up/down-sampling, on top of this package's conv2d_gradfix and
upfirdn2d.  Keeps the reference's public contract (torch_utils/ops/conv2d_resample.py:50-146): padding is given with
respect to the upsampled image and applied once; `flip_weight=True` means correlation (what conv2d does);
results equal the reference's six execution branches (1x1+down, 1x1+up, strided, transposed+FIR, plain, generic)."""
import torch
from .. import misc
from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _parse_padding
from .upfirdn2d import _get_filter_size


def _get_weight_shape(w):
    shape = [int(sz) for sz in w.shape]
    misc.assert_shape(w, shape)
    return shape


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    w = w if flip_weight else w.flip([2, 3])
    if transpose:
        return conv2d_gradfix.conv_transpose2d(x, w, stride=stride, padding=padding, groups=groups)
    return conv2d_gradfix.conv2d(x, w, stride=stride, padding=padding, groups=groups)


def _resample_pads(pads, fw, fh, up, down):
    """[x0, x1, y0, y1] after adding what the up/down FIR needs (reference :87-96)."""
    p = list(pads)
    for axis, ft in ((0, fw), (2, fh)):
        if up > 1:
            p[axis] += (ft + up - 1) // 2
            p[axis + 1] += (ft - up) // 2
        if down > 1:
            p[axis] += (ft - down + 1) // 2
            p[axis + 1] += (ft - down) // 2
    return p


def _swap_group_io(w, groups):
    """[OC, IC/g, kh, kw] -> conv_transpose2d layout [IC, OC/g, kh, kw] (reference :118-123)."""
    oc, icg, kh, kw = w.shape
    if groups == 1:
        return w.transpose(0, 1)
    return w.reshape(groups, oc // groups, icg, kh, kw).transpose(1, 2).reshape(groups * icg, oc // groups, kh, kw)


@misc.profiled_function
def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    assert isinstance(groups, int) and groups >= 1
    _, _, kh, kw = _get_weight_shape(w)
    fw, fh = _get_filter_size(f)
    pads = _resample_pads(_parse_padding(padding), fw, fh, up, down)
    fir = dict(f=f, flip_filter=flip_filter)
    conv = dict(groups=groups, flip_weight=flip_weight)
    pointwise = kh == 1 and kw == 1

    if pointwise and down > 1 and up == 1:      # filter+decimate first, then the cheap 1x1
        return _conv2d_wrapper(upfirdn2d.upfirdn2d(x, down=down, padding=pads, **fir), w, **conv)
    if pointwise and up > 1 and down == 1:      # 1x1 at low resolution, then interpolate
        return upfirdn2d.upfirdn2d(_conv2d_wrapper(x, w, **conv), up=up, padding=pads, gain=up ** 2, **fir)
    if down > 1 and up == 1:                    # low-pass, then strided convolution
        return _conv2d_wrapper(upfirdn2d.upfirdn2d(x, padding=pads, **fir), w, stride=down, **conv)
    if up > 1:                                  # transposed strided convolution, then low-pass (and optional decimate)
        px0, px1, py0, py1 = pads
        px0, px1, py0, py1 = px0 - (kw - 1), px1 - (kw - up), py0 - (kh - 1), py1 - (kh - up)
        pxt, pyt = max(min(-px0, -px1), 0), max(min(-py0, -py1), 0)
        x = _conv2d_wrapper(x, _swap_group_io(w, groups), stride=up, padding=[pyt, pxt], groups=groups, transpose=True,
                            flip_weight=not flip_weight)
        x = upfirdn2d.upfirdn2d(x, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, **fir)
        return upfirdn2d.upfirdn2d(x, down=down, **fir) if down > 1 else x
    px0, px1, py0, py1 = pads
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:   # plain convolution with symmetric padding
        return _conv2d_wrapper(x, w, padding=[py0, px0], **conv)
    # asymmetric or negative padding: pad/crop with an identity FIR, then convolve
    x = upfirdn2d.upfirdn2d(x, f=None, padding=pads, flip_filter=flip_filter)
    return _conv2d_wrapper(x, w, **conv)

"""GPU parity of the FIR kernels (fir.cu) and of the per-phase-tap transposed convolution against the oracle ops on the CPU
(oracle/ops.py restates the reference's upfirdn2d / conv2d_resample; reference upfirdn2d.py:176-199, conv2d_resample.py:117-134).
Inputs are rounded to the 16-bit storage type first, accumulation is fp32 on both sides, so the only difference is the final 16-bit
store: tolerance = 1 ulp of the storage type at the result's magnitude (+ fp32 summation order)."""
import ctypes
import math
import numpy as np
import pytest
import torch
import torch.nn.functional as F
from oracle import ops as O
import util

pytestmark = pytest.mark.gpu
FK = (ctypes.c_float * 4)(0.125, 0.375, 0.375, 0.125)


@pytest.fixture(autouse=True, params=[0, 7], ids=["vec8", "vec4"])
def fir_mode(request):
    """Every test of this file runs with both forms of the FIR walkers: 8 channels per thread and 4 channels per thread
    (`mgf_fir_set_mode` bits 0-2: fir4 / upfir2_add / upfir2_bwd); the library default is restored afterwards."""
    from morphganformer_b200 import _lib
    L = _lib.lib()
    old = L.mgf_fir_get_mode()
    L.mgf_fir_set_mode(request.param)
    yield request.param
    L.mgf_fir_set_mode(old)


def _setup(fwd):
    from morphganformer_b200 import _lib
    _lib.set_forward_dtype(fwd)
    dt = torch.float16 if fwd == "fp16" else torch.bfloat16
    return _lib, _lib.lib(), dt, (2.0 ** -10 if fwd == "fp16" else 2.0 ** -7)


def _nhwc(t, dt):
    return t.permute(0, 2, 3, 1).contiguous().to(dt).cuda()


def _nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2)


@pytest.mark.parametrize("fwd", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 32, 5, 7), (1, 64, 33, 16), (3, 8, 2, 2), (1, 512, 4, 4)])
@pytest.mark.parametrize("tail", [False, True])
def test_fir4_second_stage_of_upconv(fwd, shape, tail):
    """ct [B, 2h+2, 2w+2, C] (transposed-conv output, last row / column zero) -> z [B, 2h, 2w, C] = upfirdn2d(ct, f, pad 1, gain 4)
    (+ noise * strength + bias -> lrelu * gain when `tail`)."""
    _lib, L, dt, ulp = _setup(fwd)
    B, C, h, w = shape
    ct = util.case_tensor((B, C, 2 * h + 1, 2 * w + 1), 1).to(dt).float()
    f = O.setup_filter([1, 3, 3, 1])
    ref = O.upfirdn2d(ct, f, padding=[1, 1, 1, 1], gain=4)
    noise = util.case_tensor((2 * h, 2 * w), 2)
    nstr = torch.tensor([0.3])
    bias = util.case_tensor((C,), 3) * 0.2
    if tail:
        ref = O.bias_act(ref + noise * nstr, bias, act="lrelu", gain=math.sqrt(2) * 0.7)
    ctq = _nhwc(F.pad(ct, (0, 1, 0, 1)), dt)
    out = torch.full((B, 2 * h, 2 * w, C), float("nan"), dtype=dt, device="cuda")
    n_d, s_d, b_d = noise.cuda(), nstr.cuda(), bias.cuda()
    _lib.check(L.mgf_fir4(ctq.data_ptr(), out.data_ptr(), FK, 4.0, -1, B, 2 * h + 2, 2 * w + 2, 2 * h, 2 * w, 2 * h, 2 * w, C, 1, 1,
                          n_d.data_ptr() if tail else None, s_d.data_ptr() if tail else None, 0, b_d.data_ptr() if tail else None,
                          1 if tail else 0, 0.2, math.sqrt(2) * 0.7 if tail else 1.0, _lib.stream_ptr()), "mgf_fir4")
    got = _nchw(out)
    assert torch.isfinite(got).all()
    assert ((got - ref).abs() <= ulp * ref.abs() + 1e-5).all(), (got - ref).abs().max()


@pytest.mark.parametrize("shape", [(2, 32, 10, 14), (1, 64, 66, 32), (2, 8, 4, 4)])
def test_fir4_pad_adjoint_of_the_second_stage(shape):
    """g [B, H+2, W+2, C] = adjoint of upfirdn2d(pad 1, gain 4) applied to dy [B, H, W, C] on the (H+1) x (W+1) grid, padding row / column zero:
    checked against autograd through the oracle's upfirdn2d."""
    _lib, L, dt, ulp = _setup("fp16")
    B, C, H, W = shape
    dy = util.case_tensor((B, C, H, W), 4).to(torch.bfloat16).float()
    ct = torch.zeros(B, C, H + 1, W + 1, requires_grad=True)
    y = O.upfirdn2d(ct, O.setup_filter([1, 3, 3, 1]), padding=[1, 1, 1, 1], gain=4)
    ref, = torch.autograd.grad(y, [ct], dy)
    g = torch.full((B, H + 2, W + 2, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    dyq = _nhwc(dy, torch.bfloat16)
    _lib.check(L.mgf_fir4_pad(dyq.data_ptr(), g.data_ptr(), FK, 4.0, B, H, W, C, _lib.stream_ptr()), "mgf_fir4_pad")
    got = _nchw(g)
    assert (got[:, :, H + 1] == 0).all() and (got[:, :, :, W + 1] == 0).all()
    assert ((got[:, :, :H + 1, :W + 1] - ref).abs() <= 2.0 ** -7 * ref.abs() + 1e-5).all()


@pytest.mark.parametrize("fwd", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 32, 5, 7), (1, 64, 17, 16), (2, 512, 2, 2), (1, 8, 1, 1)])
def test_upfir2_add_and_its_adjoint(fwd, shape):
    """resnet skip branch: out = add + gain * upfirdn2d(v, f, up 2, pad [2,1,2,1], gain 4) (reference networks.py:245-250 through
    conv2d_resample's 1x1-up branch) and the adjoint of the FIR wrt v."""
    _lib, L, dt, ulp = _setup(fwd)
    B, C, h, w = shape
    gain = math.sqrt(0.5)
    v = util.case_tensor((B, C, h, w), 5).to(dt).float().requires_grad_(True)
    add = util.case_tensor((B, C, 2 * h, 2 * w), 6).to(dt).float()
    up = O.upfirdn2d(v, O.setup_filter([1, 3, 3, 1]), up=2, padding=[2, 1, 2, 1], gain=4) * gain
    ref = up + add
    out = torch.full((B, 2 * h, 2 * w, C), float("nan"), dtype=dt, device="cuda")
    vq, addq = _nhwc(v.detach(), dt), _nhwc(add, dt)          # keep the device tensors alive across the launches
    _lib.check(L.mgf_upfir2_add(vq.data_ptr(), addq.data_ptr(), out.data_ptr(), FK, 4.0 * gain, B, h, w, C,
                                _lib.stream_ptr()), "mgf_upfir2_add")
    got = _nchw(out)
    assert ((got - ref.detach()).abs() <= ulp * ref.detach().abs() + 1e-5).all(), (got - ref.detach()).abs().max()
    out2 = torch.full_like(out, float("nan"))
    _lib.check(L.mgf_upfir2_add(vq.data_ptr(), None, out2.data_ptr(), FK, 4.0 * gain, B, h, w, C, _lib.stream_ptr()), "mgf_upfir2_add")
    assert ((_nchw(out2) - up.detach()).abs() <= ulp * up.detach().abs() + 1e-5).all()
    dout = util.case_tensor((B, C, 2 * h, 2 * w), 7).to(torch.bfloat16).float()
    gref, = torch.autograd.grad(up, [v], dout)
    dv = torch.full((B, h, w, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    doutq = _nhwc(dout, torch.bfloat16)
    _lib.check(L.mgf_upfir2_bwd(doutq.data_ptr(), dv.data_ptr(), FK, 4.0 * gain, B, h, w, C, _lib.stream_ptr()), "mgf_upfir2_bwd")
    assert ((_nchw(dv) - gref).abs() <= 2.0 ** -7 * gref.abs() + 1e-5).all()


@pytest.mark.parametrize("fwd", ["fp16", "bf16"])
@pytest.mark.parametrize("cfg", [(2, 64, 32, 9, 12), (1, 128, 64, 16, 16), (3, 32, 32, 4, 4), (1, 512, 256, 8, 8)])
def test_transposed_conv_as_four_parity_gemms(fwd, cfg):
    """tc.conv_tc(phase_ntaps=(4,2,2,1)): ct = conv_transpose2d(x, w, stride 2) on the (2h+1) x (2w+1) grid inside a [2h+2, 2w+2] buffer
    (reference conv2d_resample.py:117-127 with flip_weight=False: the un-flipped weights go straight into conv_transpose2d)."""
    _lib, L, dt, ulp = _setup(fwd)
    from morphganformer_b200 import tc
    B, I, Oc, h, w = cfg
    x = util.case_tensor((B, I, h, w), 8).to(dt).float()
    W = (util.case_tensor((Oc, I, 3, 3), 9) / math.sqrt(9 * I)).to(dt).float()
    ref = F.conv_transpose2d(x, W.transpose(0, 1), stride=2)                                     # [B, O, 2h+1, 2w+1]
    taps = ([(0, -a, -b, (2 * a) * 3 + 2 * b) for a in (0, 1) for b in (0, 1)] + [(0, -a, 0, (2 * a) * 3 + 1) for a in (0, 1)]
            + [(0, 0, -b, 3 + 2 * b) for b in (0, 1)] + [(0, 0, 0, 4)])
    wk = W.reshape(Oc, I, 9).permute(2, 0, 1).reshape(1, 9, Oc, I).to(dt).contiguous().cuda()
    ct = torch.full((B, 2 * h + 2, 2 * w + 2, Oc), float("nan"), dtype=dt, device="cuda")
    xq = _nhwc(x, dt)
    tc.conv_tc([xq], wk, taps, (B, h + 1, w + 1), 4, Oc, ct, osy=2, osx=2, ofy=(0, 0, 1, 1), ofx=(0, 1, 0, 1), phase_ntaps=(4, 2, 2, 1))
    got = _nchw(ct)
    assert (got[:, :, 2 * h + 1] == 0).all() and (got[:, :, :, 2 * w + 1] == 0).all()
    got = got[:, :, :2 * h + 1, :2 * w + 1]
    scale = ref.abs().max().item()
    assert ((got - ref).abs() <= 2 * ulp * ref.abs() + 2e-3 * ulp * 128 * scale + 1e-6).all(), ((got - ref).abs().max().item(), scale)

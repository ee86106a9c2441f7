"""GPU parity tests proper: every op of the torch_utils.ops mirror, called through the C ABI on cuda:0, against the
oracle (oracle/ops.py, CPU) on the same seeded inputs and against the committed golden vectors from the real
reference.  Tolerances: fp32 1e-5 abs (exact-fp32 kernels), bf16/fp16 by ulp of the dtype."""
import os
import numpy as np
import pytest
import torch
from oracle import ops as O
import util

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(util.GOLDEN, "ops_golden.npz"))


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("i", range(len(util.UPFIRDN_CASES)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upfirdn2d(i, dtype):
    from morphganformer_b200.torch_utils.ops import upfirdn2d as U
    name, xs, taps, kw = util.UPFIRDN_CASES[i]
    x = util.case_tensor(xs, 100 + i)
    f = U.setup_filter(taps) if taps is not None else None
    xg = x.to(_dev(), dtype).requires_grad_(True)
    y = U.upfirdn2d(xg, f.to(_dev()) if f is not None else None, **kw)
    tol = 1e-5 if dtype == torch.float32 else 0.05
    ref = G["upfirdn/" + name]
    xr = x.to(dtype).float() if dtype != torch.float32 else x
    ref_t = O.upfirdn2d(xr, f, **kw)
    np.testing.assert_allclose(y.float().cpu().detach().numpy(), ref_t.numpy(), rtol=0, atol=tol * max(1.0, float(ref_t.abs().max())))
    if dtype == torch.float32:
        np.testing.assert_allclose(y.cpu().detach().numpy(), ref, rtol=0, atol=1e-5)
        gy = util.case_tensor(y.shape, 500 + i)
        gx, = torch.autograd.grad((y * gy.to(_dev())).sum(), [xg])
        xr2 = x.clone().requires_grad_(True)
        gx_ref, = torch.autograd.grad((O.upfirdn2d(xr2, f, **kw) * gy).sum(), [xr2])
        np.testing.assert_allclose(gx.cpu().numpy(), gx_ref.numpy(), rtol=0, atol=1e-5)


def test_upfirdn2d_channels_last_and_empty():
    from morphganformer_b200.torch_utils.ops import upfirdn2d as U
    f = U.setup_filter([1, 3, 3, 1])
    x = util.case_tensor((2, 8, 9, 9), 1)
    y = U.upfirdn2d(x.to(_dev()).contiguous(memory_format=torch.channels_last), f.to(_dev()), up=2, padding=[2, 1, 2, 1], gain=4)
    np.testing.assert_allclose(y.cpu().numpy(), O.upfirdn2d(x, f, up=2, padding=[2, 1, 2, 1], gain=4).numpy(), atol=1e-5, rtol=0)
    assert y.is_contiguous(memory_format=torch.channels_last)
    e = U.upfirdn2d(torch.zeros(0, 3, 8, 8, device=_dev()), f.to(_dev()), up=2, padding=[2, 1, 2, 1])
    assert tuple(e.shape) == (0, 3, 16, 16)
    with pytest.raises(RuntimeError):
        U.upfirdn2d(torch.zeros(1, 1, 2, 2, device=_dev()), f.to(_dev()))  # output would be empty


@pytest.mark.parametrize("act", util.BIAS_ACT_ACTS)
@pytest.mark.parametrize("clamp", [None, 0.4])
def test_bias_act_fp32_all_orders(act, clamp):
    from morphganformer_b200.torch_utils.ops import bias_act as B
    x = util.case_tensor((3, 5, 4, 6), 7).to(_dev()).requires_grad_(True)
    b = util.case_tensor((5,), 8).to(_dev()).requires_grad_(True)
    y = B.bias_act(x, b, dim=1, act=act, clamp=clamp)
    key = f"bias_act/{act}/{'clamp' if clamp else 'noclamp'}"
    np.testing.assert_allclose(y.detach().cpu().numpy(), G[key + "/y"], rtol=0, atol=2e-6)
    gy = util.case_tensor(y.shape, 9).to(_dev())
    gx, gb = torch.autograd.grad((y * gy).sum(), [x, b], create_graph=True)
    np.testing.assert_allclose(gx.detach().cpu().numpy(), G[key + "/gx"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(gb.detach().cpu().numpy(), G[key + "/gb"], rtol=0, atol=1e-4)
    if gx.requires_grad:
        ggx, = torch.autograd.grad((gx * util.case_tensor(gx.shape, 10).to(_dev())).sum(), [x], allow_unused=True)
        ggx = ggx if ggx is not None else torch.zeros_like(x)
        np.testing.assert_allclose(ggx.cpu().numpy(), G[key + "/ggx"], rtol=0, atol=2e-5)


def test_bias_act_dims_dtypes_edges():
    from morphganformer_b200.torch_utils.ops import bias_act as B
    x = util.case_tensor((3, 5, 4, 6), 7)
    y = B.bias_act(x.permute(0, 2, 3, 1).contiguous().to(_dev()), util.case_tensor((6,), 11).to(_dev()), dim=2, act="lrelu", alpha=0.1, gain=0.7)
    np.testing.assert_allclose(y.cpu().numpy(), G["bias_act/dim2_alpha_gain/y"], rtol=0, atol=2e-6)
    # fp64 goes through the same ABI whose alpha/gain/clamp are C floats (like the reference plugin, bias_act.cpp:24), hence 1e-6
    for dt, tol in ((torch.bfloat16, 2e-2), (torch.float16, 3e-3), (torch.float64, 1e-6)):
        xx = util.case_tensor((2, 7, 5, 3), 3).to(dt)          # odd sizes: exercises the scalar tail + per-element bias index
        bb = util.case_tensor((7,), 4).to(dt)
        yy = B.bias_act(xx.to(_dev()), bb.to(_dev()), act="lrelu")
        ref = O.bias_act(xx.double(), bb.double(), act="lrelu")
        np.testing.assert_allclose(yy.double().cpu().numpy(), ref.numpy(), rtol=tol, atol=tol)
    xl = x.to(_dev()).contiguous(memory_format=torch.channels_last)
    bl = util.case_tensor((5,), 8).to(_dev())
    np.testing.assert_allclose(B.bias_act(xl, bl, act="relu").cpu().numpy(), O.bias_act(x, bl.cpu(), act="relu").numpy(), atol=2e-6, rtol=0)
    assert B.bias_act(torch.zeros(0, 4, device=_dev()), torch.zeros(4, device=_dev()), act="lrelu").numel() == 0
    big = torch.randn(4, 32, 256, 256, device=_dev())
    np.testing.assert_allclose(B.bias_act(big, None, act="lrelu")[1, 3, 5, :8].cpu().numpy(),
                               O.bias_act(big[1, 3, 5, :8].cpu(), None, act="lrelu").numpy(), atol=1e-6, rtol=0)


CONV_CASES = [
    # N, IC, H, W, OC, k, stride, pad, dil, groups
    (2, 8, 9, 11, 12, 3, 1, 1, 1, 1), (1, 6, 10, 10, 9, 3, 2, 1, 1, 3), (2, 4, 7, 7, 6, 1, 1, 0, 1, 1),
    (1, 16, 12, 12, 16, 3, 1, 2, 2, 4), (3, 3, 8, 8, 5, 5, 2, 2, 1, 1), (1, 70, 6, 6, 130, 3, 1, 1, 1, 1),
]


@pytest.mark.parametrize("c", CONV_CASES)
def test_conv2d_gradfix_forward_and_grads(c):
    from morphganformer_b200.torch_utils.ops import conv2d_gradfix as C
    n, ic, h, w, oc, k, s, p, d, g = c
    x = util.case_tensor((n, ic, h, w), 1).requires_grad_(True)
    wt = (util.case_tensor((oc, ic // g, k, k), 2) * 0.2).requires_grad_(True)
    b = util.case_tensor((oc,), 3).requires_grad_(True)
    xr, wr, br = [t.detach().clone().to(_dev()).requires_grad_(True) for t in (x, wt, b)]
    y = C.conv2d(xr, wr, br, stride=s, padding=p, dilation=d, groups=g)
    yref = torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), stride=s, padding=p, dilation=d, groups=g)
    np.testing.assert_allclose(y.detach().cpu().numpy(), yref.detach().numpy(), rtol=1e-5, atol=2e-5)
    gy = util.case_tensor(y.shape, 4)
    gx, gw, gb = torch.autograd.grad((y * gy.to(_dev())).sum(), [xr, wr, br])
    rx, rw, rb = torch.autograd.grad((yref * gy.double()).sum(), [x, wt, b])
    np.testing.assert_allclose(gx.cpu().numpy(), rx.numpy(), rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(gw.cpu().numpy(), rw.numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(gb.cpu().numpy(), rb.numpy(), rtol=1e-5, atol=1e-4)


def test_conv_transpose2d_and_double_backward():
    from morphganformer_b200.torch_utils.ops import conv2d_gradfix as C
    x = util.case_tensor((2, 6, 5, 5), 1).requires_grad_(True)
    wt = (util.case_tensor((6, 4, 3, 3), 2) * 0.3).requires_grad_(True)     # groups=2: [IC, OC/g, kh, kw]
    xr, wr = [t.detach().clone().to(_dev()).requires_grad_(True) for t in (x, wt)]
    y = C.conv_transpose2d(xr, wr, stride=2, padding=0, groups=2)
    yref = torch.nn.functional.conv_transpose2d(x, wt, stride=2, padding=0, groups=2)
    np.testing.assert_allclose(y.detach().cpu().numpy(), yref.detach().numpy(), rtol=0, atol=2e-5)
    gy = util.case_tensor(y.shape, 4)
    gx, gw = torch.autograd.grad((y * gy.to(_dev())).sum(), [xr, wr], create_graph=True)
    rx, rw = torch.autograd.grad((yref * gy).sum(), [x, wt], create_graph=True)
    np.testing.assert_allclose(gx.detach().cpu().numpy(), rx.detach().numpy(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(gw.detach().cpu().numpy(), rw.detach().numpy(), rtol=0, atol=1e-4)
    # second order: d/dw of sum(gx^2) (what path-length / R1 regularisers need)
    g2, = torch.autograd.grad(gx.square().sum(), [wr])
    r2, = torch.autograd.grad(rx.square().sum(), [wt])
    np.testing.assert_allclose(g2.cpu().numpy(), r2.numpy(), rtol=0, atol=1e-3)
    with C.no_weight_gradients():
        y2 = C.conv2d(xr, torch.ones(4, 6, 1, 1, device=_dev(), requires_grad=True))
        assert torch.autograd.grad(y2.sum(), [xr])[0] is not None


@pytest.mark.parametrize("i", range(len(util.RESAMPLE_CASES)))
def test_conv2d_resample(i):
    from morphganformer_b200.torch_utils.ops import conv2d_resample as R, upfirdn2d as U
    name, xs, ws, kw = util.RESAMPLE_CASES[i]
    kw = dict(kw)
    f = kw.pop("f", None)
    f = U.setup_filter(f).to(_dev()) if f is not None else None
    x = util.case_tensor(xs, 200 + i).to(_dev()).requires_grad_(True)
    w = (util.case_tensor(ws, 300 + i) * 0.2).to(_dev()).requires_grad_(True)
    y = R.conv2d_resample(x, w, f=f, **kw)
    np.testing.assert_allclose(y.detach().cpu().numpy(), G[f"resample/{name}/y"], rtol=0, atol=2e-5)
    gx, gw = torch.autograd.grad((y * util.case_tensor(y.shape, 400 + i).to(_dev())).sum(), [x, w])
    np.testing.assert_allclose(gx.cpu().numpy(), G[f"resample/{name}/gx"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(gw.cpu().numpy(), G[f"resample/{name}/gw"], rtol=0, atol=2e-4)


def test_fma():
    from morphganformer_b200.torch_utils.ops import fma as F_
    a, b, c = util.case_tensor((2, 3, 4, 5), 20), util.case_tensor((2, 3, 1, 1), 21), util.case_tensor((4, 5), 22)
    ag, bg, cg = [t.to(_dev()).requires_grad_(True) for t in (a, b, c)]
    y = F_.fma(ag, bg, cg)
    np.testing.assert_allclose(y.detach().cpu().numpy(), G["fma/y"], rtol=0, atol=1e-6)
    ga, gb, gc = torch.autograd.grad(y.sum(), [ag, bg, cg])
    assert ga.shape == a.shape and gb.shape == b.shape and gc.shape == c.shape
    np.testing.assert_allclose(gb.cpu().numpy(), a.sum(dim=[2, 3], keepdim=True).numpy(), rtol=1e-5, atol=1e-5)

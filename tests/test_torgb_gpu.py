"""GPU: mgf_torgb_bwd (reference ToRGBLayer.forward training/networks.py:1054-1065, backward of the modulated 1x1 convolution without
demodulation) against PyTorch fp32 autograd."""
import pytest
import torch

from morphganformer_b200 import _lib

pytestmark = pytest.mark.gpu


def _p(t):
    return t.data_ptr()


@pytest.mark.parametrize("B,HW,C", [(2, 37 * 41, 32), (1, 64 * 64, 64), (2, 300, 512), (3, 5000, 128)])
def test_torgb_bwd_vs_autograd(B, HW, C):
    L = _lib.lib()
    try:
        g = torch.Generator(device="cuda").manual_seed(C + HW)
        y = torch.randn(B, HW, C, device="cuda", generator=g).to(torch.float16)
        wrgb = torch.randn(3, C, device="cuda", generator=g) * 0.2
        s = torch.randn(B, C, device="cuda", generator=g)
        dimg = torch.randn(B, 3, HW, device="cuda", generator=g)
        dy = torch.empty(B, HW, C, device="cuda", dtype=torch.bfloat16); ds = torch.zeros(B, C, device="cuda"); R = torch.zeros(B, C, device="cuda")
        _lib.set_forward_dtype("fp16")
        _lib.check(L.mgf_torgb_bwd(_p(dimg), _p(y), _p(wrgb), _p(s), _p(dy), _p(ds), _p(R), B, HW, C, torch.cuda.current_stream().cuda_stream), "torgb_bwd")
        torch.cuda.synchronize()
        yf = y.float().requires_grad_(True); sf = s.clone().requires_grad_(True)
        img = torch.einsum("bpc,oc,bc->bop", yf, wrgb, sf)                  # img[b,o,p] = sum_c y[b,p,c] w[o,c] s[b,c]
        gy, gs = torch.autograd.grad(img, [yf, sf], grad_outputs=dimg)
        assert (dy.float() - gy).abs().max().item() <= 2.0 ** -7 * gy.abs().max().item()
        torch.testing.assert_close(ds, gs, rtol=2e-4, atol=2e-3 * gs.abs().max().item())
        torch.testing.assert_close(R, (gy * y.float()).sum(1), rtol=2e-4, atol=2e-3 * gs.abs().max().item())
    finally:
        _lib.set_forward_dtype(_lib.DEFAULT_FORWARD_DTYPE)

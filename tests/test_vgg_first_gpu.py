"""Fused LPIPS input stage + VGG conv1_1 (mgf_vgg_conv1_fwd / mgf_vgg_conv1_bwd) against PyTorch fp32:
ScalingLayer (reference lpips/networks_basic.py:94-101) -> conv3x3(3->64) + bias -> ReLU (pretrained_networks.py slice1), and the
MSE sum / gradient of the projection loss.  Operands are 16-bit on the tensor cores: tolerance 2^-9 (fp16) / 2^-6 (bf16) of the scale."""
import pytest
import torch

from morphganformer_b200 import _lib

pytestmark = pytest.mark.gpu
SHIFT = torch.tensor([-.030, -.088, -.188]).view(1, 3, 1, 1)
SCALE = torch.tensor([.458, .448, .450]).view(1, 3, 1, 1)


def _p(t):
    return t.data_ptr() if t is not None else None


@pytest.mark.parametrize("fwd", ["fp16", "bf16"])
@pytest.mark.parametrize("B,R", [(1, 16), (2, 40), (2, 64), (1, 200), (2, 520)])      # last: ragged AND several tiles per persistent CTA (register prefetch path)
def test_vgg_conv1_fwd_bwd(B, R, fwd):
    L = _lib.lib()
    _lib.set_forward_dtype(fwd)
    try:
        dt = torch.float16 if fwd == "fp16" else torch.bfloat16
        g = torch.Generator(device="cuda").manual_seed(R)
        img = torch.tanh(torch.randn(B, 3, R, R, device="cuda", generator=g)); tgt = torch.tanh(torch.randn(B, 3, R, R, device="cuda", generator=g))
        w = torch.randn(64, 3, 3, 3, device="cuda", generator=g) * 0.2; bias = torch.randn(64, device="cuda", generator=g) * 0.1
        wc = torch.zeros(64, 32, device="cuda"); wc[:, :27] = w.permute(0, 2, 3, 1).reshape(64, 27)
        out = torch.empty(B, R, R, 64, device="cuda", dtype=dt); mse = torch.zeros(B, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(L.mgf_vgg_conv1_fwd(_p(img), _p(tgt), _p(mse), _p(wc), _p(bias), _p(out), B, R, s), "fwd")
        gy = (torch.randn(B, R, R, 64, device="cuda", generator=g)).to(torch.bfloat16)
        dimg = torch.empty_like(img)
        _lib.check(L.mgf_vgg_conv1_bwd(_p(gy), _p(wc), _p(img), _p(tgt), 0.37, _p(dimg), B, R, s), "bwd")
        dimg2 = torch.empty_like(img)
        _lib.check(L.mgf_vgg_conv1_bwd(_p(gy), _p(wc), None, None, 0.0, _p(dimg2), B, R, s), "bwd")
        torch.cuda.synchronize()
        x = img.clone().requires_grad_(True)
        pre = torch.nn.functional.conv2d((x - SHIFT.cuda()) / SCALE.cuda(), w, bias, padding=1)
        ref = torch.relu(pre).permute(0, 2, 3, 1)
        eps = 2.0 ** -9 if fwd == "fp16" else 2.0 ** -6
        assert (out.float() - ref).abs().max().item() < eps * ref.abs().max().item()
        torch.testing.assert_close(mse, (img - tgt).square().sum(dim=[1, 2, 3]), rtol=1e-5, atol=1e-3)
        gx, = torch.autograd.grad(pre, [x], grad_outputs=gy.float().permute(0, 3, 1, 2))
        assert (dimg2 - gx).abs().max().item() < 2.0 ** -7 * gx.abs().max().item()
        assert (dimg - (gx + 0.37 * (img - tgt))).abs().max().item() < 2.0 ** -7 * gx.abs().max().item()
    finally:
        _lib.set_forward_dtype(_lib.DEFAULT_FORWARD_DTYPE)

"""GPU parity of the streaming 1x1 convolution (pointwise.cu; the resnet-skip Conv2dLayer of the reference, training/networks.py:245-250 with
kernel_size 1, and its input gradient) against an fp32 matmul of the same 16-bit-rounded operands: fp32 accumulation on both sides, so the
only difference is the final 16-bit store (1 ulp of the storage type at the result's magnitude) -- and against mgf_conv_tc on the same call."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fwd", ["fp16", "bf16"])
@pytest.mark.parametrize("is_fwd", [1, 0])
@pytest.mark.parametrize("P,K,N", [(16 * 16 * 2, 64, 32), (1000, 32, 64), (37, 128, 64), (4096, 64, 128), (513, 256, 128), (2048, 128, 256)])
def test_pointwise_vs_fp32_matmul(fwd, is_fwd, P, K, N):
    from morphganformer_b200 import _lib
    _lib.set_forward_dtype(fwd)
    try:
        L = _lib.lib()
        dt = (torch.float16 if fwd == "fp16" else torch.bfloat16) if is_fwd else torch.bfloat16
        ulp = 2.0 ** -10 if dt == torch.float16 else 2.0 ** -7
        g = torch.Generator().manual_seed(P + K + N)
        x = torch.randn(P, K, generator=g).to(dt).cuda()
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt).cuda()
        out = torch.full((P, N), float("nan"), dtype=dt, device="cuda")
        assert L.mgf_pointwise_supported(K, N) == 1
        _lib.check(L.mgf_pointwise(x.data_ptr(), w.data_ptr(), out.data_ptr(), P, K, N, is_fwd, _lib.stream_ptr(x.device)), "mgf_pointwise")
        ref = x.float() @ w.float().t()
        err = (out.float() - ref).abs()
        tol = ulp * ref.abs().clamp(min=1.0) + 1e-5
        assert torch.isfinite(out.float()).all()
        assert (err <= tol).all(), "max err %g at magnitude %g" % (err.max().item(), ref.abs().max().item())
    finally:
        _lib.set_forward_dtype(_lib.DEFAULT_FORWARD_DTYPE)


def test_pointwise_equals_conv_tc_one_tap():
    from morphganformer_b200 import _lib, tc
    L = _lib.lib()
    dt = _lib.forward_torch_dtype()
    B, h, K, N = 2, 32, 64, 32
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, h, h, K, generator=g).to(dt).cuda()
    w = (torch.randn(1, 1, N, K, generator=g) / 8).to(dt).cuda()
    a = torch.empty(B, h, h, N, dtype=dt, device="cuda"); b = torch.empty_like(a)
    tc.conv_tc([x], w, [(0, 0, 0, 0)], (B, h, h), 1, N, a)
    _lib.check(L.mgf_pointwise(x.data_ptr(), w.data_ptr(), b.data_ptr(), B * h * h, K, N, 1, _lib.stream_ptr(x.device)), "mgf_pointwise")
    # both accumulate in fp32 and round once; summation order differs
    assert (a.float() - b.float()).abs().max().item() <= 2.0 ** -9 * max(1.0, a.float().abs().max().item())


def test_pointwise_rejects_unsupported_shapes():
    from morphganformer_b200 import _lib
    L = _lib.lib()
    assert L.mgf_pointwise_supported(512, 512) == 0 and L.mgf_pointwise_supported(48, 32) == 0
    x = torch.zeros(16, 512, dtype=torch.bfloat16, device="cuda")
    rc = L.mgf_pointwise(x.data_ptr(), x.data_ptr(), x.data_ptr(), 16, 512, 512, 0, None)
    assert rc == -5 and b"outside the supported shapes" in L.mgf_last_error()

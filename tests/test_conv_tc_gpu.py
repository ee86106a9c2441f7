"""GPU parity: the tcgen05 implicit-GEMM convolution (mgf_conv_tc) against fp32 convolution of the same bf16-rounded
operands (torch CPU, the oracle's arithmetic).  bf16 products are exact in fp32, so only accumulation order and the final
bf16 store differ: tolerance = 2 bf16 ulps of the result magnitude."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
import util

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16)


def _check(out_nhwc, ref_nchw, name=""):
    got = out_nhwc.float().cpu().permute(0, 3, 1, 2)
    scale = ref_nchw.abs().max().item()
    err = (got - ref_nchw).abs().max().item()
    assert err <= 2 ** -7 * scale + 1e-6, "%s: max err %g vs scale %g" % (name, err, scale)


@pytest.mark.parametrize("cfg", [
    # B, H, W, Cin, Cout, bn
    (2, 16, 16, 64, 64, 0), (1, 32, 32, 128, 256, 0), (3, 8, 8, 64, 128, 0), (2, 4, 4, 64, 64, 0),
    (1, 64, 64, 32, 32, 0), (2, 16, 32, 512, 512, 0), (1, 16, 16, 64, 256, 128), (1, 40, 24, 64, 32, 0),
    (5, 16, 16, 64, 64, 32),
])
def test_conv3x3_shared_weights(cfg):
    from morphganformer_b200 import tc
    b, h, w, ci, co, bn = cfg
    x = _bf(util.case_tensor((b, h, w, ci), 1))
    wt = _bf(util.case_tensor((co, ci, 3, 3), 2) * (1.0 / np.sqrt(9 * ci)))
    out = torch.empty(b, h, w, co, dtype=torch.bfloat16, device="cuda")
    tc.conv_tc([x.cuda()], tc.pack_w3x3(wt).cuda(), tc.TAPS_3X3, (b, h, w), 1, co, out, bn=bn)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1)
    _check(out, ref, str(cfg))


def test_conv3x3_per_sample_weights_and_epilogue():
    from morphganformer_b200 import tc
    b, h, w, ci, co = 3, 16, 16, 64, 128
    x = _bf(util.case_tensor((b, h, w, ci), 1))
    wt = _bf(util.case_tensor((b, co, ci, 3, 3), 2) * (1.0 / np.sqrt(9 * ci)))
    noise = util.case_tensor((h, w), 3)
    ns = torch.tensor(0.3)
    bias = util.case_tensor((co,), 4)
    add = _bf(util.case_tensor((b, h, w, co), 5))
    wp = wt.permute(0, 3, 4, 1, 2).reshape(b, 9, co, ci).contiguous()
    out = torch.empty(b, h, w, co, dtype=torch.bfloat16, device="cuda")
    tc.conv_tc([x.cuda()], wp.cuda(), tc.TAPS_3X3, (b, h, w), 1, co, out, noise=noise.cuda(), noise_strength=ns.cuda(),
               bias=bias.cuda(), act=1, alpha=0.2, gain=1.3, add=add.cuda())
    torch.cuda.synchronize()
    ref = torch.cat([F.conv2d(x[i:i + 1].float().permute(0, 3, 1, 2), wt[i].float(), padding=1) for i in range(b)])
    ref = F.leaky_relu(ref + noise * ns + bias.reshape(1, -1, 1, 1), 0.2) * 1.3 + add.float().permute(0, 3, 1, 2)
    _check(out, ref)


def test_dgrad_style_epilogue_scale_reduce_actgrad():
    from morphganformer_b200 import tc
    b, h, w, ci, co = 2, 16, 16, 128, 64          # "dgrad": A = dy [.., ci], output channels co
    dy = _bf(util.case_tensor((b, h, w, ci), 1))
    wt = _bf(util.case_tensor((b, co, ci, 3, 3), 2) * (1.0 / np.sqrt(9 * ci)))
    X = _bf(util.case_tensor((b, h, w, co), 3))
    s = util.case_tensor((b, co), 4)
    red = torch.zeros(b, co, device="cuda")
    wp = wt.permute(0, 3, 4, 1, 2).reshape(b, 9, co, ci).contiguous()
    out = torch.empty(b, h, w, co, dtype=torch.bfloat16, device="cuda")
    tc.conv_tc([dy.cuda()], wp.cuda(), tc.TAPS_3X3, (b, h, w), 1, co, out, scale_n=s.cuda(), reduce_out=red, X=X.cuda(),
               actgrad=True, ag_alpha=0.2, ag_gain=1.4, reduce_per_sample=True)
    torch.cuda.synchronize()
    acc = torch.cat([F.conv2d(dy[i:i + 1].float().permute(0, 3, 1, 2), wt[i].float(), padding=1) for i in range(b)])
    Xn = X.float().permute(0, 3, 1, 2)
    ref_red = (acc * Xn).sum(dim=[2, 3])
    ref = acc * s.reshape(b, co, 1, 1) * torch.where(Xn > 0, 1.0, 0.2) * 1.4
    _check(out, ref)
    assert (red.cpu() - ref_red).abs().max() <= 1e-3 * ref_red.abs().max() + 1e-3


def test_phase_output_and_multi_amap():
    """4-phase output mapping (up-sampling conv form) and a 2-map K-concatenation."""
    from morphganformer_b200 import tc
    b, h, w, ci, co = 2, 8, 8, 64, 64
    x = _bf(util.case_tensor((b, h, w, ci), 1))
    wt = _bf(util.case_tensor((4, co, ci, 3, 3), 2) * (1.0 / np.sqrt(9 * ci)))       # one 3x3 kernel per output phase
    wp = wt.permute(3, 4, 0, 1, 2).reshape(1, 9, 4 * co, ci).contiguous()
    out = torch.zeros(b, 2 * h, 2 * w, co, dtype=torch.bfloat16, device="cuda")
    tc.conv_tc([x.cuda()], wp.cuda(), tc.TAPS_3X3, (b, h, w), 4, co, out, osy=2, osx=2, ofy=(0, 0, 1, 1), ofx=(0, 1, 0, 1))
    torch.cuda.synchronize()
    ref = torch.zeros(b, co, 2 * h, 2 * w)
    for ph, (py, px) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]):
        ref[:, :, py::2, px::2] = F.conv2d(x.float().permute(0, 3, 1, 2), wt[ph].float(), padding=1)
    _check(out, ref)
    # strided phase views as A operands: out = sum_ph conv3x3(big[:, py::2, px::2], w[ph])
    big = _bf(util.case_tensor((b, 2 * h, 2 * w, ci), 7)).cuda()
    acts = [tc.phase_view(big, py, px) for (py, px) in [(0, 0), (0, 1), (1, 0), (1, 1)]]
    taps = [(ph, ky - 1, kx - 1, ph * 9 + ky * 3 + kx) for ph in range(4) for ky in range(3) for kx in range(3)]
    wp2 = wt.permute(0, 3, 4, 1, 2).reshape(1, 36, co, ci).contiguous()
    out2 = torch.empty(b, h, w, co, dtype=torch.bfloat16, device="cuda")
    tc.conv_tc(acts, wp2.cuda(), taps, (b, h, w), 1, co, out2)
    torch.cuda.synchronize()
    bigc = big.float().cpu().permute(0, 3, 1, 2)
    ref2 = sum(F.conv2d(bigc[:, :, py::2, px::2], wt[ph].float(), padding=1) for ph, (py, px) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]))
    _check(out2, ref2)

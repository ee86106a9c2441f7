"""GPU parity: the tcgen05 implicit-GEMM convolution (mgf_conv_tc) against fp32 convolution of the same bf16-rounded
operands (torch CPU, the oracle's arithmetic).  bf16 products are exact in fp32, so only accumulation order and the final
bf16 store differ: tolerance = 2 bf16 ulps of the result magnitude."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
import util

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _bf16_operands():
    """Kernel-level tests of mgf_conv_tc build their operands as bf16 tensors: run the library in bf16 forward storage here (the fp16
    forward mode of the same kernel -- one instruction-descriptor bit and the pack/unpack type -- is covered by the engine tests)."""
    from morphganformer_b200 import _lib
    _lib.set_forward_dtype("bf16")
    yield


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


# ---- halo variant (C = 64/128, images >= 256^2): same contract, different kernel; also A/B against the generic kernel
@pytest.mark.parametrize("cfg", [
    # B, H, W, Cin, Cout
    (1, 256, 256, 64, 64), (2, 256, 256, 64, 64), (1, 256, 256, 64, 128), (1, 272, 264, 64, 32), (1, 256, 512, 64, 256),
    (2, 256, 256, 32, 32), (1, 264, 272, 32, 64), (1, 256, 256, 128, 64), (2, 264, 272, 128, 64),      # last two: two 64-channel chunks, resident weights
])
def test_conv3x3_halo_variant(cfg):
    from morphganformer_b200 import tc, _lib
    b, h, w, ci, co = cfg
    x = _bf(util.case_tensor((b, h, w, ci), 1))
    wt = _bf(util.case_tensor((co, ci, 3, 3), 2) * (1.0 / np.sqrt(9 * ci)))
    bias = util.case_tensor((co,), 3)
    outs = []
    for halo in (1, 0):
        _lib.lib().mgf_conv_tc_set_halo(halo)
        out = torch.empty(b, h, w, co, dtype=torch.bfloat16, device="cuda")
        tc.conv_tc([x.cuda()], tc.pack_w3x3(wt).cuda(), tc.TAPS_3X3, (b, h, w), 1, co, out, bias=bias.cuda(), act=2)
        torch.cuda.synchronize()
        outs.append(out)
    _lib.lib().mgf_conv_tc_set_halo(1)
    ref = torch.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1))
    _check(outs[1], ref, "generic " + str(cfg))
    _check(outs[0], ref, "halo " + str(cfg))


def test_halo_per_sample_dgrad_epilogue_and_phases():
    from morphganformer_b200 import tc
    b, h, w, ci, co = 2, 256, 256, 64, 64
    dy = _bf(util.case_tensor((b, h, w, ci), 1))
    wt = _bf(util.case_tensor((b, co, ci, 3, 3), 2) * (1.0 / np.sqrt(9 * ci)))
    X = _bf(util.case_tensor((b, h, w, co), 3))
    s = util.case_tensor((b, co), 4)
    red = torch.zeros(b, co, device="cuda")
    wp = wt.permute(0, 3, 4, 1, 2).reshape(b, 9, co, ci).contiguous()
    taps_b = [(0, 1 - ky, 1 - kx, ky * 3 + kx) for ky in range(3) for kx in range(3)]     # flipped taps, as the engine's dgrad uses
    out = torch.empty(b, h, w, co, dtype=torch.bfloat16, device="cuda")
    tc.conv_tc([dy.cuda()], wp.cuda(), taps_b, (b, h, w), 1, co, out, scale_n=s.cuda(), reduce_out=red, X=X.cuda(),
               actgrad=True, ag_alpha=0.2, ag_gain=1.4, reduce_per_sample=True)
    torch.cuda.synchronize()
    acc = torch.cat([F.conv2d(dy[i:i + 1].float().permute(0, 3, 1, 2), wt[i].float().flip([2, 3]), padding=1) for i in range(b)])
    Xn = X.float().permute(0, 3, 1, 2)
    _check(out, acc * s.reshape(b, co, 1, 1) * torch.where(Xn > 0, 1.0, 0.2) * 1.4)
    ref_red = (acc * Xn).sum(dim=[2, 3])
    assert (red.cpu() - ref_red).abs().max() <= 2e-3 * ref_red.abs().max() + 1e-2
    # four output phases (up-conv form) through the halo kernel: 64 -> 32 channels, 256^2 -> 512^2
    co2 = 32
    x = _bf(util.case_tensor((1, h, w, ci), 5))
    w4 = _bf(util.case_tensor((4, co2, ci, 3, 3), 6) * (1.0 / np.sqrt(9 * ci)))
    wp4 = w4.permute(3, 4, 0, 1, 2).reshape(1, 9, 4 * co2, ci).contiguous()
    out4 = torch.zeros(1, 2 * h, 2 * w, co2, dtype=torch.bfloat16, device="cuda")
    tc.conv_tc([x.cuda()], wp4.cuda(), tc.TAPS_3X3, (1, h, w), 4, co2, out4, osy=2, osx=2, ofy=(0, 0, 1, 1), ofx=(0, 1, 0, 1))
    torch.cuda.synchronize()
    ref4 = torch.zeros(1, co2, 2 * h, 2 * w)
    for ph, (py, px) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]):
        ref4[:, :, py::2, px::2] = F.conv2d(x.float().permute(0, 3, 1, 2), w4[ph].float(), padding=1)
    _check(out4, ref4)


@pytest.mark.parametrize("cfg", [(2, 40, 24, 64, 32), (1, 32, 32, 128, 64), (2, 16, 16, 256, 128), (1, 8, 8, 64, 256)])
@pytest.mark.parametrize("halo_mode", [1, 9])
def test_four_phase_upconv_form_in_one_tile_with_fused_tail(cfg, halo_mode):
    """FIR-folded up-convolution form: 4 output phases x Cout columns.  The tile spans whole phases when 4 * Cout fits 128 / 256 columns
    (32-channel phases use the 32-column staged groups), so activation tiles are fetched once for the four phases; noise (per phase!),
    bias and leaky-ReLU run in the epilogue.  halo_mode 9 forces the older one-phase-per-tile halo kernel where it applies (same result)."""
    from morphganformer_b200 import tc, _lib
    b, h, w, ci, co = cfg
    x = _bf(util.case_tensor((b, h, w, ci), 11))
    w4 = _bf(util.case_tensor((4, co, ci, 3, 3), 12) * (1.0 / np.sqrt(9 * ci)))
    wp4 = w4.permute(3, 4, 0, 1, 2).reshape(1, 9, 4 * co, ci).contiguous()
    noise = util.case_tensor((2 * h, 2 * w), 13)
    nstr = torch.tensor([0.4])
    bias = util.case_tensor((co,), 14) * 0.3
    scale = util.case_tensor((b, 4 * co), 15).abs() + 0.5
    out4 = torch.full((b, 2 * h, 2 * w, co), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.lib().mgf_conv_tc_set_halo(halo_mode)
    try:
        tc.conv_tc([x.cuda()], wp4.cuda(), tc.TAPS_3X3, (b, h, w), 4, co, out4, osy=2, osx=2, ofy=(0, 0, 1, 1), ofx=(0, 1, 0, 1),
                   scale_n=scale.cuda(), noise=noise.cuda(), noise_strength=nstr.cuda(), bias=bias.cuda(), act=1, alpha=0.2, gain=1.3)
        torch.cuda.synchronize()
    finally:
        _lib.lib().mgf_conv_tc_set_halo(1)
    ref4 = torch.zeros(b, co, 2 * h, 2 * w)
    for ph, (py, px) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]):
        y = F.conv2d(x.float().permute(0, 3, 1, 2), w4[ph].float(), padding=1) * scale[:, ph * co:(ph + 1) * co].reshape(b, co, 1, 1)
        ref4[:, :, py::2, px::2] = y
    ref4 = F.leaky_relu(ref4 + noise * nstr + bias.reshape(1, co, 1, 1), 0.2) * 1.3
    _check(out4, ref4)


@pytest.mark.parametrize("cfg", [(2, 32, 32, 64, 128), (1, 64, 48, 128, 256), (3, 16, 16, 256, 512), (2, 24, 40, 512, 256), (1, 16, 8, 64, 128)])
def test_cta_pair_kernel_equals_single_cta_kernel(cfg):
    """Wide tiles (BN = 128 / 256) run on the CTA-pair kernel (tcgen05.mma.cta_group::2, M = 256 over two SMs, each CTA staging half of the weight
    tile): same result as the one-CTA kernel (mode bit 4 switches the pair kernel off) and as the fp32 reference, including odd tile counts
    (the last pair's second tile falls outside the tensor), the bias + ReLU tail and the shared-weight d(style)-free dgrad form."""
    from morphganformer_b200 import tc, _lib
    b, h, w, ci, co = cfg
    x = _bf(util.case_tensor((b, h, w, ci), 21))
    wt = _bf(util.case_tensor((co, ci, 3, 3), 22) * (1.0 / np.sqrt(9 * ci)))
    bias = util.case_tensor((co,), 23) * 0.2
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1) + bias.reshape(1, co, 1, 1))
    outs = []
    try:
        for mode in (1, 1 | 16):
            _lib.lib().mgf_conv_tc_set_halo(mode)
            out = torch.full((b, h, w, co), float("nan"), dtype=torch.bfloat16, device="cuda")
            tc.conv_tc([x.cuda()], tc.pack_w3x3(wt).cuda(), tc.TAPS_3X3, (b, h, w), 1, co, out, bias=bias.cuda(), act=2, gain=1.0)
            torch.cuda.synchronize()
            outs.append(out.float().cpu())
    finally:
        _lib.lib().mgf_conv_tc_set_halo(1)
    assert torch.equal(outs[0], outs[1]), (outs[0] - outs[1]).abs().max()
    _check(outs[0], ref)


def test_halo_two_chunk_variant_with_saved_activation_mask():
    """128 -> 64 channels through the two-chunk halo kernel with the ReLU mask of a saved activation in the tail (the shape of the input
    gradient of VGG conv2_1 feeding conv1_2's ReLU backward): X is read per thread here (one staging tile per group), so halo and generic
    kernels must agree exactly up to summation order."""
    from morphganformer_b200 import tc, _lib
    b, h, w, ci, co = 2, 256, 256, 128, 64
    x = _bf(util.case_tensor((b, h, w, ci), 11))
    wt = _bf(util.case_tensor((co, ci, 3, 3), 12) * (1.0 / np.sqrt(9 * ci)))
    X = util.case_tensor((b, h, w, co), 13).to(_lib.forward_torch_dtype()).cuda()
    outs = []
    for halo in (1, 0):
        _lib.lib().mgf_conv_tc_set_halo(halo)
        out = torch.empty(b, h, w, co, dtype=torch.bfloat16, device="cuda")
        tc.conv_tc([x.cuda()], tc.pack_w3x3(wt).to(torch.bfloat16).cuda(), tc.TAPS_3X3, (b, h, w), 1, co, out, X=X, actgrad=True, ag_alpha=0.0, fwd=False)
        torch.cuda.synchronize()
        outs.append(out)
    _lib.lib().mgf_conv_tc_set_halo(1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1) * (X.float().cpu().permute(0, 3, 1, 2) > 0)
    _check(outs[1], ref, "generic")
    _check(outs[0], ref, "halo KC=2")


@pytest.mark.parametrize("tail", ["plain", "fwd", "bwd"])
def test_cta_pair_halo_kernel_equals_one_cta_halo_kernel(tail):
    """The CTA-pair form of the halo kernel (tcgen05.mma.cta_group::2, half of the resident weights per CTA; off by default because it measured
    slower, mgf_conv_tc_set_halo bit 6) against the one-CTA halo kernel on the same launch: identical arithmetic, so bit-identical outputs
    (the d(style) reduction differs by atomic order only)."""
    from morphganformer_b200 import tc, _lib
    b, h, w, c = 2, 512, 256, 64
    x = _bf(util.case_tensor((b, h, w, c), 21)).cuda()
    wt = _bf(util.case_tensor((b, 9, c, c), 22) * (1.0 / np.sqrt(9 * c))).cuda()
    bias = util.case_tensor((c,), 23).cuda()
    X = util.case_tensor((b, h, w, c), 24).to(_lib.forward_torch_dtype()).cuda()
    s2 = (util.case_tensor((b, c), 25).abs() + 0.5).cuda()
    outs, reds = [], []
    for mode in (1 | 64, 1):
        _lib.lib().mgf_conv_tc_set_halo(mode)
        out = torch.empty(b, h, w, c, dtype=torch.bfloat16, device="cuda")
        red = torch.zeros(b, c, device="cuda")
        if tail == "plain":
            tc.conv_tc([x], wt, tc.TAPS_3X3, (b, h, w), 1, c, out, fwd=False)
        elif tail == "fwd":
            tc.conv_tc([x], wt, tc.TAPS_3X3, (b, h, w), 1, c, out, bias=bias, act=1, gain=1.4, fwd=False)
        else:
            tc.conv_tc([x], wt, tc.TAPS_3X3, (b, h, w), 1, c, out, scale_n=s2, reduce_out=red, X=X, actgrad=True, ag_gain=1.4, reduce_per_sample=True, fwd=False)
        torch.cuda.synchronize()
        outs.append(out); reds.append(red)
    _lib.lib().mgf_conv_tc_set_halo(1)
    assert torch.equal(outs[0], outs[1])
    assert torch.allclose(reds[0], reds[1], rtol=1e-4, atol=1e-3)

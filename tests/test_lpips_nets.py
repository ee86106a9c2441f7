"""LPIPS alex / squeeze: oracle vs golden vectors from the real reference (CPU), product module vs golden (GPU)."""
import os
import numpy as np
import pytest
import torch

import util
from oracle import lpips_ref, refimport

GOLD = np.load(os.path.join(util.GOLDEN, "lpips_nets_golden.npz"))


def _inputs():
    return torch.tanh(util.case_tensor((2, 3, 96, 96), 31)), torch.tanh(util.case_tensor((2, 3, 96, 96), 32))


@pytest.mark.parametrize("net", ["alex", "squeeze"])
def test_oracle_matches_reference_golden(net):
    sd = util.build_lpips_sd(net, 4)
    a, b = _inputs()
    a.requires_grad_(True)
    d = lpips_ref.lpips(sd, a, b, net)
    ga, = torch.autograd.grad(d.sum(), [a])
    np.testing.assert_allclose(d.detach().numpy(), GOLD[net + "_d"], rtol=2e-5)
    np.testing.assert_allclose(ga.numpy(), GOLD[net + "_ga"], rtol=1e-3, atol=1e-8)


@pytest.mark.skipif(not refimport.available(), reason="reference tree not present")
@pytest.mark.parametrize("net", ["alex", "squeeze"])
def test_seeded_state_dict_equals_reference(net):
    ref = {k: v for k, v in refimport.build_lpips(seed=4, net_type=net).state_dict().items() if k.startswith(("net.", "lin"))}
    mine = util.build_lpips_sd(net, 4)
    assert sorted(ref) == sorted(mine)
    for k in ref:
        assert torch.equal(ref[k], mine[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("net", ["alex", "squeeze"])
def test_product_matches_reference_golden_gpu(net):
    from morphganformer_b200.lpips_nets import PerceptualLoss
    P = PerceptualLoss(util.build_lpips_sd(net, 4), model="net-lin", net=net).cuda()
    a, b = _inputs()
    a = a.cuda().requires_grad_(True)
    d = P(a, b.cuda())
    assert tuple(d.shape) == (2, 1, 1, 1)
    ga, = torch.autograd.grad(d.sum(), [a])
    np.testing.assert_allclose(d.detach().cpu().numpy(), GOLD[net + "_d"], rtol=1e-4)       # fp32 kernels, different summation order
    g = GOLD[net + "_ga"]
    assert np.abs(ga.cpu().numpy() - g).max() < 1e-3 * np.abs(g).max()


@pytest.mark.gpu
def test_downsample_prestep_gpu():
    """projection_example_v1.py:150-155: box-average to 256^2 before the distance (here 128 -> 64 with downsample_to=64)."""
    from morphganformer_b200.lpips_nets import LpipsNet
    sd = util.build_lpips_sd("alex", 4)
    a, b = torch.tanh(util.case_tensor((1, 3, 128, 128), 33)), torch.tanh(util.case_tensor((1, 3, 128, 128), 34))
    P = LpipsNet(sd, "alex", downsample_to=64).cuda()
    got = P(a.cuda(), b.cuda()).cpu()
    pool = lambda t: t.reshape(1, 3, 64, 2, 64, 2).mean([3, 5])
    want = lpips_ref.lpips(sd, pool(a), pool(b), "alex")
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4)

"""GPU parity: G.synthesis / G(z) through the ops engine (exact fp32 kernels) against the oracle and the golden vectors
from the real reference.  north_star tolerance for fp32 images: 1e-4 max-abs."""
import os
import numpy as np
import pytest
import torch
from oracle import ganformer
import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [(32, 512, 32, "gen32_golden.npz"), (64, 32768, 512, "gen64_golden.npz")])
def test_ops_engine_image_matches_reference_golden(cfg):
    res, cb, cm, fn = cfg
    g = np.load(os.path.join(util.GOLDEN, fn))
    G = util.build_G(res, 0, cb, cm).cuda()
    z = torch.from_numpy(g["z"]).cuda()
    img, ws = G(z, return_ws=True, noise_mode="const")
    np.testing.assert_allclose(ws.cpu().numpy(), g["ws"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(img.cpu().numpy(), g["img"], rtol=0, atol=1e-4)     # fp32 tolerance of the north_star


def test_ops_engine_attention_maps_and_signature():
    g = np.load(os.path.join(util.GOLDEN, "gen32_golden.npz"))
    G = util.build_G(32, 0, 512, 32).cuda()
    ws = torch.from_numpy(g["ws"]).cuda()
    img, att = G.synthesis(ws, pos=G.pos, mask=torch.ones(2, 16, device="cuda"), noise_mode="const")
    assert tuple(att.shape) == tuple(g["att_shape"])
    np.testing.assert_allclose(att[0, :, :, 0, ::8, ::8].cpu().numpy(), g["att_sample"], rtol=0, atol=1e-5)
    out = G(ws=ws, subnet="synthesis", noise_mode="const")
    assert torch.equal(out, img)


def test_ops_engine_grads_match_reference_golden():
    g = np.load(os.path.join(util.GOLDEN, "gen32_golden.npz"))
    G = util.build_G(32, 0, 512, 32).cuda()
    ws = torch.from_numpy(g["ws"]).cuda().requires_grad_(True)
    img, _ = G.synthesis(ws, pos=G.pos, mask=torch.ones(2, 16, device="cuda"), noise_mode="const", return_att_maps=False)
    gws, = torch.autograd.grad(img.square().mean(), [ws])
    scale = np.abs(g["gws"]).max()
    np.testing.assert_allclose(gws.cpu().numpy(), g["gws"], rtol=0, atol=1e-6 + 2e-4 * scale)
    z = torch.from_numpy(g["z"]).cuda().requires_grad_(True)
    gz, = torch.autograd.grad(G(z, noise_mode="const")[0].square().mean(), [z])
    np.testing.assert_allclose(gz.cpu().numpy(), g["gz"], rtol=0, atol=1e-7 + 2e-4 * np.abs(g["gz"]).max())


def test_nonfused_modconv_equals_fused():
    G = util.build_G(32, 0, 512, 32).cuda()
    ws = torch.randn(2, 17, G.num_ws, 32, device="cuda")
    m = torch.ones(2, 16, device="cuda")
    a, _ = G.synthesis(ws, pos=G.pos, mask=m, noise_mode="const", fused_modconv=True, return_att_maps=False)
    b, _ = G.synthesis(ws, pos=G.pos, mask=m, noise_mode="const", fused_modconv=False, return_att_maps=False)
    assert (a - b).abs().max() < 1e-4

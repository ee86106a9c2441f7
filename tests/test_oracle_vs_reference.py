"""Build-container only: pins the oracle restatements and the host mirror against the REAL reference code imported
read-only from /root/reference (skipped on the GPU box, where the tree does not exist)."""
import pytest
import torch
from oracle import refimport, ganformer, lpips_ref, ops as O
import util

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not present")


def test_mirror_init_bit_identical_to_reference():
    from morphganformer_b200.training import networks as N
    Gr = refimport.build_generator(32, seed=5, channel_base=512, channel_max=32)
    torch.manual_seed(5)
    Gm = N.Generator(**N.ganformer_default_kwargs(32, 512, 32))
    a, b = Gr.state_dict(), Gm.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_oracle_generator_matches_reference_fp32():
    Gr = util.randomize(refimport.build_generator(32, seed=2, channel_base=512, channel_max=32), 3)
    sd = {k: v.detach() for k, v in Gr.state_dict().items()}
    z = util.case_tensor((3, 17, 32), 9)
    img_r, ws_r = Gr(z, return_ws=True, noise_mode="const")
    img_o, ws_o = ganformer.generator(sd, z, 32)
    assert (ws_r - ws_o).abs().max() < 1e-6
    assert (img_r - img_o).abs().max() < 2e-5


def test_oracle_lpips_matches_reference():
    net = refimport.build_lpips(seed=4)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    a, b = torch.tanh(util.case_tensor((1, 3, 32, 32), 1)), torch.tanh(util.case_tensor((1, 3, 32, 32), 2))
    assert (net(a, b) - lpips_ref.lpips(sd, a, b)).abs().max() < 1e-6
    sd2 = util.build_vgg_lpips_sd(4)
    for k in sd2:
        assert torch.equal(sd2[k], sd[k]), k

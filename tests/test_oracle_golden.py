"""CPU: the oracle restatements (oracle/) against the golden vectors generated from the real reference
(tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os
import numpy as np
import pytest
import torch
from oracle import ops as O, ganformer, lpips_ref
import util

G = np.load(os.path.join(util.GOLDEN, "ops_golden.npz"))


@pytest.mark.parametrize("i", range(len(util.UPFIRDN_CASES)))
def test_upfirdn2d_oracle(i):
    name, xs, taps, kw = util.UPFIRDN_CASES[i]
    f = O.setup_filter(taps) if taps is not None else None
    y = O.upfirdn2d(util.case_tensor(xs, 100 + i), f, **kw)
    np.testing.assert_allclose(y.numpy(), G["upfirdn/" + name], rtol=0, atol=1e-6)


@pytest.mark.parametrize("act", util.BIAS_ACT_ACTS)
@pytest.mark.parametrize("clamp", [None, 0.4])
def test_bias_act_oracle(act, clamp):
    x = util.case_tensor((3, 5, 4, 6), 7).requires_grad_(True)
    b = util.case_tensor((5,), 8).requires_grad_(True)
    y = O.bias_act(x, b, dim=1, act=act, clamp=clamp)
    key = f"bias_act/{act}/{'clamp' if clamp else 'noclamp'}"
    np.testing.assert_allclose(y.detach().numpy(), G[key + "/y"], rtol=0, atol=1e-6)
    gx, gb = torch.autograd.grad((y * util.case_tensor(y.shape, 9)).sum(), [x, b])
    np.testing.assert_allclose(gx.numpy(), G[key + "/gx"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(gb.numpy(), G[key + "/gb"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("i", range(len(util.RESAMPLE_CASES)))
def test_conv2d_resample_oracle(i):
    name, xs, ws, kw = util.RESAMPLE_CASES[i]
    kw = dict(kw)
    f = kw.pop("f", None)
    f = O.setup_filter(f) if f is not None else None
    x = util.case_tensor(xs, 200 + i)
    w = util.case_tensor(ws, 300 + i) * 0.2
    y = O.conv2d_resample(x, w, f=f, **kw)
    np.testing.assert_allclose(y.numpy(), G[f"resample/{name}/y"], rtol=0, atol=2e-6)


def test_generator_init_matches_reference_checksum():
    g = np.load(os.path.join(util.GOLDEN, "gen32_golden.npz"))
    Gm = util.build_G(32, 0, 512, 32)
    assert abs(util.sd_checksum(util.state_dict_cpu(Gm)) - float(g["sd_checksum"])) < 1e-6 * abs(float(g["sd_checksum"]))


@pytest.mark.parametrize("cfg", [(32, 512, 32, "gen32_golden.npz"), (64, 32768, 512, "gen64_golden.npz"), (256, 32768, 512, "gen256_golden.npz")])
def test_generator_oracle_matches_golden(cfg):
    res, cb, cm, fn = cfg
    g = np.load(os.path.join(util.GOLDEN, fn))
    sd = util.state_dict_cpu(util.build_G(res, 0, cb, cm))
    assert abs(util.sd_checksum(sd) - float(g["sd_checksum"])) < 1e-6 * abs(float(g["sd_checksum"]))
    z = torch.from_numpy(g["z"])
    img, ws = ganformer.generator(sd, z, res)
    np.testing.assert_allclose(ws.numpy(), g["ws"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(img.numpy(), g["img"], rtol=0, atol=2e-5)


def test_generator_oracle_grads_match_golden():
    g = np.load(os.path.join(util.GOLDEN, "gen32_golden.npz"))
    sd = util.state_dict_cpu(util.build_G(32, 0, 512, 32))
    ws = torch.from_numpy(g["ws"]).requires_grad_(True)
    img = ganformer.synthesis(sd, ws, sd["pos"], torch.ones(2, 16), 32)
    gws, = torch.autograd.grad(img.square().mean(), [ws])
    np.testing.assert_allclose(gws.numpy(), g["gws"], rtol=0, atol=1e-6 + 1e-4 * np.abs(g["gws"]).max())


def test_generator_oracle_grads_match_golden_at_config1_size():
    """BASELINE configs[0] size (256^2, default widths, batch 1): the oracle's d(mean img^2)/d(ws) and d/dz against the real reference."""
    g = np.load(os.path.join(util.GOLDEN, "gen256_golden.npz"))
    sd = util.state_dict_cpu(util.build_G(256, 0))
    ws = torch.from_numpy(g["ws"]).requires_grad_(True)
    img = ganformer.synthesis(sd, ws, sd["pos"], torch.ones(1, 16), 256)
    gws, = torch.autograd.grad(img.square().mean(), [ws])
    np.testing.assert_allclose(gws.numpy(), g["gws"], rtol=0, atol=1e-6 + 1e-4 * np.abs(g["gws"]).max())
    z = torch.from_numpy(g["z"]).requires_grad_(True)
    gz, = torch.autograd.grad(ganformer.generator(sd, z, 256)[0].square().mean(), [z])
    np.testing.assert_allclose(gz.numpy(), g["gz"], rtol=0, atol=1e-6 + 1e-4 * np.abs(g["gz"]).max())


def test_lpips_oracle_matches_golden():
    g = np.load(os.path.join(util.GOLDEN, "lpips_golden.npz"))
    sd = util.build_vgg_lpips_sd(4)
    assert abs(float(sd["net.slice1.0.weight"].double().abs().sum()) - float(g["conv0_checksum"])) < 1e-6
    a = torch.tanh(util.case_tensor((2, 3, 64, 64), 60)).requires_grad_(True)
    b = torch.tanh(util.case_tensor((2, 3, 64, 64), 61))
    d = lpips_ref.lpips(sd, a, b)
    np.testing.assert_allclose(d.detach().numpy(), g["d"], rtol=1e-5, atol=1e-7)
    ga, = torch.autograd.grad(d.sum(), [a])
    np.testing.assert_allclose(ga.numpy(), g["ga"], rtol=0, atol=1e-7 + 1e-4 * np.abs(g["ga"]).max())

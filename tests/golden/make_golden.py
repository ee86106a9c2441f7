"""Generates the committed golden vectors from the REAL reference (read-only tree at /root/reference, build container
only).  Run:  python tests/golden/make_golden.py      (records torch version + thread count in every file)

  ops_golden.npz       reference `_upfirdn2d_ref`, `_bias_act_ref`, `conv2d_resample`, `fma` outputs for the case tables
                       in tests/util.py
  gen32_golden.npz     reference Generator (GANformer-default, res 32, channel_base 512 / max 32, seed 0, randomised
                       noise strengths/biases): state-dict checksum, ws, img, att map sample, d(mean img^2)/d ws, d/dz
  gen256_golden.npz    BASELINE configs[0]: res 256, default widths, batch 1: image, ws, d(ws), d(z)   (python make_golden.py --gen256)
  gen64_golden.npz     same at res 64 with the full 512-channel widths (checksum + outputs only; weights come from the seed)
  lpips_golden.npz     reference PNetLin (vgg, random torchvision init seed 4 + shipped lin weights) distances
  lpips_lin_vgg_v0.1.npz  the five 1x1 `lin` layers shipped in the reference (lpips/weights/v0.1/vgg.pth, 1472 floats)
"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refimport  # noqa: E402
import util  # noqa: E402

META = dict(torch_version=torch.__version__, threads=torch.get_num_threads())


def ops_golden(ref):
    out = {}
    for i, (name, xs, taps, kw) in enumerate(util.UPFIRDN_CASES):
        x = util.case_tensor(xs, 100 + i)
        f = ref.upfirdn2d.setup_filter(taps) if taps is not None else None
        out["upfirdn/" + name] = ref.upfirdn2d._upfirdn2d_ref(x, f, **kw).numpy()
    x = util.case_tensor((3, 5, 4, 6), 7)
    b = util.case_tensor((5,), 8)
    for act in util.BIAS_ACT_ACTS:
        for clamp in (None, 0.4):
            xx = x.clone().requires_grad_(True)
            bb = b.clone().requires_grad_(True)
            y = ref.bias_act._bias_act_ref(xx, bb, dim=1, act=act, clamp=clamp)
            gy = util.case_tensor(y.shape, 9)
            gx, gb = torch.autograd.grad((y * gy).sum(), [xx, bb], create_graph=True)
            ggx = None
            if gx.requires_grad:
                ggx, = torch.autograd.grad((gx * util.case_tensor(gx.shape, 10)).sum(), [xx], allow_unused=True)
            key = f"bias_act/{act}/{'clamp' if clamp else 'noclamp'}"
            out[key + "/y"] = y.detach().numpy(); out[key + "/gx"] = gx.detach().numpy(); out[key + "/gb"] = gb.detach().numpy()
            out[key + "/ggx"] = (ggx if ggx is not None else torch.zeros_like(xx)).detach().numpy()
    y = ref.bias_act._bias_act_ref(x.permute(0, 2, 3, 1), util.case_tensor((6,), 11), dim=2, act="lrelu", alpha=0.1, gain=0.7)
    out["bias_act/dim2_alpha_gain/y"] = y.numpy()
    for i, (name, xs, ws, kw) in enumerate(util.RESAMPLE_CASES):
        kw = dict(kw)
        f = kw.pop("f", None)
        f = ref.upfirdn2d.setup_filter(f) if f is not None else None
        x = util.case_tensor(xs, 200 + i).requires_grad_(True)
        w = (util.case_tensor(ws, 300 + i) * 0.2).requires_grad_(True)
        y = ref.conv2d_resample.conv2d_resample(x, w, f=f, **kw)
        gx, gw = torch.autograd.grad((y * util.case_tensor(y.shape, 400 + i)).sum(), [x, w])
        out[f"resample/{name}/y"] = y.detach().numpy(); out[f"resample/{name}/gx"] = gx.numpy(); out[f"resample/{name}/gw"] = gw.numpy()
    a, bq, c = util.case_tensor((2, 3, 4, 5), 20), util.case_tensor((2, 3, 1, 1), 21), util.case_tensor((4, 5), 22)
    out["fma/y"] = ref.fma.fma(a, bq, c).numpy()
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **out, **{"meta/" + k: np.array(str(v)) for k, v in META.items()})
    print("ops_golden:", len(out), "arrays")


def gen_golden(ref, res, cb, cm, fname, with_grads, batch=2):
    G = util.randomize(refimport.build_generator(res, seed=0, channel_base=cb, channel_max=cm), 1)
    sd = {k: v.detach() for k, v in G.state_dict().items()}
    z = util.case_tensor((batch, 17, 32), 50)
    out = dict(z=z.numpy(), sd_checksum=np.array(util.sd_checksum(sd)))
    zz = z.clone().requires_grad_(with_grads)
    ws = G.mapping(zz, None, pos=G.pos, mask=torch.ones(batch, 16))
    wsl = ws.detach().clone().requires_grad_(with_grads)
    img, att = G.synthesis(wsl, pos=G.pos, mask=torch.ones(batch, 16), noise_mode="const")
    out.update(ws=ws.detach().numpy(), img=img.detach().numpy(), att_shape=np.array(att.shape), att_sample=att.detach()[0, :, :, 0, ::8, ::8].numpy())
    if with_grads:
        loss = img.square().mean()
        gws, = torch.autograd.grad(loss, [wsl])
        img2 = G(zz, noise_mode="const")[0]
        gz, = torch.autograd.grad(img2.square().mean(), [zz])
        out.update(gws=gws.numpy(), gz=gz.numpy(), loss=np.array(loss.item()))
    np.savez_compressed(os.path.join(HERE, fname), **out, **{"meta/" + k: np.array(str(v)) for k, v in META.items()})
    print(fname, "img absmax", float(img.abs().max()))


def lpips_golden(ref):
    net = refimport.build_lpips(seed=4)
    sd = net.state_dict()
    np.savez(os.path.join(HERE, "lpips_lin_vgg_v0.1.npz"), **{f"lin{k}": sd[f"lin{k}.model.1.weight"].flatten().numpy() for k in range(5)})
    a = torch.tanh(util.case_tensor((2, 3, 64, 64), 60)).requires_grad_(True)
    b = torch.tanh(util.case_tensor((2, 3, 64, 64), 61))
    d = net(a, b)
    ga, = torch.autograd.grad(d.sum(), [a])
    np.savez_compressed(os.path.join(HERE, "lpips_golden.npz"), d=d.detach().numpy(), ga=ga.numpy(),
                        conv0_checksum=np.array(float(sd["net.slice1.0.weight"].double().abs().sum())),
                        **{"meta/" + k: np.array(str(v)) for k, v in META.items()})
    print("lpips d", d.flatten().tolist())


if __name__ == "__main__":
    ref = refimport.load()
    if "--gen256" in sys.argv:      # BASELINE.json configs[0]: generator forward at 256x256, batch 1, default widths (+ d(ws) of mean(img^2))
        gen_golden(ref, 256, 32768, 512, "gen256_golden.npz", True, batch=1)
        sys.exit(0)
    ops_golden(ref)
    gen_golden(ref, 32, 512, 32, "gen32_golden.npz", True)
    gen_golden(ref, 64, 32768, 512, "gen64_golden.npz", False)
    lpips_golden(ref)

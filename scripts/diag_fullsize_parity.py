"""Diagnostic (GPU box): tc engine + LpipsEngine against the CPU oracle at the benchmarked sizes.
Prints absolute image error, d(ws) relative L2, per-image projection loss / LPIPS relative error.
    python scripts/diag_fullsize_parity.py [256 1024]"""
import os
import sys
import time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from oracle import ganformer, lpips_ref
from morphganformer_b200 import _lib
from morphganformer_b200.lpips_engine import LpipsEngine

torch.set_num_threads(os.cpu_count())


def run(res, B, fwd):
    G = util.build_G(res, 0)
    sd = util.state_dict_cpu(G)
    lsd = util.build_vgg_lpips_sd(4)
    z = util.case_tensor((B, 17, 32), 31)
    tgt = torch.tanh(util.case_tensor((B, 3, res, res), 32))
    t0 = time.time()
    zr = z.clone().requires_grad_(True)
    img_ref, ws_ref = ganformer.generator(sd, zr, res)
    ws_ref.retain_grad()
    mse = (img_ref - tgt).square().mean(dim=[1, 2, 3])
    lp = lpips_ref.lpips(lsd, img_ref, tgt).reshape(B)
    loss = 0.5 * lp + 0.5 * mse
    loss.sum().backward()
    gws_ref = ws_ref.grad.detach()
    print("[%d B=%d] oracle fwd+bwd %.1f s; img range %.3f" % (res, B, time.time() - t0, img_ref.abs().max().item()), flush=True)
    _lib.set_forward_dtype(fwd)
    Gc = G.cuda(); Gc.synthesis.engine = "tc"
    w = ws_ref.detach().cuda().requires_grad_(True)
    img, _ = Gc.synthesis(w, pos=Gc.pos, mask=torch.ones(B, 16, device="cuda"), noise_mode="const")
    eng = LpipsEngine(lsd)
    eng.set_target(tgt.cuda())
    val, mse_sum = eng.forward(img.detach())
    n = 3 * res * res
    per = 0.5 * val + 0.5 * mse_sum / n
    dimg = eng.backward(torch.full((B,), 0.5, device="cuda"), 0.5 * 2.0 / n)
    img.backward(dimg)
    e = (img.detach().cpu() - img_ref.detach())
    g = w.grad.cpu()
    print("  %s: img max-abs %.4g  rel-rms %.3g | lpips rel %s | mse rel %s | loss rel %s | dws relL2 %.4g cos %.6f" % (
        fwd, e.abs().max().item(), (e.square().mean().sqrt() / img_ref.square().mean().sqrt()).item(),
        ((val.cpu() - lp.detach()).abs() / lp.detach().abs()).tolist(), ((mse_sum.cpu() / n - mse.detach()).abs() / mse.detach()).tolist(),
        ((per.cpu() - loss.detach()).abs() / loss.detach().abs()).tolist(),
        ((g - gws_ref).norm() / gws_ref.norm()).item(), torch.nn.functional.cosine_similarity(g.flatten(), gws_ref.flatten(), dim=0).item()), flush=True)
    # LPIPS engine alone on the oracle's image (isolates the VGG path)
    val2, _ = eng.forward(img_ref.detach().cuda())
    print("     lpips engine on the oracle image: rel %s" % (((val2.cpu() - lp.detach()).abs() / lp.detach().abs()).tolist(),), flush=True)
    _lib.set_forward_dtype("bf16")


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1024]
    for r in sizes:
        for fwd in ("fp16", "bf16"):
            run(r, 2 if r <= 256 else 1, fwd)

"""Build-container only: golden vectors for LPIPS with the AlexNet / SqueezeNet backbones from the REAL reference
(lpips/networks_basic.PNetLin, pnet_rand=True under seed 4 + the shipped lin weights lpips/weights/v0.1/{alex,squeeze}.pth).
Writes lpips_lin_{alex,squeeze}_v0.1.npz (the shipped 1x1 `lin` weights, 1152 / 2368 floats) and lpips_nets_golden.npz
(distances and d/d(pred) for fixed inputs).      python tests/golden/make_golden_lpips_nets.py"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np, torch
from oracle import refimport
import util

out = {}
for nt in ("alex", "squeeze"):
    net = refimport.build_lpips(seed=4, net_type=nt)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    L = 5 if nt == "alex" else 7
    np.savez(os.path.join(HERE, f"lpips_lin_{nt}_v0.1.npz"), **{f"lin{k}": sd[f"lin{k}.model.1.weight"].flatten().numpy() for k in range(L)})
    a = torch.tanh(util.case_tensor((2, 3, 96, 96), 31)).requires_grad_(True)
    b = torch.tanh(util.case_tensor((2, 3, 96, 96), 32))
    d = net(a, b)
    ga, = torch.autograd.grad(d.sum(), [a])
    out[nt + "_d"] = d.detach().numpy(); out[nt + "_ga"] = ga.numpy()
    print(nt, d.flatten().tolist())
np.savez_compressed(os.path.join(HERE, "lpips_nets_golden.npz"), torch_version=torch.__version__, threads=torch.get_num_threads(), **out)

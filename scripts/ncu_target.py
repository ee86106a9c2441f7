"""Workload for ncu captures: two eager projection steps (8 images, 1024^2) -- the second one is what the --launch-skip counts aim at.
    ncu --set full --clock-control none --import-source on -k regex:<kernels> --launch-skip N -c M -o gpurun_out/x python scripts/ncu_target.py"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from morphganformer_b200.projection import Projector, latent_stats
B, R = int(os.environ.get("NCU_B", "8")), int(os.environ.get("NCU_R", "1024"))
G = util.build_G(R, 0).cuda()
mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
P = Projector(G, util.build_vgg_lpips_sd(4), B, 100, latent_mean=mean, latent_std=std, step_noise=torch.zeros(100, B, 17, 32))
P.set_targets(torch.tanh(torch.randn(B, 3, R, R)).cuda())
for _ in range(int(os.environ.get("NCU_STEPS", "2"))):
    P.step(use_graph=False)
torch.cuda.synchronize()
print("done")

"""Kernel-level time table of one eager projection step (torch.profiler / CUPTI; no ncu needed), optionally for several halo-kernel modes.
    python scripts/prof_step.py [--batch 8] [--res 1024] [--halo-modes 1,5,0] [--gen-only]
Prints, per mode: total GPU kernel time of one step and the per-kernel table (name, launches, ms)."""
import argparse
import os
import sys
import collections
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from morphganformer_b200 import _lib, tc
from morphganformer_b200.projection import Projector, latent_stats

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--halo-modes", default="1")
ap.add_argument("--dump-conv", action="store_true")
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--detail", default="", help="substring: also list every launch of the kernels whose name contains it, in launch order")
args = ap.parse_args()

B, R = args.batch, args.res
G = util.build_G(R, 0).cuda()
lsd = util.build_vgg_lpips_sd(4)
mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
P = Projector(G, lsd, B, 100, latent_mean=mean, latent_std=std, step_noise=torch.zeros(100, B, 17, 32))
P.set_targets(torch.tanh(torch.randn(B, 3, R, R)).cuda())
for _ in range(3):
    P.step(use_graph=False)
torch.cuda.synchronize()


def short(name):
    name = name.replace("mgf::", "").replace("(anonymous namespace)::", "")
    i = name.find("(")
    name = name[:i] if i > 0 else name
    return name.replace("void ", "")[:70]


for mode in [int(m) for m in args.halo_modes.split(",")]:
    _lib.lib().mgf_conv_tc_set_halo(mode)
    for _ in range(2):
        P.step(use_graph=False)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        P.step(use_graph=False)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            d = agg.setdefault(short(ev.name), [0, 0.0])
            d[0] += 1; d[1] += ev.device_time / 1000.0 if hasattr(ev, "device_time") else ev.cuda_time / 1000.0
    if args.detail:
        for ev in sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and args.detail in e.name), key=lambda e: e.time_range.start):
            print("   launch %-60s %8.1f us" % (short(ev.name), ev.device_time if hasattr(ev, "device_time") else ev.cuda_time))
    tot = sum(v[1] for v in agg.values())
    print("=== halo mode %d: %d kernels, %.3f ms summed kernel time" % (mode, sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print("  %-70s %4d  %8.3f ms  %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    if args.dump_conv:
        tc.PROFILE = []
        P.step(use_graph=False)
        torch.cuda.synchronize()
        recs, tc.PROFILE = tc.PROFILE, None
        for rec in recs:
            print("CONV %-8s %.3f ms  %7.1f alg TF/s  %s" % (rec[0], rec[1].elapsed_time(rec[2]), rec[3] / rec[1].elapsed_time(rec[2]) / 1e9, rec[5]))
_lib.lib().mgf_conv_tc_set_halo(1)

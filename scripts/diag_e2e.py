"""Where does the end-to-end projection loop lose time against the device-resident loop?  Times each phase with a host clock
around explicit synchronisation (diagnostic, not a bench value)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util
from morphganformer_b200 import _lib
from morphganformer_b200.projection import Projector, latent_stats

dev = torch.device("cuda", 0)
B, R, K = 8, 1024, 10
_lib.set_forward_dtype("fp16")
G = util.build_G(R, 0).to(dev)
lsd = util.build_vgg_lpips_sd(4)
mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
gen = torch.Generator(device="cpu").manual_seed(1000)
target_host = torch.tanh(torch.randn(B, 3, R, R, generator=gen)).pin_memory()
noise_host = torch.randn(64, B, 17, 32, generator=gen).pin_memory()
P = Projector(G, lsd, B, 1000, latent_mean=mean, latent_std=std, step_noise=torch.zeros(1000, B, 17, 32))
P.set_targets(target_host.to(dev))
P.capture()

def T(name, fn, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3 / n
    print("%-46s %8.2f ms" % (name, dt), flush=True)

for _ in range(3):
    P.step()
T("graph step x10 back-to-back (per step)", P.step, 10)
def step_sync():
    P.step(); torch.cuda.current_stream().synchronize()
T("graph step + stream sync (per step)", step_sync, 10)
loss_host = torch.empty(B).pin_memory()
def step_full():
    P.step_noise[P.i + 1].copy_(noise_host[P.i % 64], non_blocking=True)
    per = P.step(); loss_host.copy_(per, non_blocking=True); torch.cuda.current_stream().synchronize()
T("noise h2d + graph step + loss d2h + sync", step_full, 10)
T("P.reset()", P.reset)
T("target h2d (100 MB pinned)", lambda: target_host.to(dev, non_blocking=True))
tdev = target_host.to(dev)
T("P.set_targets(device tensor) 1st", lambda: P.set_targets(tdev))
T("P.set_targets(device tensor) 2nd", lambda: P.set_targets(tdev))
T("graph step after set_targets (per step)", P.step, 10)
T("eager step (per step)", lambda: P.step(use_graph=False), 5)
print("mem allocated %.1f GB reserved %.1f GB" % (torch.cuda.memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30))

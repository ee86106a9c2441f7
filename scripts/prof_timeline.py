"""Timeline of ONE CUDA-graph replay of the projection step (torch.profiler / CUPTI kernel records; no ncu): where the step's
wall time goes beyond the summed kernel time.
    python scripts/prof_timeline.py [--batch 8] [--res 1024] [--gaps 30]
Prints the span of the replay, busy time per stream, the time covered by at least one kernel, and the largest idle gaps with the
kernels on either side -- launch gaps between dependent graph nodes and waits of the main stream on the side stream show up here,
not in the per-kernel tables of prof_step.py."""
import argparse
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from morphganformer_b200.projection import Projector, latent_stats

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--gaps", type=int, default=30)
ap.add_argument("--dump", action="store_true", help="print every kernel record (start, duration, stream, name)")
args = ap.parse_args()

B, R = args.batch, args.res
G = util.build_G(R, 0).cuda()
lsd = util.build_vgg_lpips_sd(4)
mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
P = Projector(G, lsd, B, 100, latent_mean=mean, latent_std=std, step_noise=torch.zeros(100, B, 17, 32))
P.set_targets(torch.tanh(torch.randn(B, 3, R, R)).cuda())
P.capture()
for _ in range(5):
    P.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    P.step()
e1.record(); torch.cuda.synchronize()
print("graph replay: %.3f ms / step (10 replays, CUDA events)" % (e0.elapsed_time(e1) / 10))

with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    P.step()
    torch.cuda.synchronize()


def short(name):
    name = name.replace("mgf::", "").replace("(anonymous namespace)::", "")
    i = name.find("(")
    name = name[:i] if i > 0 else name
    return name.replace("void ", "")[:60]


evs = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        evs.append((ev.time_range.start, ev.time_range.start + dur, getattr(ev, "device_resource_id", getattr(ev, "stream", -1)), short(ev.name)))
evs.sort()
t0, t1 = evs[0][0], max(e[1] for e in evs)
print("kernel records: %d, span %.3f ms" % (len(evs), (t1 - t0) / 1000.0))
streams = {}
for s, e, st, n in evs:
    d = streams.setdefault(st, [0, 0.0]); d[0] += 1; d[1] += e - s
for st, d in sorted(streams.items(), key=lambda kv: -kv[1][1]):
    print("  stream %-6s %4d kernels  %8.3f ms busy" % (st, d[0], d[1] / 1000.0))
# union coverage + idle gaps (no kernel of any stream running)
covered, cur_end, gaps, last_name = 0.0, t0, [], "(start)"
ends = []
for s, e, st, n in evs:
    if s > cur_end:
        gaps.append((s - cur_end, last_name, n, (cur_end - t0) / 1000.0))
        covered += 0
        cur_start = s
    covered += max(0.0, e - max(s, cur_end))
    if e > cur_end:
        cur_end, last_name = e, n
print("covered by >= 1 kernel: %.3f ms; idle inside the span: %.3f ms in %d gaps" % (covered / 1000.0, (t1 - t0 - covered) / 1000.0, len(gaps)))
hist = {}
for g in gaps:
    k = "<2us" if g[0] < 2 else "<4us" if g[0] < 4 else "<8us" if g[0] < 8 else "<16us" if g[0] < 16 else ">=16us"
    d = hist.setdefault(k, [0, 0.0]); d[0] += 1; d[1] += g[0]
for k in ("<2us", "<4us", "<8us", "<16us", ">=16us"):
    if k in hist:
        print("  gaps %-6s %4d  %8.3f ms" % (k, hist[k][0], hist[k][1] / 1000.0))
print("largest gaps (us, at ms, after -> before):")
for g in sorted(gaps, reverse=True)[:args.gaps]:
    print("  %7.1f us @ %7.3f ms  %-50s -> %s" % (g[0], g[3], g[1], g[2]))
# main stream = the stream with the most busy time: kernels on it that START while nothing else of it runs but later than the previous end
main = max(streams.items(), key=lambda kv: kv[1][1])[0]
prev_end, main_gap = None, 0.0
for s, e, st, n in evs:
    if st != main:
        continue
    if prev_end is not None and s > prev_end:
        main_gap += s - prev_end
    prev_end = e if prev_end is None else max(prev_end, e)
print("main stream %s: idle between its own kernels %.3f ms" % (main, main_gap / 1000.0))
if args.dump:
    for s, e, st, n in evs:
        print("K %9.1f %8.1f %s %s" % (s - t0, e - s, st, n))

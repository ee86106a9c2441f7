"""Stage-isolation experiments on the halo convolution kernel (64-channel 3x3 layers at 512^2 / 1024^2): the same launch timed with the
epilogue reduced to a TMEM drain (dbg 1), without MMAs (dbg 2), without activation loads (dbg 4), to see which stage bounds the tile rate.
    python scripts/bench_halo.py [--stages] [names...]       (without --stages: only the CTA-pair vs one-CTA comparison)"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morphganformer_b200 import tc, _lib

L = _lib.lib()
B = 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fdt = _lib.forward_torch_dtype()


def time_it(fn):
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return min(ts)


def cases():
    # generator 1024^2 32 -> 32 layer as super-pixel rows [B,1024,512,64], per-sample weights
    H, W = 1024, 512
    x = torch.randn(B, H, W, 64, device="cuda").to(fdt); out = torch.empty_like(x)
    wt = (torch.randn(B, 9, 64, 64, device="cuda") * 0.05).to(fdt)
    noise = torch.randn(H, 2 * W, device="cuda"); nstr = torch.tensor([0.1], device="cuda"); bias = torch.randn(64, device="cuda") * 0.1
    yield "g1024.fwd.plain", lambda: tc.conv_tc([x], wt, tc.TAPS_3X3, (B, H, W), 1, 64, out, tag="x")
    yield "g1024.fwd.tail", lambda: tc.conv_tc([x], wt, tc.TAPS_3X3, (B, H, W), 1, 64, out, noise=noise, noise_strength=nstr, bias=bias, act=1,
                                               gain=math.sqrt(2.0), superpix=True, tag="x")
    dy = torch.randn(B, H, W, 64, device="cuda").to(torch.bfloat16); dx = torch.empty_like(dy)
    wb = (torch.randn(B, 9, 64, 64, device="cuda") * 0.05).to(torch.bfloat16)
    s2 = torch.rand(B, 64, device="cuda") + 0.5; ds2 = torch.zeros(B, 64, device="cuda")
    yield "g1024.bwd.red+X", lambda: tc.conv_tc([dy], wb, tc.TAPS_3X3, (B, H, W), 1, 64, dx, scale_n=s2, reduce_out=ds2, X=x, reduce_per_sample=True, fwd=False, tag="x")
    yield "g1024.bwd.red+X+ag", lambda: tc.conv_tc([dy], wb, tc.TAPS_3X3, (B, H, W), 1, 64, dx, scale_n=s2, reduce_out=ds2, X=x, actgrad=True, ag_gain=1.4,
                                                   reduce_per_sample=True, fwd=False, tag="x")
    # VGG conv1_2: 1024^2 64 -> 64, shared weights, bias + ReLU; backward with the ReLU mask of the saved activation
    H, W = 1024, 1024
    xv = torch.randn(B, H, W, 64, device="cuda").to(fdt); ov = torch.empty_like(xv)
    wv = (torch.randn(1, 9, 64, 64, device="cuda") * 0.05).to(fdt)
    yield "vgg1_2.fwd", lambda: tc.conv_tc([xv], wv, tc.TAPS_3X3, (B, H, W), 1, 64, ov, bias=bias, act=2, tag="x")
    gv = torch.randn(B, H, W, 64, device="cuda").to(torch.bfloat16); dv = torch.empty_like(gv)
    wvb = (torch.randn(1, 9, 64, 64, device="cuda") * 0.05).to(torch.bfloat16)
    yield "vgg1_2.bwd.X", lambda: tc.conv_tc([gv], wvb, tc.TAPS_3X3, (B, H, W), 1, 64, dv, X=xv, actgrad=True, ag_alpha=0.0, fwd=False, tag="x")


want = [a for a in sys.argv[1:] if not a.startswith("--")]
for name, fn in cases():
    if want and not any(w in name for w in want):
        continue
    L.mgf_conv_tc_set_halo(1 | 64); t_pair = time_it(fn)
    L.mgf_conv_tc_set_halo(1); t_single = time_it(fn)
    L.mgf_conv_tc_set_halo(1)
    print("%-20s CTA-pair kernel %.3f ms, one-CTA kernel %.3f ms" % (name, t_pair, t_single), flush=True)
    if "--stages" not in sys.argv:
        continue
    for groups in (3, 2):
        row = []
        for dbg in (0, 2, 5, 6):
            L.mgf_conv_tc_set_halo(1 | (32 if groups == 2 else 0) | (dbg << 8))
            row.append("dbg%d %.3f" % (dbg, time_it(fn)))
        L.mgf_conv_tc_set_halo(1)
        print("%-20s groups%d %s ms" % (name, groups, "  ".join(row)), flush=True)

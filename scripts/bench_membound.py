"""Micro-benchmark of the HBM-bound engine kernels at the 1024^2 x 8 bench shapes: achieved algorithmic GB/s (CUDA events, best of 5)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morphganformer_b200 import _lib
_lib.set_forward_dtype("fp16")      # the bench mode (16-bit payloads: only the timing matters here)
L = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
def p(t): return t.data_ptr() if t is not None else None
def run(name, fn, nbytes):
    ts = []
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = fn(); e1.record(); torch.cuda.synchronize()
        if rc:
            print("%-34s FAILED rc=%d %s" % (name, rc, L.mgf_last_error()), flush=True)
            return
        if i >= 2: ts.append(e0.elapsed_time(e1))
    print("%-34s %.3f ms  %6.0f GB/s" % (name, min(ts), nbytes / min(ts) / 1e6), flush=True)
B = 8
bf = lambda *sh: torch.randn(*sh, device="cuda").to(torch.bfloat16)
for (H, C) in [(1024, 64), (512, 128), (256, 256), (128, 512)]:
    f = bf(B, H, H, C).abs(); n1 = bf(B, H, H, C).abs() * 0.1; lin = torch.rand(C, device="cuda"); val = torch.zeros(B, device="cuda"); coef = torch.full((B,), 0.5, device="cuda")
    out = torch.empty_like(f); n = f.numel() * 2
    run("lpips_head fwd %dx%d C%d" % (H, H, C), lambda: L.mgf_lpips_head(1, p(f), p(n1), p(lin), None, None, p(val), 0, B, H * H, C, s), 2 * n)
    dy = bf(B, H // 2, H // 2, C)
    run("lpips_tap_pool_bwd %dx%d C%d" % (H, H, C), lambda: L.mgf_lpips_tap_pool_bwd(p(f), p(n1), p(lin), p(coef), p(dy), p(out), None, B, H, H, C, s), 3.25 * n)
    y = torch.empty(B, H // 2, H // 2, C, dtype=torch.bfloat16, device="cuda")
    run("maxpool2_fwd %dx%d C%d" % (H, H, C), lambda: L.mgf_maxpool2_fwd(p(f), p(y), B, H, H, C, s), 1.25 * n)
for (H, C) in [(1024, 32), (512, 64), (256, 128)]:
    z = bf(B, H, H, C); dz = bf(B, H, H, C); dy = torch.empty_like(z); R = torch.zeros(B, C, device="cuda"); noise = torch.randn(H, H, device="cuda"); ns = torch.tensor([0.1], device="cuda"); bias = torch.zeros(C, device="cuda")
    n = z.numel() * 2
    run("act_bwd mode0 %dx%d C%d" % (H, H, C), lambda: L.mgf_act_bwd(p(dz), p(z), p(dy), p(R), p(noise), p(ns), p(bias), 0.2, 1.0, 0, B, H * H, C, 0, s), 3 * n)
    run("act_bwd reduce-only %dx%d C%d" % (H, H, C), lambda: L.mgf_act_bwd(p(dz), p(z), None, p(R), p(noise), p(ns), p(bias), 0.2, 1.4, 1, B, H * H, C, 0, s), 2 * n)
    v = bf(B, H // 2, H // 2, C); fk = (ctypes.c_float * 4)(0.125, 0.375, 0.375, 0.125)
    run("upfir2_add -> %dx%d C%d" % (H, H, C), lambda: L.mgf_upfir2_add(p(v), p(z), p(dy), fk, 2.8, B, H // 2, H // 2, C, s), 2.25 * n)
    dv = torch.empty_like(v)
    run("upfir2_bwd %dx%d C%d" % (H, H, C), lambda: L.mgf_upfir2_bwd(p(dz), p(dv), fk, 2.8, B, H // 2, H // 2, C, s), 1.25 * n)
R_ = 1024
img = torch.randn(B, 3, R_, R_, device="cuda"); tgt = torch.randn_like(img); col = torch.empty(B, R_, R_, 32, dtype=torch.bfloat16, device="cuda"); mse = torch.zeros(B, device="cuda"); dimg = torch.empty_like(img)
run("lpips_prep 1024", lambda: L.mgf_lpips_prep(p(img), p(tgt), p(col), p(mse), B, R_, s), img.numel() * 8 + col.numel() * 2)
run("lpips_prep_bwd 1024", lambda: L.mgf_lpips_prep_bwd(p(col), p(img), p(tgt), 0.1, p(dimg), B, R_, s), img.numel() * 12 + col.numel() * 2)
y = bf(B, R_, R_, 32); wr = torch.randn(3, 32, device="cuda"); sr = torch.randn(B, 32, device="cuda"); br = torch.zeros(3, device="cuda")
run("torgb_fwd 1024 C32", lambda: L.mgf_torgb_fwd(p(y), p(wr), p(sr), p(br), p(img), B, R_ * R_, 32, s), y.numel() * 2 + img.numel() * 4)
dyy = torch.empty_like(y); ds = torch.zeros(B, 32, device="cuda"); RR = torch.zeros(B, 32, device="cuda")
run("torgb_bwd 1024 C32", lambda: L.mgf_torgb_bwd(p(dimg), p(y), p(wr), p(sr), p(dyy), p(ds), p(RR), B, R_ * R_, 32, s), y.numel() * 4 + img.numel() * 4)

# fused LPIPS input stage + VGG conv1_1 (vgg_first.cu): bytes = image (+ target) + the 64-channel tensor once (+ image gradient)
wc = torch.randn(64, 32, device="cuda") * 0.1; b64 = torch.zeros(64, device="cuda"); h0 = torch.empty(B, R_, R_, 64, dtype=torch.bfloat16, device="cuda")
run("vgg_conv1_fwd 1024 (+mse)", lambda: L.mgf_vgg_conv1_fwd(p(img), p(tgt), p(mse), p(wc), p(b64), p(h0), B, R_, s), img.numel() * 8 + h0.numel() * 2)
run("vgg_conv1_fwd 1024 (no mse)", lambda: L.mgf_vgg_conv1_fwd(p(img), None, None, p(wc), p(b64), p(h0), B, R_, s), img.numel() * 4 + h0.numel() * 2)
run("vgg_conv1_bwd 1024 (+mse grad)", lambda: L.mgf_vgg_conv1_bwd(p(h0), p(wc), p(img), p(tgt), 0.1, p(dimg), B, R_, s), img.numel() * 12 + h0.numel() * 2)

# LPIPS tap + 2x2 max-pool forward, then the backward fed with the forward's per-pixel reductions (as lpips_engine.py runs them):
# forward bytes = x + n1 read, pooled y written (2.25 x tensor bytes) + 8 B / pixel of stats; backward = x, n1, dx + dy/4 + stats (3.25 x)
for (H, C) in [(1024, 64), (512, 128), (256, 256), (128, 512)]:
    f = bf(B, H, H, C).abs().to(torch.float16).view(torch.bfloat16); n1 = (bf(B, H, H, C).abs() * 0.1).to(torch.float16).view(torch.bfloat16)
    lin = torch.rand(C, device="cuda"); val = torch.zeros(B, device="cuda"); coef = torch.full((B,), 0.5, device="cuda")
    y = torch.empty(B, H // 2, H // 2, C, dtype=torch.bfloat16, device="cuda"); stats = torch.empty(B, H, H, 2, device="cuda")
    dy = bf(B, H // 2, H // 2, C); dx = torch.empty(B, H, H, C, dtype=torch.bfloat16, device="cuda"); n = f.numel() * 2
    run("lpips_tap_pool_fwd %dx%d C%d" % (H, H, C), lambda: L.mgf_lpips_tap_pool_fwd(p(f), p(n1), p(lin), p(y), p(val), p(stats), B, H, H, C, s), 2.25 * n + stats.numel() * 4)
    run("lpips_tap_pool_bwd+stats %dx%d C%d" % (H, H, C), lambda: L.mgf_lpips_tap_pool_bwd(p(f), p(n1), p(lin), p(coef), p(dy), p(dx), p(stats), B, H, H, C, s), 3.25 * n + stats.numel() * 4)
# first stage of the up-convolution input gradient: dy [B,H,W,C] -> g [B,H+2,W+2,C] (bf16), bytes = both tensors once
fk = (ctypes.c_float * 4)(0.125, 0.375, 0.375, 0.125)
for (H, C) in [(1024, 32), (512, 64), (256, 128), (128, 256)]:
    dy = bf(B, H, H, C); gq = torch.empty(B, H + 2, H + 2, C, dtype=torch.bfloat16, device="cuda")
    run("fir4_pad %dx%d C%d" % (H, H, C), lambda: L.mgf_fir4_pad(p(dy), p(gq), fk, 4.0, B, H, H, C, s), dy.numel() * 2 + gq.numel() * 2)
# resnet-skip 1x1 convolution (pointwise.cu) and its input gradient: bytes = input + output
for (H, K, N) in [(512, 64, 32), (256, 128, 64), (128, 256, 128)]:
    x = bf(B, H, H, K); w = bf(N, K) * 0.1; o = torch.empty(B, H, H, N, dtype=torch.bfloat16, device="cuda")
    run("pointwise fwd %dx%d %d->%d" % (H, H, K, N), lambda: L.mgf_pointwise(p(x), p(w), p(o), B * H * H, K, N, 0, s), x.numel() * 2 + o.numel() * 2)
    wt = bf(K, N) * 0.1
    run("pointwise dgrad %dx%d %d->%d" % (H, H, N, K), lambda: L.mgf_pointwise(p(o), p(wt), p(x), B * H * H, N, K, 0, s), x.numel() * 2 + o.numel() * 2)

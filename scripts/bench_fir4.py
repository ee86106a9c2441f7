"""Micro-benchmark of mgf_fir4_pad (first stage of the up-convolution's input gradient) + check against a PyTorch statement."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morphganformer_b200 import _lib
L = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
fk = (ctypes.c_float * 4)(0.125, 0.375, 0.375, 0.125)
f2 = torch.tensor([0.125, 0.375, 0.375, 0.125], device="cuda")
x = torch.randn(2, 10, 14, 32, device="cuda").to(torch.bfloat16); g = torch.empty(2, 12, 16, 32, device="cuda", dtype=torch.bfloat16)
_lib.check(L.mgf_fir4_pad(x.data_ptr(), g.data_ptr(), fk, 4.0, 2, 10, 14, 32, s))
xp = torch.nn.functional.pad(x.float().permute(0, 3, 1, 2), (2, 2, 2, 2))
ker = (torch.outer(f2, f2) * 4.0).flip(0, 1).reshape(1, 1, 4, 4).repeat(32, 1, 1, 1)
ref = torch.nn.functional.conv2d(xp, ker, groups=32)                      # [2,32,11,15]: g[u,v] = sum Ff[a,b] dy[u-a+1, v-b+1]
err = (g[:, :11, :15].float().permute(0, 3, 1, 2) - ref).abs().max().item()
print("fir4_pad max err vs torch %.3g (scale %.3g); border zero: %s" % (err, ref.abs().max().item(), bool((g[:, 11] == 0).all() and (g[:, :, 15] == 0).all())))
for (H, C) in [(1024, 32), (512, 64), (256, 128), (128, 256)]:
    dy = torch.randn(8, H, H, C, device="cuda").to(torch.bfloat16); gq = torch.empty(8, H + 2, H + 2, C, device="cuda", dtype=torch.bfloat16)
    ts = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); L.mgf_fir4_pad(dy.data_ptr(), gq.data_ptr(), fk, 4.0, 8, H, H, C, s); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("fir4_pad %d C%d: %.3f ms  %.0f GB/s" % (H, C, min(ts[2:]), (dy.numel() + gq.numel()) * 2 / min(ts[2:]) / 1e6))

"""A/B of the FIR walkers with 8 and 4 channels per thread (mgf_fir_set_mode) at the bench shapes: CUDA events, best of 5, algorithmic GB/s."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morphganformer_b200 import _lib
_lib.set_forward_dtype("fp16")
L = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
def p(t): return t.data_ptr() if t is not None else None
def run(name, fn, nbytes):
    ts = []
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = fn(); e1.record(); torch.cuda.synchronize()
        if rc:
            print("%-34s FAILED rc=%d %s" % (name, rc, L.mgf_last_error()), flush=True); return
        if i >= 2: ts.append(e0.elapsed_time(e1))
    print("%-34s %.3f ms  %6.0f GB/s" % (name, min(ts), nbytes / min(ts) / 1e6), flush=True)
B = 8
bf = lambda *sh: torch.randn(*sh, device="cuda").to(torch.bfloat16)
fk = (ctypes.c_float * 4)(0.125, 0.375, 0.375, 0.125)
for mode in (0, 7):
    L.mgf_fir_set_mode(mode)
    print("=== mgf_fir_set_mode(%d): %d channels per thread" % (mode, 4 if mode else 8), flush=True)
    for (H, C) in [(1024, 32), (512, 64), (256, 128), (128, 256), (64, 512)]:
        dy = bf(B, H, H, C); gq = torch.empty(B, H + 2, H + 2, C, dtype=torch.bfloat16, device="cuda")
        run("fir4_pad %dx%d C%d" % (H, H, C), lambda: L.mgf_fir4_pad(p(dy), p(gq), fk, 4.0, B, H, H, C, s), dy.numel() * 2 + gq.numel() * 2)
        z = bf(B, H, H, C); out = torch.empty_like(z); v = bf(B, H // 2, H // 2, C); n = z.numel() * 2
        run("upfir2_add -> %dx%d C%d" % (H, H, C), lambda: L.mgf_upfir2_add(p(v), p(z), p(out), fk, 2.8, B, H // 2, H // 2, C, s), 2.25 * n)
        dv = torch.empty_like(v)
        run("upfir2_bwd %dx%d C%d" % (H, H, C), lambda: L.mgf_upfir2_bwd(p(z), p(dv), fk, 2.8, B, H // 2, H // 2, C, s), 1.25 * n)
L.mgf_fir_set_mode(2)      # library default

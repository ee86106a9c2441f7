"""Micro-benchmark of the fused attention kernels (CUDA events, best of 5)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morphganformer_b200 import _lib
L = _lib.lib()
def p(t): return t.data_ptr() if t is not None else None
for (B, HW, C) in [(8, 16384, 256), (8, 4096, 512), (8, 1024, 512), (8, 256, 512), (8, 64, 512), (8, 16, 512)]:
    X = torch.randn(B, HW, C, device="cuda").to(torch.bfloat16); dz = torch.randn_like(X)
    Kf = torch.randn(16, C, device="cuda") * 0.05; Sc = torch.randn(HW, 16, device="cuda"); mb = torch.zeros(B, 16, device="cuda")
    VM = torch.randn(B, 16, C, device="cuda") * 0.1; bm = torch.zeros(C, device="cuda"); noise = torch.randn(HW, device="cuda"); ns = torch.tensor([0.1], device="cuda")
    bias = torch.zeros(C, device="cuda"); out = torch.empty_like(X); dX = torch.empty_like(X); dVM = torch.zeros_like(VM); R = torch.zeros(B, C, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    # pre-built coefficient tables, as the engine passes them
    tabK = torch.empty(int(L.mgf_attn_table_bytes(0, C)), dtype=torch.uint8, device="cuda"); tabV = torch.empty(B * int(L.mgf_attn_table_bytes(1, C)), dtype=torch.uint8, device="cuda")
    _lib.check(L.mgf_attn_tables(p(Kf), p(VM), p(tabK), p(tabV), B, C, s))
    def fwd(): _lib.check(L.mgf_attn_fwd(p(X), p(Kf), p(Sc), p(mb), p(VM), p(bm), p(noise), p(ns), p(bias), 1.4, 0.2, p(out), None, None, p(tabK), p(tabV), B, HW, C, 0, s))
    def bwd(): _lib.check(L.mgf_attn_bwd(p(X), p(dz), p(Kf), p(Sc), p(mb), p(VM), p(bm), p(noise), p(ns), p(bias), 1.4, 0.2, p(dX), p(dVM), p(R), None, p(tabK), p(tabV), B, HW, C, 0, s))
    def bwd1():
        L.mgf_attn_set_split(0); bwd(); L.mgf_attn_set_split(1)
    for name, fn in (("fwd", fwd), ("bwd", bwd), ("bwd(one warp per tile)", bwd1)):
        ts = []
        for i in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if i >= 2: ts.append(e0.elapsed_time(e1))
        by = B * HW * C * 2 * (2 if name == "fwd" else 3)
        print("attn %s B=%d HW=%d C=%d: %.3f ms  (%.0f GB/s algorithmic)" % (name, B, HW, C, min(ts), by / min(ts) / 1e6), flush=True)

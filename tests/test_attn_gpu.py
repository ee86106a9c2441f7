"""Tensor-core duplex-attention kernels (mgf_attn_fwd / mgf_attn_bwd) against a plain PyTorch fp32 statement of the same folded
layer (reference training/networks.py:748-822 + :1036-1040 with the host-side constant folding of engine.py).  Floating-point
kernel -> torch fp32 reference; tolerances: outputs are stored in 16 bits (fp16: 2^-11, bf16: 2^-8 relative), gradients in bf16."""
import pytest
import torch

from morphganformer_b200 import _lib

pytestmark = pytest.mark.gpu


def _ref(X, Kf, Sc, mb, VM, bm, noise, ns, bias, gain, alpha, dmask=None):
    S = X @ Kf.t() + Sc[None] + mb[:, None, :]
    A = torch.softmax(S, -1)
    if dmask is not None:          # attention dropout (reference networks.py:505-513): applied after the softmax, no renormalisation
        A = A * dmask
    ctl = A @ VM + bm
    xn = X * torch.rsqrt(X.square().mean(-1, keepdim=True) + 1e-8)
    u = xn * (1 + ctl) + noise[None, :, None] * ns + bias
    return torch.nn.functional.leaky_relu(u, alpha) * gain, A


def _p(t):
    return t.data_ptr() if t is not None else None


@pytest.mark.parametrize("fwd", ["fp16", "bf16"])
@pytest.mark.parametrize("dropout,tables", [(False, False), (True, False), (False, True), (True, True)])
# the 512-channel cases below 148 * 4 pixel tiles take the split-channel backward kernel: 8 warps per tile (2 x 1000 ragged, 8 x 16),
# 4 (8 x 512), 2 (8 x 1024, 8 x 1000 ragged); 2 x 4096 x 512 stays on the one-warp-per-tile kernel
@pytest.mark.parametrize("B,HW,C", [(2, 16, 32), (2, 100, 64), (3, 1032, 128), (2, 4096, 256), (1, 500, 384), (2, 1000, 512), (8, 16, 512),
                                    (8, 512, 512), (8, 1024, 512), (8, 1000, 512), (2, 4096, 512)])
def test_attn_fwd_bwd_vs_torch(B, HW, C, fwd, dropout, tables):
    L = _lib.lib()
    _lib.set_forward_dtype(fwd)
    try:
        dt = torch.float16 if fwd == "fp16" else torch.bfloat16
        g = torch.Generator(device="cuda").manual_seed(B * 1000 + HW + C)
        r = lambda *s: torch.randn(*s, device="cuda", generator=g)
        X16 = (r(B, HW, C) * 1.5).to(dt)
        Kf, Sc, mb = r(16, C) * (2.0 / C ** 0.5), r(HW, 16), r(B, 16) * 0.3
        VM, bm, noise, ns, bias = r(B, 16, C) * 0.3, r(C) * 0.1, r(HW), torch.tensor([0.2], device="cuda"), r(C) * 0.1
        dz16 = r(B, HW, C).to(torch.bfloat16)
        gain, alpha = 1.4142135, 0.2
        dmask = None
        if dropout:                # cell mask x column mask, each scaled by 1/(1-p), p = 0.06 (attention_dropout / 2)
            keep = lambda *sh: (torch.rand(*sh, device="cuda", generator=g) >= 0.06).float() / 0.94
            dmask = (keep(B, HW, 16) * keep(B, 1, 16)).contiguous()
        s = torch.cuda.current_stream().cuda_stream
        tabK = tabV = None
        if tables:                 # pre-built shared-memory images of the coefficient tables (what the engine passes)
            tabK = torch.empty(L.mgf_attn_table_bytes(0, C), dtype=torch.uint8, device="cuda")
            tabV = torch.empty(B * L.mgf_attn_table_bytes(1, C), dtype=torch.uint8, device="cuda")
            _lib.check(L.mgf_attn_tables(_p(Kf), _p(VM), _p(tabK), _p(tabV), B, C, s), "tables")
        out = torch.empty_like(X16); probs = torch.empty(B, HW, 16, device="cuda")
        _lib.check(L.mgf_attn_fwd(_p(X16), _p(Kf), _p(Sc), _p(mb), _p(VM), _p(bm), _p(noise), _p(ns), _p(bias), gain, alpha, _p(out), _p(probs), _p(dmask), _p(tabK), _p(tabV), B, HW, C, 0, s), "fwd")
        dX = torch.empty(B, HW, C, device="cuda", dtype=torch.bfloat16); dVM = torch.zeros(B, 16, C, device="cuda"); R = torch.zeros(B, C, device="cuda")
        _lib.check(L.mgf_attn_bwd(_p(X16), _p(dz16), _p(Kf), _p(Sc), _p(mb), _p(VM), _p(bm), _p(noise), _p(ns), _p(bias), gain, alpha, _p(dX), _p(dVM), _p(R), _p(dmask), _p(tabK), _p(tabV), B, HW, C, 0, s), "bwd")
        torch.cuda.synchronize()
        Xr = X16.float().requires_grad_(True); VMr = VM.clone().requires_grad_(True)
        ref, A = _ref(Xr, Kf, Sc, mb, VMr, bm, noise, ns, bias, gain, alpha, dmask)
        gX, gVM = torch.autograd.grad(ref, [Xr, VMr], grad_outputs=dz16.float())
        Rr = (gX * Xr.detach()).sum(1)
        eps = 2.0 ** -10 if fwd == "fp16" else 2.0 ** -7
        scale = ref.abs().max().item()
        assert (probs - A).abs().max().item() < (2e-3 if fwd == "fp16" else 1.5e-2)
        # lrelu kink: an element whose pre-activation is within rounding of 0 may flip slope; compare in the large
        eo = (out.float() - ref).abs()
        assert eo.max().item() < 4 * eps * scale and eo.mean().item() < eps * ref.abs().mean().item()
        rel = lambda a, b: ((a - b).norm() / b.norm()).item()
        assert rel(dX.float(), gX) < 1.5e-2, rel(dX.float(), gX)
        assert rel(dVM, gVM) < 1.5e-2, rel(dVM, gVM)
        assert rel(R, Rr) < 1.5e-2, rel(R, Rr)
    finally:
        _lib.set_forward_dtype(_lib.DEFAULT_FORWARD_DTYPE)


def test_attn_rejects_bad_shapes():
    L = _lib.lib()
    x = torch.zeros(1, 16, 48, device="cuda", dtype=torch.bfloat16)
    f = torch.zeros(16, 48, device="cuda")
    rc = L.mgf_attn_fwd(_p(x), _p(f), _p(f), _p(f), _p(f), _p(f), None, None, None, 1.0, 0.2, _p(x), None, None, None, None, 1, 16, 48, 0, 0)
    assert rc != 0 and b"C=48" in L.mgf_last_error()


@pytest.mark.parametrize("res,C,B", [(8, 32, 2), (16, 64, 3)])
def test_attn_kernel_vs_oracle_transformer_layer(res, C, B):
    """The kernel against the ORACLE's restatement of the reference layer (oracle/ganformer.py transformer_layer = networks.py:748-822 with
    its separate Q / positional / centroid / value / modulation projections), not against the builder's own folded formula: the folding of
    engine.SynthesisEngine._fold_layer (Kf, Sc, VM, bm) is applied to a real layer's parameters, the kernel runs on the folded constants, and
    output + d(X) must match the oracle layer evaluated on the un-folded parameters (fp32 autograd on the CPU)."""
    import math
    import util
    from oracle import ganformer
    G = util.build_G(res, 0, 64 * res, C)             # last block: res x res with C channels
    sd = util.state_dict_cpu(G)
    pre = f"synthesis.b{res}.conv1.transformer"
    HW = res * res
    g = torch.Generator().manual_seed(res)
    X = torch.randn(B, HW, C, generator=g) * 1.3
    Y = torch.randn(B, 16, 32, generator=g)
    mask = torch.ones(B, 16); mask[0, 3] = 0
    Xr = X.half().float().requires_grad_(True)
    ref, probs_ref = ganformer.transformer_layer(sd, pre, Xr, Y, sd[f"synthesis.b{res}.conv1.grid_pos"], sd["pos"], mask.unsqueeze(1), HW, 16)
    dz = torch.randn(B, HW, C, generator=g)
    gX, = torch.autograd.grad(ref, [Xr], dz.bfloat16().float())
    # fold exactly as the engine does
    from morphganformer_b200 import engine as E
    Gc = G.cuda(); Gc.synthesis.engine = "tc"
    eng = E.SynthesisEngine(Gc.synthesis)
    Lr = [e for e in eng.blocks if e["res"] == res][0]["conv1"]
    assert Lr.attn
    VM = (Y.cuda() @ Lr.WVM.t() + Lr.bVM).contiguous()
    mb = ((1.0 - mask.cuda()) * -10000.0).contiguous()
    L, s = _lib.lib(), torch.cuda.current_stream().cuda_stream
    X16 = X.half().cuda()
    out = torch.empty_like(X16); probs = torch.empty(B, HW, 16, device="cuda")
    # no noise, no bias, linear tail (alpha = 1, gain = 1): the pure attention layer
    _lib.check(L.mgf_attn_fwd(_p(X16), _p(Lr.Kf), _p(Lr.Sc), _p(mb), _p(VM), _p(Lr.bm), None, None, None, 1.0, 1.0, _p(out), _p(probs), None, None, None, B, HW, C, 0, s), "fwd")
    dX = torch.empty(B, HW, C, device="cuda", dtype=torch.bfloat16); dVM = torch.zeros(B, 16, C, device="cuda"); R = torch.zeros(B, C, device="cuda")
    dz16 = dz.bfloat16().cuda()
    _lib.check(L.mgf_attn_bwd(_p(X16), _p(dz16), _p(Lr.Kf), _p(Lr.Sc), _p(mb), _p(VM), _p(Lr.bm), None, None, None, 1.0, 1.0, _p(dX), _p(dVM), _p(R), None, None, None, B, HW, C, 0, s), "bwd")
    torch.cuda.synchronize()
    assert (probs.cpu() - probs_ref.detach().reshape(B, HW, 16)).abs().max().item() < 2e-3
    scale = ref.detach().abs().max().item()
    assert (out.float().cpu() - ref.detach()).abs().max().item() < 4 * 2.0 ** -10 * scale
    assert ((dX.float().cpu() - gX).norm() / gX.norm()).item() < 1.5e-2


@pytest.mark.parametrize("B,HW", [(8, 16), (8, 256), (8, 512), (8, 1024)])
def test_attn_bwd_split_kernel_equals_one_warp_per_tile_kernel(B, HW):
    """The split-channel backward kernel (several warps per pixel tile, partial sums exchanged through shared memory) against the
    one-warp-per-tile kernel on the same inputs: same arithmetic up to the order of the partial sums."""
    L = _lib.lib()
    C = 512
    g = torch.Generator(device="cuda").manual_seed(HW)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    X16 = (r(B, HW, C) * 1.5).to(_lib.forward_torch_dtype())
    Kf, Sc, mb = r(16, C) * (2.0 / C ** 0.5), r(HW, 16), r(B, 16) * 0.3
    VM, bm, noise, ns, bias = r(B, 16, C) * 0.3, r(C) * 0.1, r(HW), torch.tensor([0.2], device="cuda"), r(C) * 0.1
    dz16 = r(B, HW, C).to(torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream
    outs = []
    for split in (1, 0):
        L.mgf_attn_set_split(split)
        dX = torch.empty(B, HW, C, device="cuda", dtype=torch.bfloat16); dVM = torch.zeros(B, 16, C, device="cuda"); R = torch.zeros(B, C, device="cuda")
        _lib.check(L.mgf_attn_bwd(_p(X16), _p(dz16), _p(Kf), _p(Sc), _p(mb), _p(VM), _p(bm), _p(noise), _p(ns), _p(bias), 1.4142135, 0.2, _p(dX), _p(dVM), _p(R),
                                  None, None, None, B, HW, C, 0, s), "bwd")
        torch.cuda.synchronize()
        outs.append((dX.float(), dVM, R))
    L.mgf_attn_set_split(1)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(outs[0][0], outs[1][0]) < 4e-3 and rel(outs[0][1], outs[1][1]) < 2e-3 and rel(outs[0][2], outs[1][2]) < 2e-3

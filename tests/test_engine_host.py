"""CPU: host-side constant folding of the bf16 engine against the oracle ops (no GPU needed)."""
import math
import numpy as np
import torch
import torch.nn.functional as F
from oracle import ops as O
import util


def test_up_phase_matrix_reproduces_reference_upconv():
    from morphganformer_b200.engine import up_phase_matrix
    Cm = torch.from_numpy(up_phase_matrix()).float()                 # [4, 9, 9]
    x = util.case_tensor((2, 5, 7, 6), 1)
    w = util.case_tensor((4, 5, 3, 3), 2) * 0.3
    f = O.setup_filter([1, 3, 3, 1])
    ref = O.conv2d_resample(x, w, f=f, up=2, padding=1, flip_weight=False)
    Weff = torch.einsum("ptk,oik->ptoi", Cm, w.reshape(4, 5, 9))     # [4, 9, O, I]
    out = torch.zeros_like(ref)
    for ph in range(4):
        py, px = ph // 2, ph % 2
        wk = Weff[ph].reshape(3, 3, 4, 5).permute(2, 3, 0, 1)        # [O, I, ty, tx] correlation taps
        out[:, :, py::2, px::2] = F.conv2d(x, wk, padding=1)
    assert (out - ref).abs().max() < 1e-5


def test_skip_path_formula():
    """1x1 conv then FIR up2 with flipped normalised taps and gain 4*sqrt(.5) (what mgf_upfir2_add computes)."""
    x = util.case_tensor((1, 3, 5, 5), 3)
    w = util.case_tensor((4, 3, 1, 1), 4)
    f = O.setup_filter([1, 3, 3, 1])
    ref = O.bias_act(O.conv2d_resample(x, w, f=f, up=2, padding=0, flip_weight=False), None, act="linear", gain=math.sqrt(0.5))
    v = F.conv2d(x, w)
    fk = (np.array([1, 3, 3, 1.0]) / 8)[::-1].copy()
    h = 5
    out = torch.zeros(1, 4, 10, 10)
    for Y in range(10):
        for X in range(10):
            acc = 0
            for fy in range(Y & 1, 4, 2):
                iy = (Y + fy - 2) >> 1
                if iy < 0 or iy >= h:
                    continue
                for fx in range(X & 1, 4, 2):
                    ix = (X + fx - 2) >> 1
                    if ix < 0 or ix >= h:
                        continue
                    acc = acc + fk[fy] * fk[fx] * v[:, :, iy, ix]
            out[:, :, Y, X] = acc * 4 * math.sqrt(0.5)
    assert (out - ref).abs().max() < 1e-5


def test_attention_fold_matches_oracle_layer():
    """S = X Kf^T + Sc and ctl = A VM + bm reproduce TransformerLayer.forward (oracle restatement) in fp64."""
    from oracle import ganformer
    G = util.build_G(32, 0, 512, 32)
    sd = {k: v.double() for k, v in util.state_dict_cpu(G).items()}
    pre = "synthesis.b8.conv1"
    C, HW = 32, 64
    X = util.case_tensor((2, HW, C), 5).double()
    Y = util.case_tensor((2, 16, 32), 6).double()
    mask = torch.ones(2, 16).double(); mask[1, 3] = 0
    ref, probs = ganformer.transformer_layer(sd, pre + ".transformer", X, Y, sd[pre + ".grid_pos"], sd["pos"], mask.unsqueeze(1), HW, 16)
    t = pre + ".transformer"
    rs = 1 / math.sqrt(C)
    Wq = sd[t + ".to_queries.weight"] / math.sqrt(C); bq = sd[t + ".to_queries.bias"]
    aw = sd[t + ".att_weight"].reshape(-1); cen = sd[t + ".centroids"][0, 0]
    a1, a2 = cen[:, :C] * aw[:C], cen[:, C:] * aw[C:]
    Kf = (a1 @ Wq) * rs
    P = sd[pre + ".grid_pos"].reshape(-1, 32) @ (sd[t + ".from_pos_map.weight"] / math.sqrt(32)).t() + sd[t + ".from_pos_map.bias"]
    Sc = ((a1 @ bq).unsqueeze(0) + P @ a2.t()) * rs
    S = X @ Kf.t() + Sc + ((1 - mask) * -10000.0).unsqueeze(1)
    A = torch.softmax(S, -1)
    assert (A - probs[:, 0]).abs().max() < 1e-9
    Wv = sd[t + ".to_values.weight"] / math.sqrt(32); bv = sd[t + ".to_values.bias"]
    Wm = sd[t + ".modulation.weight"] / math.sqrt(C); bm = sd[t + ".modulation.bias"]
    VM = Y @ (Wm @ Wv).t() + Wm @ bv
    ctl = A @ VM + bm
    xn = X * torch.rsqrt(X.square().mean(-1, keepdim=True) + 1e-8)
    assert (xn * (1 + ctl) - ref).abs().max() < 1e-9


def test_two_stage_upconv_input_gradient_formula():
    """What the engine computes for an up-convolution's input gradient up to 128^2 (mgf_fir4_pad + a stride-2 3x3 conv over four phase views):
    g[u,v] = sum_{a,b} F[a,b] dy[u-a+1, v-b+1] on the (2h+1) x (2w+1) grid (F = outer([1,3,3,1]/8)^2 * 4), dx[i,j] = sum_k W_k^T g[2i+ky, 2j+kx]
    equals autograd through the oracle's conv2d_resample(up=2) (reference conv2d_resample.py:117-134)."""
    B, I, Oc, h, w = 2, 5, 4, 6, 7
    x = util.case_tensor((B, I, h, w), 1).requires_grad_(True)
    W = util.case_tensor((Oc, I, 3, 3), 2) * 0.3
    f = O.setup_filter([1, 3, 3, 1])
    y = O.conv2d_resample(x, W, f=f, up=2, padding=1, flip_weight=False)
    dy = util.case_tensor(tuple(y.shape), 3)
    gx_ref, = torch.autograd.grad(y, [x], dy)
    f1 = np.array([1, 3, 3, 1], dtype=np.float64) / 8.0
    Ff = torch.from_numpy(np.outer(f1, f1) * 4.0).float()
    dyp = F.pad(dy, (2, 2, 2, 2))
    g = torch.zeros(B, Oc, 2 * h + 1, 2 * w + 1)
    for a in range(4):
        for b in range(4):
            g += Ff[a, b] * dyp[:, :, 3 - a:3 - a + 2 * h + 1, 3 - b:3 - b + 2 * w + 1]
    # the engine's tap table: phase view (ky%2, kx%2) of g, shifted by (ky//2, kx//2)
    gq = F.pad(g, (0, 1, 0, 1))                                       # [.., 2h+2, 2w+2] like the kernel's output buffer
    dx = torch.zeros_like(x)
    for ky in range(3):
        for kx in range(3):
            view = gq[:, :, ky % 2::2, kx % 2::2]                     # (h+1) x (w+1)
            sl = view[:, :, ky // 2:ky // 2 + h, kx // 2:kx // 2 + w]
            dx = dx + torch.einsum("bohw,oc->bchw", sl, W[:, :, ky, kx])
    assert (dx - gx_ref).abs().max() < 1e-5 * gx_ref.abs().max().clamp(min=1.0)


def test_two_stage_upconv_forward_tap_table_and_fir_offsets():
    """What the engine computes for an up-convolution's FORWARD pass: the transposed convolution as four parity GEMMs over the (h+1) x (w+1)
    grid with 4 / 2 / 2 / 1 taps on the plain [9, O, I] weights (tc.conv_tc(phase_ntaps=...), written into the [2h+2, 2w+2] buffer through the
    strided phase views), then mgf_fir4(off = -1): out[Y, X] = 4 * sum_{t,u} f[t] f[u] ct[Y-1+t, X-1+u] with zero padding -- equals the oracle's
    conv2d_resample(up=2, padding=1, flip_weight=False) (reference conv2d_resample.py:117-134)."""
    B, I, Oc, h, w = 2, 5, 4, 6, 7
    x = util.case_tensor((B, I, h, w), 1)
    W = util.case_tensor((Oc, I, 3, 3), 2) * 0.3
    f = O.setup_filter([1, 3, 3, 1])
    ref = O.conv2d_resample(x, W, f=f, up=2, padding=1, flip_weight=False)
    taps = ([(0, -a, -b, (2 * a) * 3 + 2 * b) for a in (0, 1) for b in (0, 1)] + [(0, -a, 0, (2 * a) * 3 + 1) for a in (0, 1)]
            + [(0, 0, -b, 3 + 2 * b) for b in (0, 1)] + [(0, 0, 0, 4)])             # engine._fold_layer, two_stage_fwd
    counts, ofy, ofx = (4, 2, 2, 1), (0, 0, 1, 1), (0, 1, 0, 1)
    Wk = W.reshape(Oc, I, 9)
    ct = torch.full((B, Oc, 2 * h + 2, 2 * w + 2), float("nan"))
    xp = F.pad(x, (1, 1, 1, 1))                                                  # TMA zero fill outside the image
    t0 = 0
    for ph in range(4):
        acc = torch.zeros(B, Oc, h + 1, w + 1)
        for (_, dy, dx, wz) in taps[t0:t0 + counts[ph]]:
            sl = xp[:, :, 1 + dy:1 + dy + h + 1, 1 + dx:1 + dx + w + 1]         # x[m + dy, n + dx] for m in [0, h], n in [0, w]
            acc = acc + torch.einsum("bihw,oi->bohw", sl, Wk[:, :, wz])
        ct[:, :, ofy[ph]::2, ofx[ph]::2] = acc
        t0 += counts[ph]
    assert torch.isfinite(ct).all() and ct[:, :, 2 * h + 1].abs().max() == 0 and ct[:, :, :, 2 * w + 1].abs().max() == 0
    f1 = torch.tensor([1, 3, 3, 1.0]) / 8
    cp = F.pad(ct, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for t in range(4):
        for u in range(4):
            out += 4 * f1[t] * f1[u] * cp[:, :, t:t + 2 * h, u:u + 2 * w]          # ct[Y - 1 + t, X - 1 + u]
    assert (out - ref).abs().max() < 1e-5


def test_mapping_pack_layout_reproduces_oracle_mapping():
    """pack_mapping's flat layout (what mgf_mapping_fwd reads: mapping.cu) interpreted in numpy reproduces the oracle's ws: pins the gain /
    positional-map folding and the offsets on the CPU."""
    from morphganformer_b200 import mapping_engine
    from oracle import ganformer
    G = util.build_G(16, 3, 512, 32)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n, p in G.mapping.named_parameters():
            if n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * (30.0 if ".l" in n or "out_layer" in n else 0.3))
    P = mapping_engine.pack_mapping(G).double().numpy()
    z = util.case_tensor((2, 17, 32), 40)
    with torch.no_grad():
        _, ws_ref = ganformer.generator(util.state_dict_cpu(G), z, 16)
    D, T, W = 32, 16, 1024
    lrelu = lambda v: np.where(v > 0, v, 0.2 * v)
    take = lambda off, n, shape: P[off:off + n].reshape(shape)
    LSZ = 6 * W + 2 * T * D + 4 * D
    out = np.zeros((2, 17, 32))
    for b in range(2):
        zb = z[b].double().numpy()
        x = zb[:16] / np.sqrt((zb[:16] ** 2).mean() + 1e-8)
        for l in range(4):
            o = l * LSZ
            Wq, Wk, Wv, Wm, W0, W1 = (take(o + i * W, W, (D, D)) for i in range(6))
            Cq, Ck = take(o + 6 * W, T * D, (T, D)), take(o + 6 * W + T * D, T * D, (T, D))
            bv, bm, b0, b1 = (take(o + 6 * W + 2 * T * D + i * D, D, (D,)) for i in range(4))
            q, k, v = x @ Wq.T + Cq, x @ Wk.T + Ck, x @ Wv.T + bv
            s = q @ k.T / np.sqrt(32.0)
            A = np.exp(s - s.max(1, keepdims=True)); A /= A.sum(1, keepdims=True)
            xs = x + (A @ v) @ Wm.T + bm
            h0 = lrelu(xs @ W0.T + b0) * np.sqrt(2.0)
            x = lrelu(h0 @ W1.T + b1 + x)
        o = 4 * LSZ
        out[b, :16] = lrelu(x @ take(o, W, (D, D)).T + take(o + W, D, (D,))) * np.sqrt(2.0)
        go = o + W + D
        xg = zb[16] / np.sqrt((zb[16] ** 2).mean() + 1e-8)
        GSZ = 2 * W + 2 * D
        for l in range(4):
            o2 = go + l * GSZ
            W0, W1 = take(o2, W, (D, D)), take(o2 + W, W, (D, D))
            b0, b1 = take(o2 + 2 * W, D, (D,)), take(o2 + 2 * W + D, D, (D,))
            xg = lrelu(lrelu(xg @ W0.T + b0) * np.sqrt(2.0) @ W1.T + b1 + xg)
        o2 = go + 4 * GSZ
        out[b, 16] = lrelu(xg @ take(o2, W, (D, D)).T + take(o2 + W, D, (D,))) * np.sqrt(2.0)
    assert o2 + W + D == P.size
    np.testing.assert_allclose(out, ws_ref[:, :, 0].double().numpy(), rtol=1e-5, atol=1e-6)

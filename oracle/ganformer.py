"""TEST INFRASTRUCTURE ONLY (never imported by the product package).

Functional CPU restatement (plain PyTorch CPU ops, fp32 or fp64) of the reference's GANformer
generator forward for the GANformer-default configuration (run_network.py:61-77): mapping network
with latent self-attention, synthesis network with duplex ("kmeans", parametric centroids) attention,
fused modulated convolution with weight demodulation, resnet blocks, ToRGB on the last block.

Works on a *state_dict* with the reference's parameter names (training/networks.py), so the same
weights can be pushed through the real reference (container only), through this oracle, and through
the CUDA product path.  Pinned against the real reference by tests/test_oracle_vs_reference.py
(container only) and by tests/golden/gen32_golden.npz (committed, made by tests/golden/make_golden.py).

The restatement keeps the reference's op order (including work that is dead under the default
config only where it changes rounding), so that fp32 results agree to ~1e-6.
"""
import math
import torch
import torch.nn.functional as F
from . import ops

SQRT2 = math.sqrt(2.0)


def _fc(sd, name, x, lrmul=1.0, act="linear"):
    """FullyConnectedLayer.forward, networks.py:138-150 (+ get_weight :69-84 runtime_coef = lrmul/sqrt(fan_in))."""
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    w = w.to(x.dtype) * (lrmul / math.sqrt(w.shape[1]))
    if b is not None:
        b = b.to(x.dtype) * lrmul if lrmul != 1 else b.to(x.dtype)
    if x.ndim > 2:
        x = x.flatten(1)
    if act == "linear" and b is not None:
        return torch.addmm(b.unsqueeze(0), x, w.t())
    x = x.matmul(w.t())
    return ops.bias_act(x, b, act=act)


def _normalize(x, eps=1e-8):
    """normalize(), networks.py:30-37 (l2 mode)."""
    dims = list(range(1, x.ndim))
    return x * (x.square().mean(dim=dims, keepdim=True) + eps).rsqrt()


def transformer_layer(sd, pre, from_tensor, to_tensor, from_pos, to_pos, att_mask, from_len, to_len,
                      integration="mul", norm="layer", kmeans=True, lrmul=1.0, dmask=None):
    """TransformerLayer.forward, networks.py:748-822, for num_heads=1, kmeans_iters=1, eval mode
    (dropout = multiply by ones, :374-376), parametric centroids (:715-717), no gates (:546-547).
    from_tensor [B*F or B,F, C]; to_tensor [B,T,Dt]; from_pos [F, P] or None; to_pos [T,P] or None.
    Returns (new from_tensor (same shape), att_probs [B,1,F,T])."""
    shape = from_tensor.shape
    ft = from_tensor.reshape(-1, shape[-1])
    tt = to_tensor.reshape(-1, to_tensor.shape[-1])
    bsz = ft.shape[0] // from_len
    dim = sd[pre + ".to_queries.weight"].shape[0]
    q = _fc(sd, pre + ".to_queries", ft, lrmul)
    k = _fc(sd, pre + ".to_keys", tt, lrmul)
    v = _fc(sd, pre + ".to_values", tt, lrmul)
    q0 = q
    if from_pos is not None:
        q = q + _fc(sd, pre + ".from_pos_map", from_pos.reshape(-1, from_pos.shape[-1]).repeat(bsz, 1), lrmul)
    if to_pos is not None:
        k = k + _fc(sd, pre + ".to_pos_map", to_pos.repeat(bsz, 1), lrmul)
    v = v.reshape(bsz, 1, to_len, dim)
    qh = q.reshape(bsz, 1, from_len, dim)
    kh = k.reshape(bsz, 1, to_len, dim)
    scores = qh.matmul(kh.transpose(-1, -2))                           # :776
    if kmeans:
        fe = torch.cat([q0, q - q0], dim=-1).reshape(bsz, 1, from_len, 2 * dim)   # :688-689
        cen = sd[pre + ".centroids"].to(ft.dtype).repeat(bsz, 1, 1, 1)             # :717
        scores = (fe * sd[pre + ".att_weight"].to(ft.dtype)).matmul(cen.transpose(-1, -2))  # :792
    scores = scores / math.sqrt(float(dim))                            # :795
    if att_mask is not None:
        scores = scores + (1 - att_mask.unsqueeze(1).to(ft.dtype)) * -10000.0   # :799, :379-380
    probs = F.softmax(scores, dim=-1)                                  # :507
    if dmask is not None:            # training mode, :510-512: probs = dropout(probs) over cells, then over whole 'to' columns; the two
        probs = probs * dmask        # torch.nn.Dropout masks (keep / (1 - p)) are injected pre-multiplied, broadcastable to [B,1,F,T]
    control = probs.matmul(v).permute(0, 2, 1, 3).reshape(-1, dim)     # :812-814
    # integrate(), :657-672, with att_norm :341-358
    x = ft
    if norm is not None:
        xs = x.reshape(bsz, from_len, -1)
        ax = 1 if norm == "instance" else 2
        if integration in ("add", "both"):
            xs = xs - xs.mean(dim=ax, keepdim=True)
        if integration in ("mul", "both"):
            xs = xs * torch.rsqrt(xs.square().mean(dim=ax, keepdim=True) + 1e-8)
        x = xs.reshape(ft.shape)
    control = _fc(sd, pre + ".modulation", control, lrmul)
    gain = bias = control
    if integration == "both":
        gain, bias = torch.split(control, 2, dim=-1)
    if integration != "add":
        x = x * (gain + 1)
    if integration != "mul":
        x = x + bias
    return x.reshape(shape), probs


def mapping(sd, z, pos, mask, k=17, num_layers=8, lrmul=0.01, num_ws=None):
    """MappingNetwork.forward networks.py:894-942 (transformer=True, resnet=True, ltnt2ltnt=True, use_pos=True,
    normalize_global=True, c_dim=0, truncation_psi=1); MLP :179-221; ResnetLayer :154-172."""
    zl, g = torch.split(z, [k - 1, 1], dim=1)
    g = _normalize(g)
    zl = _normalize(zl)

    def mlp(pre, x, sa):
        shape = x.shape
        x = x.reshape(-1, shape[-1])
        nl = num_layers // 2
        for i in range(nl):
            x0 = x
            if sa:
                x, _ = transformer_layer(sd, f"{pre}.sa{i}", x, x, pos, pos, mask.unsqueeze(1), k - 1, k - 1,
                                         integration="add", norm=None, kmeans=False)
            h = _fc(sd, f"{pre}.l{i}.fc0", x, lrmul, act="lrelu")
            h = _fc(sd, f"{pre}.l{i}.fc1", h, lrmul)
            x = F.leaky_relu(h + x0, negative_slope=0.2)
        x = _fc(sd, f"{pre}.out_layer", x, lrmul, act="lrelu")
        return x.reshape(*shape[:-1], -1)

    xg = mlp("mapping.global_mlp", g, False)
    xl = mlp("mapping.mlp", zl, True)
    x = torch.cat([xl, xg], dim=1)
    if num_ws is not None:
        x = x.unsqueeze(2).repeat(1, 1, num_ws, 1)
    return x


def modulated_conv2d(x, weight, styles, up=1, padding=0, f=None, demodulate=True, flip_weight=True):
    """modulated_conv2d fused path, networks.py:252-308."""
    b = x.shape[0]
    oc, ic, kh, kw = weight.shape
    w = weight.unsqueeze(0) * styles.reshape(b, 1, -1, 1, 1)
    if demodulate:
        d = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
        w = w * d.reshape(b, -1, 1, 1, 1)
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, ic, kh, kw)
    x = ops.conv2d_resample(x, w, f=f, up=up, padding=padding, groups=b, flip_weight=flip_weight)
    return x.reshape(b, -1, *x.shape[2:])


def synthesis_layer(sd, pre, x, w, pos, mask, out_res, up=1, gain=1.0, attention=True, bias=True, noise=True,
                    noise_mode="const", f=None, rand_noise=None, dmask=None):
    """SynthesisLayer.forward, networks.py:1010-1042."""
    nz = None
    if noise and noise_mode != "none":
        if noise_mode == "random":
            nz = rand_noise
        else:
            nz = sd[pre + ".noise_const"].to(x.dtype)
        nz = nz * sd[pre + ".noise_strength"].to(x.dtype)
    wt = sd[pre + ".weight"].to(x.dtype)
    wt = wt * (1.0 / math.sqrt(wt[0].numel()))
    styles = _fc(sd, pre + ".affine", w[:, -1])
    x = modulated_conv2d(x, wt, styles, up=up, padding=wt.shape[-1] // 2, f=f, flip_weight=(up == 1))
    att = None
    if attention:
        shape = x.shape
        xt = x.reshape(shape[0], shape[1], -1).permute(0, 2, 1)
        xt, att = transformer_layer(sd, pre + ".transformer", xt, w[:, :-1], sd[pre + ".grid_pos"].to(x.dtype), pos,
                                    mask.unsqueeze(1), out_res * out_res, w.shape[1] - 1, dmask=dmask)
        x = xt.permute(0, 2, 1).reshape(shape)
    if nz is not None:
        x = x + nz
    if bias:
        x = ops.bias_act(x, sd[pre + ".biasAct.bias"].to(x.dtype), act="lrelu", gain=SQRT2 * gain)
    return x, att


def synthesis(sd, ws, pos, mask, res, architecture="resnet", end_res=8, noise_mode="const", return_att=False,
              dtype=torch.float32, trace=None, inject=None):
    """SynthesisNetwork.forward networks.py:1244-1264 + SynthesisBlock.forward :1132-1174 (resnet architecture,
    const stem, ToRGB only on the last block).  ws [B,k,num_ws,w_dim]."""
    assert architecture == "resnet"
    # inject (training-mode runs): {"noise": {ws index: [B,1,r,r] plane of noise_mode='random'}, "dmask": {ws index: [B,1,r*r,16] dropout mask}}
    inj_n = (inject or {}).get("noise", {})
    inj_d = (inject or {}).get("dmask", {})
    ws = ws.to(dtype)
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    pos = pos.to(dtype)
    f = ops.setup_filter([1, 3, 3, 1]).to(ws.device, dtype)
    b = ws.shape[0]
    resolutions = [2 ** i for i in range(2, int(math.log2(res)) + 1)]
    w_idx = 0
    x = None
    atts = []
    img = None
    for r in resolutions:
        pre = f"synthesis.b{r}"
        attn = math.log2(r) < end_res
        last = r == res
        if r == 4:
            x = sd[pre + ".const"].unsqueeze(0).repeat(b, 1, 1, 1)
            x, a = synthesis_layer(sd, pre + ".conv1", x, ws[:, :, w_idx], pos, mask, r, attention=attn,
                                   noise_mode=noise_mode, f=f, rand_noise=inj_n.get(w_idx), dmask=inj_d.get(w_idx))
            atts.append(a)
            if trace is not None:
                trace[f"z{w_idx}"] = x
            w_idx += 1
        else:
            wsk = sd[pre + ".skip.weight"]
            wsk = wsk * (1.0 / math.sqrt(wsk[0].numel()))
            y = ops.conv2d_resample(x, wsk, f=f, up=2, padding=0, flip_weight=False)
            y = ops.bias_act(y, None, act="linear", gain=math.sqrt(0.5))
            x, a0 = synthesis_layer(sd, pre + ".conv0", x, ws[:, :, w_idx], pos, mask, r, up=2, attention=attn,
                                    noise_mode=noise_mode, f=f, rand_noise=inj_n.get(w_idx), dmask=inj_d.get(w_idx))
            if trace is not None:
                trace[f"z{w_idx}"] = x
            x, a1 = synthesis_layer(sd, pre + ".conv1", x, ws[:, :, w_idx + 1], pos, mask, r, gain=math.sqrt(0.5),
                                    attention=attn, noise_mode=noise_mode, f=f, rand_noise=inj_n.get(w_idx + 1), dmask=inj_d.get(w_idx + 1))
            atts += [a0, a1]
            if trace is not None:
                trace[f"z{w_idx + 1}"] = x
            x = y + x
            if trace is not None:
                trace[f"xout{r}"] = x
            w_idx += 2
        if last:
            x, _ = synthesis_layer(sd, pre + ".conv_last", x, ws[:, :, w_idx], pos, mask, r, attention=False,
                                   bias=False, noise=False, f=f)
            w_idx += 1
            wr = sd[pre + ".torgb.weight"]
            styles = _fc(sd, pre + ".torgb.affine", ws[:, -1, w_idx]) * (1.0 / math.sqrt(wr[0].numel()))
            y = modulated_conv2d(x, wr, styles, demodulate=False)
            img = ops.bias_act(y, sd[pre + ".torgb.biasAct.bias"])
    if return_att:
        return img, atts
    return img


def num_ws_for(res):
    """SynthesisNetwork.__init__ networks.py:1207-1218: 1 for b4, 2 per block, +conv_last +torgb on the last."""
    n = int(math.log2(res)) - 1
    return 1 + 2 * (n - 1) + 2


def generator(sd, z, res, noise_mode="const", dtype=torch.float32):
    """Generator.forward networks.py:1304-1331 with c=None, truncation_psi=1, eval mode (mask = ones, :366-368)."""
    b = z.shape[0]
    k = z.shape[1]
    sdd = {kk: (v.to(dtype) if v.is_floating_point() else v) for kk, v in sd.items()}
    mask = torch.ones(b, k - 1, dtype=dtype, device=z.device)
    ws = mapping(sdd, z.to(dtype), sdd["pos"], mask, k=k, num_ws=num_ws_for(res))
    return synthesis(sdd, ws, sdd["pos"], mask, res, noise_mode=noise_mode, dtype=dtype), ws

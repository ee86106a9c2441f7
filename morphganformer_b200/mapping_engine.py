"""Mapping network z -> ws on two kernels (mgf_mapping_fwd / mgf_mapping_bwd) instead of ~400 eager PyTorch launches.

Covers the GANformer-default MappingNetwork (reference training/networks.py:833-942: 16 local + 1 global latents of 32 dims, 8-layer
resnet MLPs = 4 blocks, latent-to-latent self-attention with positional maps, lrmul 0.01, truncation_psi 1).  `supported(G)` says
whether a generator fits; other configurations keep the PyTorch module (`G.mapping`, autograd) -- that is a different GPU path of
the same mirror, not a CPU fallback."""
import torch
from . import _lib


def supported(G):
    m = G.mapping
    try:
        ok = (m.transformer and m.z_dim == 32 and m.w_dim == 32 and m.k == 17 and m.c_dim == 0 and m.num_layers == 8 and m.normalize_global
              and m.num_broadcast is not None and m.mlp.sa and not m.mlp.pool and m.mlp.layers_num == 4 and m.global_mlp.layers_num == 4
              and not m.global_mlp.sa)
        for i in range(4):
            sa = getattr(m.mlp, f"sa{i}")
            ok = ok and (not sa.kmeans) and sa.integration == "add" and sa.norm is None and sa.num_heads == 1 and sa.dim == 32
            for mlp in (m.mlp, m.global_mlp):
                blk = getattr(mlp, f"l{i}")
                ok = ok and blk.fc0.act == "lrelu" and blk.fc1.act == "linear"
        ok = ok and m.mlp.out_layer.act == "lrelu" and m.global_mlp.out_layer.act == "lrelu"
        return bool(ok)
    except AttributeError:
        return False


def _wb(fc):
    """FullyConnectedLayer -> (W [out,in] * w_gain, b * b_gain) as used at run time (reference get_weight :69-84, get_param :58-66)."""
    w = fc.weight.detach().float() * fc.w_gain
    b = fc.bias.detach().float() * fc.b_gain if fc.bias is not None else torch.zeros(w.shape[0], device=w.device)
    return w, b


@torch.no_grad()
def pack_mapping(G):
    """Packs the mapping weights into the flat fp32 layout of mapping.cu (gains folded, positional maps folded into per-token biases)."""
    m, dev = G.mapping, G.pos.device
    pos = G.pos.detach().float() if m.use_pos else None
    parts = []
    for i in range(4):
        sa, blk = getattr(m.mlp, f"sa{i}"), getattr(m.mlp, f"l{i}")
        (wq, bq), (wk, bk), (wv, bv), (wm, bm) = _wb(sa.to_queries), _wb(sa.to_keys), _wb(sa.to_values), _wb(sa.modulation)
        (w0, b0), (w1, b1) = _wb(blk.fc0), _wb(blk.fc1)
        cq, ck = bq.expand(16, 32).clone(), bk.expand(16, 32).clone()
        if pos is not None:
            wfp, bfp = _wb(sa.from_pos_map); wtp, btp = _wb(sa.to_pos_map)
            cq = cq + pos @ wfp.t() + bfp
            ck = ck + pos @ wtp.t() + btp
        parts += [wq, wk, wv, wm, w0, w1, cq, ck, bv, bm, b0, b1]
    parts += list(_wb(m.mlp.out_layer))
    for i in range(4):
        blk = getattr(m.global_mlp, f"l{i}")
        (w0, b0), (w1, b1) = _wb(blk.fc0), _wb(blk.fc1)
        parts += [w0, w1, b0, b1]
    parts += list(_wb(m.global_mlp.out_layer))
    flat = torch.cat([p.reshape(-1).to(dev) for p in parts]).contiguous()
    n = _lib.lib().mgf_mapping_param_floats()
    if flat.numel() != n:
        raise _lib.MgfError("pack_mapping: %d floats packed, the library expects %d" % (flat.numel(), n))
    return flat


class MappingEngine:
    """ws = forward(z); dz = backward(dws).  Buffers are allocated once per batch size (stable addresses: CUDA-graph capturable)."""

    def __init__(self, G):
        if not supported(G):
            raise _lib.MgfError("MappingEngine: this mapping configuration is not the GANformer default (use G.mapping)")
        self.G, self.num_ws = G, G.num_ws
        self.params = pack_mapping(G)
        self.dev = self.params.device
        self._buf = {}

    def refresh(self):
        self.params.copy_(pack_mapping(self.G))

    def _get(self, name, shape):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self._buf[name] = torch.empty(shape, device=self.dev, dtype=torch.float32)
        return t

    def forward(self, z, mask):
        _lib.require_cuda(z, "MappingEngine.forward")
        B = z.shape[0]
        self.z = z.detach().float().contiguous()
        self.maskbias = ((1.0 - mask.float()) * -10000.0).contiguous()
        ws = self._get("ws", (B, 17, self.num_ws, 32))
        _lib.check(_lib.lib().mgf_mapping_fwd(self.z.data_ptr(), self.params.data_ptr(), self.maskbias.data_ptr(), ws.data_ptr(), B, self.num_ws,
                                              _lib.stream_ptr(self.dev)), "mgf_mapping_fwd")
        return ws

    def backward(self, dws):
        B = self.z.shape[0]
        dws = dws.float().contiguous()
        dz = self._get("dz", (B, 17, 32))
        _lib.check(_lib.lib().mgf_mapping_bwd(self.z.data_ptr(), self.params.data_ptr(), self.maskbias.data_ptr(), dws.data_ptr(), dz.data_ptr(), B,
                                              self.num_ws, _lib.stream_ptr(self.dev)), "mgf_mapping_bwd")
        return dz

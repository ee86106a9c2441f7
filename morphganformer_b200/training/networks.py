"""Host-side mirror of the reference generator (training/networks.py) for the GANformer hot path.

Same class names, constructor keywords, parameter/buffer names (state_dict compatible) and parameter creation
ORDER as the reference, so `torch.manual_seed(s); Generator(**kwargs)` yields bit-identical random-init weights
and reference checkpoints (state_dict form) load with `load_state_dict`.  Same call surface:
`G(z, c, ws=..., truncation_psi=..., subnet=...)`, `G.mapping(z, c, pos=, mask=)`,
`G.synthesis(ws, pos=, mask=, noise_mode=, fused_modconv=) -> (img, att_maps)`.

Two execution engines sit behind `G.synthesis`:
  * engine="ops"  (default for fp32 / autograd): layer by layer through this package's torch_utils.ops mirror
    (bias_act, upfirdn2d, conv2d_resample -> conv2d_gradfix, fma), i.e. the reference's own op graph on the
    libmgf_sm100a.so kernels; exact fp32, double-backward capable.
  * engine="tc": the fused bf16 tcgen05 engine (morphganformer_b200/engine.py): implicit-GEMM modulated convolutions
    with demodulation folded into the weight tile, fused duplex attention, explicit backward wrt ws.

Covered configuration space: the GANformer-default generator (run_network.py:61-77) and its StyleGAN-style
sub-cases: architecture resnet|skip|orig, const stem, style modulation, duplex attention with parametric
k-means centroids, sinusoidal grid encoding, one or more heads.  Gates, latent stem, iterative (non-parametric)
centroids and trainable positional encodings raise NotImplementedError.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from ..torch_utils import misc as torch_misc
from ..torch_utils.ops import bias_act, conv2d_resample, fma, upfirdn2d


def float_dtype():
    return torch.float32


def get_gain(arch):
    return math.sqrt(0.5) if arch == "resnet" else 1


def get_global(ws):
    return ws[:, -1]


def get_components(ws):
    return ws[:, :-1]


def _make_weight(shape, gain=1, lrmul=1):
    """reference get_weight (:69-84) with use_wscale=True: N(0,1)/init_std parameter + runtime coefficient."""
    fan_in = int(np.prod(shape[1:]))
    init_std = 1.0 / lrmul
    return torch.nn.Parameter(torch.randn(shape) / init_std), gain / math.sqrt(fan_in) * lrmul


def _scaled(param, dtype, gain, reorder=False):
    """reference get_param (:58-66)."""
    if param is None:
        return None
    if gain != 1 and reorder:
        param = param * gain
    param = param.to(dtype)
    if gain != 1 and not reorder:
        param = param * gain
    return param


def normalize(x, eps=1e-8):
    dims = list(range(1, x.ndim))
    x = x.to(float_dtype())
    return x * (x.square().mean(dim=dims, keepdim=True) + eps).rsqrt()


class BiasActLayer(torch.nn.Module):
    def __init__(self, num_channels, bias=True, act="linear", lrmul=1, bias_init=0, clamp=None, gain=1):
        super().__init__()
        self.bias = torch.nn.Parameter(torch.full([num_channels], np.float32(bias_init))) if bias else None
        self.b_gain = lrmul if bias else None
        self.out_gain = bias_act.activation_funcs[act].def_gain * gain
        self.out_clamp = clamp * gain if clamp is not None else None
        self.act = act

    def forward(self, x):
        return bias_act.bias_act(x, _scaled(self.bias, x.dtype, self.b_gain), act=self.act, gain=self.out_gain, clamp=self.out_clamp)


class FullyConnectedLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, bias=True, act="linear", gain=1, lrmul=1, bias_init=0):
        super().__init__()
        self.weight, self.w_gain = _make_weight([out_channels, in_channels], gain=gain, lrmul=lrmul)
        self.bias = torch.nn.Parameter(torch.full([out_channels], np.float32(bias_init))) if bias else None
        self.b_gain = lrmul
        self.act = act

    def forward(self, x, _x=None):
        w = _scaled(self.weight, x.dtype, self.w_gain)
        b = _scaled(self.bias, x.dtype, self.b_gain)
        if x.ndim > 2:
            x = x.flatten(1)
        if self.act == "linear" and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return bias_act.bias_act(x.matmul(w.t()), b, act=self.act)


class ResnetLayer(torch.nn.Module):
    def __init__(self, channels, act="linear", lrmul=1, sa=False):
        super().__init__()
        self.fc0 = FullyConnectedLayer(channels, channels, act=act, lrmul=lrmul)
        self.fc1 = FullyConnectedLayer(channels, channels, lrmul=lrmul)

    def forward(self, x, _x):
        shape = x.shape
        h = self.fc1(self.fc0(x.reshape(-1, shape[-1]))).reshape(shape)
        return F.leaky_relu(h + _x, negative_slope=0.2)


class MLP(torch.nn.Module):
    def __init__(self, channels, act, resnet=False, sa=False, pool=False, lrmul=1, **sa_kwargs):
        super().__init__()
        self.layers_num = int(len(channels) / 2) if resnet else (len(channels) - 1)
        self.out_layer = FullyConnectedLayer(channels[-1], channels[-1], act=act, lrmul=lrmul)
        self.pool, self.sa = pool, sa
        for idx in range(self.layers_num):
            in_dim, out_dim = channels[idx], channels[idx + 1]
            if sa:
                setattr(self, f"sa{idx}", TransformerLayer(dim=in_dim, pos_dim=in_dim, from_dim=in_dim, to_dim=in_dim, **sa_kwargs))
            if resnet:
                assert in_dim == out_dim
                layer = ResnetLayer(in_dim, act=act, lrmul=lrmul)
            else:
                layer = FullyConnectedLayer(in_dim, out_dim, act=act, lrmul=lrmul)
            setattr(self, f"l{idx}", layer)

    def forward(self, x, pos=None, mask=None):
        shape = x.shape
        if x.ndim > 2:
            x = x.flatten(1) if self.pool else x.reshape(-1, shape[-1])
        for idx in range(self.layers_num):
            _x = x
            if self.sa:
                x = getattr(self, f"sa{idx}")(from_tensor=x, to_tensor=x, from_pos=pos, to_pos=pos, att_mask=mask.unsqueeze(1))[0]
            x = getattr(self, f"l{idx}")(x, _x)
        return self.out_layer(x).reshape(*shape[:-1], -1)


class Conv2dLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, act="linear", up=1, down=1,
                 resample_kernel=[1, 3, 3, 1], gain=1):
        super().__init__()
        self.up, self.down, self.kernel_size = up, down, kernel_size
        self.weight, self.w_gain = _make_weight([out_channels, in_channels, kernel_size, kernel_size])
        self.biasAct = BiasActLayer(out_channels, bias, act, gain=gain)
        self.register_buffer("resample_kernel", upfirdn2d.setup_filter(resample_kernel))

    def forward(self, x):
        w = _scaled(self.weight, x.dtype, self.w_gain, reorder=True)
        x = conv2d_resample.conv2d_resample(x=x, w=w, f=self.resample_kernel, up=self.up, down=self.down,
                                            padding=self.kernel_size // 2, flip_weight=(self.up == 1))
        return self.biasAct(x)


@torch_misc.profiled_function
def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_kernel=None, demodulate=True,
                     flip_weight=True, fused_modconv=True, modulate=True):
    """reference modulated_conv2d (:252-328): fused = per-sample weights + grouped conv; non-fused = x*s -> conv -> fma."""
    if not modulate:
        x = conv2d_resample.conv2d_resample(x, weight, f=resample_kernel, up=up, padding=padding, flip_weight=flip_weight)
        return x.add_(noise) if noise is not None else x
    batch = int(x.shape[0])
    oc, ic, kh, kw = weight.shape
    torch_misc.assert_shape(x, [batch, ic, None, None])
    torch_misc.assert_shape(styles, [batch, ic])
    w = d = None
    if demodulate or fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(batch, 1, -1, 1, 1)
    if demodulate:
        d = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if fused_modconv:
        if demodulate:
            w = w * d.reshape(batch, -1, 1, 1, 1)
        x = conv2d_resample.conv2d_resample(x=x.reshape(1, -1, *x.shape[2:]), w=w.to(x.dtype).reshape(-1, ic, kh, kw),
                                            f=resample_kernel, up=up, down=down, padding=padding, groups=batch,
                                            flip_weight=flip_weight)
        x = x.reshape(batch, -1, *x.shape[2:])
        return x.add_(noise) if noise is not None else x
    x = x * styles.reshape(batch, -1, 1, 1)
    x = conv2d_resample.conv2d_resample(x=x, w=weight, f=resample_kernel, up=up, down=down, padding=padding, flip_weight=flip_weight)
    if demodulate and noise is not None:
        return fma.fma(x, d.reshape(batch, -1, 1, 1), noise)
    if demodulate:
        return x * d.reshape(batch, -1, 1, 1)
    return x.add_(noise) if noise is not None else x


def att_norm(x, num, integration, norm):
    if norm is None:
        return x
    shape = x.shape
    x = x.reshape([-1, num] + list(shape[1:])).to(float_dtype())
    axis = 1 if norm == "instance" else 2
    if integration in ["add", "both"]:
        x = x - x.mean(dim=axis, keepdim=True)
    if integration in ["mul", "both"]:
        x = x * torch.rsqrt(torch.square(x).mean(dim=axis, keepdim=True) + 1e-8)
    return x.reshape(shape)


def random_dp_binary(shape, dropout, training, device):
    if not training or dropout == 0.0:
        return torch.ones(shape, device=device)
    return torch.rand(shape, device=device) >= dropout


def get_sinusoidal_encoding(size, dim, num=2):
    """reference :406-440, two-direction case."""
    if num != 2:
        raise NotImplementedError("sinusoidal encoding with pos_directions_num != 2")
    c = torch.linspace(-1.0, 1.0, size).unsqueeze(-1)
    i = torch.arange(int(dim / 4)).to(float_dtype())
    sin = torch.sin(c / (torch.pow(10000.0, 4 * i / dim)))
    cos = torch.cos(c / (torch.pow(10000.0, 4 * i / dim)))
    return torch.cat([sin.unsqueeze(0).tile([size, 1, 1]), cos.unsqueeze(0).tile([size, 1, 1]),
                      sin.unsqueeze(1).tile([1, size, 1]), cos.unsqueeze(1).tile([1, size, 1])], dim=-1)


class TransformerLayer(torch.nn.Module):
    """Bipartite / duplex attention layer (reference :558-822): 'from' elements (image grid or latents) attend to
    'to' elements (latents); the result modulates the normalised 'from' tensor."""

    def __init__(self, dim, pos_dim, from_len, to_len, from_dim, to_dim, from_gate=False, to_gate=False, num_heads=1,
                 attention_dropout=0.12, integration="add", norm=None, kmeans=False, kmeans_iters=1, iterative=False, **_kwargs):
        super().__init__()
        if from_gate or to_gate:
            raise NotImplementedError("gated attention (--ltnt-gate/--img-gate)")
        if kmeans and iterative:
            raise NotImplementedError("iterative (non-parametric) centroids")
        if kmeans and kmeans_iters != 1:
            # iterations after the first re-estimate the centroids from the assignments (reference networks.py:776-789); only the
            # GANformer default (one iteration) is built -- refuse instead of silently producing different images
            raise NotImplementedError("kmeans_iters = %r (only 1 is built)" % (kmeans_iters,))
        self.dim, self.pos_dim = dim, pos_dim
        self.from_len, self.to_len, self.from_dim, self.to_dim = from_len, to_len, from_dim, to_dim
        self.num_heads, self.size_head = num_heads, int(dim / num_heads)
        self.att_dp = torch.nn.Dropout(p=attention_dropout / 2)
        self.norm, self.integration = norm, integration
        self.parametric = not iterative
        self.centroid_dim = 2 * self.size_head
        self.kmeans, self.kmeans_iters = kmeans, kmeans_iters
        self.to_queries = FullyConnectedLayer(from_dim, dim)
        self.to_keys = FullyConnectedLayer(to_dim, dim)
        self.to_values = FullyConnectedLayer(to_dim, dim)
        self.from_pos_map = FullyConnectedLayer(pos_dim, dim)
        self.to_pos_map = FullyConnectedLayer(pos_dim, dim)
        self.modulation = FullyConnectedLayer(dim, (2 * dim) if integration == "both" else dim)
        if kmeans:
            self.att_weight = torch.nn.Parameter(torch.ones(num_heads, 1, self.centroid_dim))
            self.centroids = torch.nn.Parameter(torch.randn([1, num_heads, to_len, self.centroid_dim]))

    def _heads(self, x, n):
        return x.reshape(-1, n, self.num_heads, x.shape[-1] // self.num_heads).permute(0, 2, 1, 3)

    def forward(self, from_tensor, to_tensor, from_pos, to_pos, att_vars=None, att_mask=None):
        from_shape = from_tensor.shape
        ft = from_tensor.reshape(-1, from_shape[-1])
        tt = to_tensor.reshape(-1, to_tensor.shape[-1])
        bsz = ft.shape[0] // self.from_len
        q0 = q = self.to_queries(ft)
        k = self.to_keys(tt)
        v = self.to_values(tt)
        if from_pos is not None:
            q = q + self.from_pos_map(from_pos.reshape(-1, self.pos_dim).tile([bsz, 1]))
        if to_pos is not None:
            k = k + self.to_pos_map(to_pos.reshape(-1, self.pos_dim).tile([bsz, 1]))
        if self.kmeans:
            fe = self._heads(torch.cat([q0, q - q0], dim=-1), self.from_len)
            scores = (fe * self.att_weight).matmul(self.centroids.tile([bsz, 1, 1, 1]).permute(0, 1, 3, 2))
        else:
            scores = self._heads(q, self.from_len).matmul(self._heads(k, self.to_len).permute(0, 1, 3, 2))
        scores = scores / math.sqrt(float(self.size_head))
        if att_mask is not None:
            scores = scores + (1 - att_mask.unsqueeze(1).to(float_dtype())) * -10000.0
        probs = F.softmax(scores, dim=-1)
        if self.training and self.att_dp.p > 0:
            rows = list(probs.shape); rows[-2] = 1
            probs = self.att_dp(torch.ones_like(probs)) * probs
            probs = self.att_dp(torch.ones(rows, device=probs.device)) * probs
        to_from = None
        if self.kmeans:
            to_from = (probs / (probs.sum(dim=-2, keepdim=True) + 1e-8)).permute(0, 1, 3, 2)
        control = probs.matmul(self._heads(v, self.to_len)).permute(0, 2, 1, 3).reshape(-1, self.dim)
        x = att_norm(ft, self.from_len, self.integration, self.norm)
        control = self.modulation(control)
        gain = bias = control
        if self.integration == "both":
            gain, bias = torch.split(control, 2, dim=-1)
        if self.integration != "add":
            x = x * (gain + 1)
        if self.integration != "mul":
            x = x + bias
        return x.reshape(from_shape), probs, {"centroid_assignments": to_from}


class MappingNetwork(torch.nn.Module):
    def __init__(self, z_dim=512, c_dim=0, w_dim=512, k=1, num_broadcast=None, num_layers=8, embed_dim=None, layer_dim=None,
                 act="lrelu", lrmul=0.01, w_avg_beta=0.995, transformer=False, resnet=False, shared=False, ltnt2ltnt=False,
                 ltnt_gate=False, normalize_global=True, use_pos=False, **transformer_kwargs):
        super().__init__()
        if c_dim > 0:
            raise NotImplementedError("conditional generator (c_dim > 0)")
        self.z_dim, self.c_dim, self.w_dim, self.k = z_dim, c_dim, w_dim, k
        self.num_broadcast, self.num_layers, self.w_avg_beta = num_broadcast, num_layers, w_avg_beta
        self.normalize_global, self.use_pos, self.transformer = normalize_global, use_pos, transformer
        layer_dim = layer_dim or w_dim
        sa_kwargs = {"sa": ltnt2ltnt and not shared, "pool": shared, "from_len": k - 1, "to_len": k - 1,
                     "from_gate": ltnt_gate, "to_gate": ltnt_gate}
        sa_kwargs.update(transformer_kwargs)
        layers = [layer_dim] * (num_layers - 1) + [w_dim]
        self.global_mlp = MLP([z_dim] + layers, act=act, resnet=resnet, lrmul=lrmul)
        if transformer:
            self.mlp = MLP([z_dim] + layers, act=act, resnet=resnet, lrmul=lrmul, **sa_kwargs)
        if num_broadcast is not None and w_avg_beta is not None:
            self.register_buffer("w_avg", torch.zeros([w_dim]))

    def forward(self, z, c, pos=None, mask=None, truncation_psi=1, truncation_cutoff=None, skip_w_avg_update=False):
        torch_misc.assert_shape(z, [None, self.k, self.z_dim])
        if self.transformer:
            z, g = torch.split(z, [self.k - 1, 1], dim=1)
            if self.normalize_global:
                g = normalize(g)
        z = normalize(z)
        x = self.global_mlp(g if self.transformer else z)
        if self.transformer:
            x = torch.cat([self.mlp(z, pos=pos if self.use_pos else None, mask=mask), x], dim=1)
        if self.w_avg_beta is not None and self.training and not skip_w_avg_update:
            self.w_avg.copy_(x.detach().mean(dim=(0, 1)).lerp(self.w_avg, self.w_avg_beta))
        if self.num_broadcast is not None:
            x = x.unsqueeze(2).repeat([1, 1, self.num_broadcast, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_broadcast is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :, :truncation_cutoff] = self.w_avg.lerp(x[:, :, :truncation_cutoff], truncation_psi)
        return x


class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, y_dim, k, out_resolution, kernel_size=3, up=1, local_noise=True, bias=True,
                 act="lrelu", resample_kernel=[1, 3, 3, 1], gain=1, style=True, transformer=False, use_pos=False,
                 ltnt_gate=False, img_gate=False, **transformer_kwargs):
        super().__init__()
        self.affine = FullyConnectedLayer(y_dim, in_channels, bias_init=1)
        self.weight, self.w_gain = _make_weight([out_channels, in_channels, kernel_size, kernel_size])
        self.biasAct = BiasActLayer(out_channels, act=act, gain=gain) if bias else None
        self.style, self.kernel_size = style, kernel_size
        self.out_res, self.up = out_resolution, up
        self.in_res = out_resolution // up
        self.register_buffer("resample_kernel", upfirdn2d.setup_filter(resample_kernel))
        self.local_noise = local_noise
        if local_noise:
            self.register_buffer("noise_const", torch.randn([self.out_res, self.out_res]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.transformer, self.use_pos = None, use_pos
        if transformer:
            pos_dim = transformer_kwargs.get("pos_dim") or y_dim
            transformer_kwargs["pos_dim"] = pos_dim
            if transformer_kwargs.get("pos_type", "sinus") != "sinus":
                raise NotImplementedError("only the sinusoidal grid encoding is built")
            self.register_buffer("grid_pos", get_sinusoidal_encoding(out_resolution, pos_dim, transformer_kwargs.get("pos_directions_num", 2)))
            kwargs = {"from_len": self.out_res * self.out_res, "to_len": k - 1, "from_dim": out_channels, "to_dim": y_dim,
                      "from_gate": img_gate, "to_gate": ltnt_gate}
            kwargs.update(transformer_kwargs)
            self.transformer = TransformerLayer(dim=out_channels, **kwargs)

    def forward(self, x, y, att_vars=None, pos=None, mask=None, noise_mode="random", fused_modconv=True):
        assert noise_mode in ["random", "const", "none"]
        torch_misc.assert_shape(x, [None, self.weight.shape[1], self.in_res, self.in_res])
        att_map, noise = None, None
        if self.local_noise and noise_mode != "none":
            if noise_mode == "random":
                noise = torch.randn([x.shape[0], 1, self.out_res, self.out_res], device=x.device)
            else:
                noise = self.noise_const
            noise = noise * self.noise_strength
        x = modulated_conv2d(x=x, weight=self.weight * self.w_gain, styles=self.affine(get_global(y)), modulate=self.style,
                             up=self.up, padding=self.kernel_size // 2, resample_kernel=self.resample_kernel,
                             flip_weight=(self.up == 1), fused_modconv=fused_modconv)
        if self.transformer is not None:
            shape = x.shape
            x = x.reshape(shape[0], shape[1], -1).permute(0, 2, 1)
            x, att_map, att_vars = self.transformer(from_tensor=x, to_tensor=get_components(y), from_pos=self.grid_pos,
                                                    to_pos=pos if self.use_pos else None, att_vars=att_vars,
                                                    att_mask=mask.unsqueeze(1))
            x = x.permute(0, 2, 1).reshape(shape)
        if noise is not None:
            x = x + noise
        if self.biasAct:
            x = self.biasAct(x)
        return x, att_map, att_vars


class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, y_dim, kernel_size=1, style=True):
        super().__init__()
        self.affine = FullyConnectedLayer(y_dim, in_channels, bias_init=1)
        self.weight, self.w_gain = _make_weight([out_channels, in_channels, kernel_size, kernel_size])
        self.biasAct = BiasActLayer(out_channels)
        self.style = style

    def forward(self, x, y, fused_modconv):
        styles = self.affine(get_global(y))
        weight = self.weight
        if self.style:
            styles = styles * self.w_gain
        else:
            weight = self.weight * self.w_gain
        x = modulated_conv2d(x=x, weight=weight, styles=styles, modulate=self.style, demodulate=False, fused_modconv=fused_modconv)
        return self.biasAct(x).to(float_dtype())


class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture="skip",
                 resample_kernel=[1, 3, 3, 1], latent_stem=False, style=True, **layer_kwargs):
        assert architecture in ["orig", "skip", "resnet"]
        super().__init__()
        if latent_stem:
            raise NotImplementedError("latent stem")
        self.in_channels, self.img_channels, self.res, self.w_dim = in_channels, img_channels, resolution, w_dim
        self.stem, self.is_last, self.architecture = (in_channels == 0), is_last, architecture
        self.register_buffer("resample_kernel", upfirdn2d.setup_filter(resample_kernel))
        self.num_conv, self.num_torgb = 0, 0
        if self.stem:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, out_resolution=resolution, up=2, resample_kernel=resample_kernel,
                                        y_dim=w_dim, style=style, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, out_resolution=resolution,
                                    gain=1 if self.stem else get_gain(architecture), y_dim=w_dim, style=style, **layer_kwargs)
        self.num_conv += 1
        if is_last or architecture == "skip":
            self.torgb = ToRGBLayer(out_channels, img_channels, y_dim=w_dim, style=style)
            self.num_torgb += 1
        if (not self.stem) and architecture == "resnet":
            self.skip = Conv2dLayer(in_channels, out_channels, kernel_size=1, bias=False, up=2,
                                    resample_kernel=resample_kernel, gain=get_gain(architecture))
        if is_last:
            last_kwargs = dict(layer_kwargs, transformer=False, bias=False, local_noise=False)
            self.conv_last = SynthesisLayer(out_channels, out_channels, out_resolution=resolution, y_dim=w_dim, style=style, **last_kwargs)
            self.num_conv += 1

    def forward(self, x, img, ws, att_vars, fused_modconv=None, **layer_kwargs):
        torch_misc.assert_shape(ws, [None, None, self.num_conv + self.num_torgb, self.w_dim])
        w_iter = iter(ws.unbind(dim=2))
        if fused_modconv is None:
            fused_modconv = not self.training
        if self.stem:
            x = self.const.unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
        else:
            torch_misc.assert_shape(x, [None, self.in_channels, self.res // 2, self.res // 2])
        x = x.to(float_dtype())
        att_maps = [None, None]
        kw = dict(fused_modconv=fused_modconv, **layer_kwargs)
        if self.stem:
            x, att_maps[0], att_vars = self.conv1(x, next(w_iter), att_vars, **kw)
        elif self.architecture == "resnet":
            y = self.skip(x)
            x, att_maps[0], att_vars = self.conv0(x, next(w_iter), att_vars, **kw)
            x, att_maps[1], att_vars = self.conv1(x, next(w_iter), att_vars, **kw)
            x = y + x
        else:
            x, att_maps[0], att_vars = self.conv0(x, next(w_iter), att_vars, **kw)
            x, att_maps[1], att_vars = self.conv1(x, next(w_iter), att_vars, **kw)
        if img is not None:
            img = upfirdn2d.upsample2d(img, self.resample_kernel)
        if self.is_last:
            x = self.conv_last(x, next(w_iter), **kw)[0]
        if self.is_last or self.architecture == "skip":
            y = self.torgb(x, next(w_iter), fused_modconv=fused_modconv)
            img = img + y if img is not None else y
        return x, img, att_maps, att_vars


class SynthesisNetwork(torch.nn.Module):
    def __init__(self, w_dim, k, img_resolution, img_channels, channel_base=32 << 10, channel_max=512, transformer=False,
                 start_res=0, end_res=20, **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim, self.k, self.img_res, self.img_channels = w_dim, k, img_resolution, img_channels
        self.block_resolutions = [2 ** i for i in range(2, int(np.log2(img_resolution)) + 1)]
        self.architecture = block_kwargs.get("architecture", "skip")
        self.end_res, self.start_res, self.use_transformer = end_res, start_res, transformer
        nch = lambda res: min(channel_base // res, channel_max)
        self.num_ws = 0
        for res in self.block_resolutions:
            is_last = res == self.img_res
            use_tr = transformer and np.log2(res) >= start_res and np.log2(res) < end_res
            block = SynthesisBlock(nch(res // 2) if res > 4 else 0, nch(res), w_dim=w_dim, k=k, resolution=res,
                                   img_channels=img_channels, is_last=is_last, transformer=use_tr, **block_kwargs)
            self.num_ws += block.num_conv
            if is_last:
                self.num_ws += block.num_torgb
            setattr(self, f"b{res}", block)
        self.engine = "ops"       # "ops" | "tc"; see module docstring
        self._tc = None

    def list2tensor(self, att_list, device):
        att_list = [a for a in att_list if a is not None]
        if len(att_list) == 0:
            return torch.zeros([1], device=device)
        maps = []
        for a in att_list:
            heads = a.shape[1]
            s = int(math.sqrt(int(a.shape[2])))
            a = a.reshape(-1, s, s, self.k - 1).permute(0, 3, 1, 2)
            if s < self.img_res:
                factor = int(self.img_res / s)
                a = upfirdn2d.upsample2d(a, f=upfirdn2d.setup_filter([1] * factor, device=a.device), up=factor)
            maps.append(a.reshape(-1, heads, self.k - 1, self.img_res, self.img_res))
        return torch.stack(maps, dim=1).permute(0, 3, 1, 2, 4, 5)

    def forward(self, ws, return_att_maps=None, **block_kwargs):
        """return_att_maps: the reference default is True (networks.py:1244).  Left unset (None) it means True on the ops engine and
        False on the tc engine, where the maps are 738 MB per image at 1024^2 -- pass True to get them there too;
        `torch.zeros([1])` is the reference's own "no maps" value (:1224-1225)."""
        torch_misc.assert_shape(ws, [None, self.k, self.num_ws, self.w_dim])
        if self.engine == "tc":
            from .. import engine as _engine
            if self._tc is None:
                self._tc = _engine.SynthesisEngine(self)
            img = self._tc(ws, want_probs=bool(return_att_maps), **block_kwargs)
            if not return_att_maps:
                return img, torch.zeros([1], device=ws.device)
            probs = [p.reshape(p.shape[0], 1, p.shape[1], p.shape[2]) for p in self._tc.last_probs]
            return img, self.list2tensor(probs, ws.device)
        if return_att_maps is None:
            return_att_maps = True
        ws = ws.to(torch.float32)
        block_ws, w_idx = [], 0
        for res in self.block_resolutions:
            block = getattr(self, f"b{res}")
            block_ws.append(ws.narrow(2, w_idx, block.num_conv + block.num_torgb))
            w_idx += block.num_conv
        x, img, att_maps = None, None, []
        att_vars = {"centroid_assignments": None}
        for res, cur_ws in zip(self.block_resolutions, block_ws):
            x, img, maps, att_vars = getattr(self, f"b{res}")(x, img, cur_ws, att_vars, **block_kwargs)
            att_maps += maps
        att = self.list2tensor(att_maps, ws.device) if return_att_maps else torch.zeros([1], device=ws.device)
        return img, att


class Generator(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, k, img_resolution, img_channels, component_dropout=0.0, mapping_kwargs={},
                 synthesis_kwargs={}, **_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.k = z_dim, c_dim, w_dim, k
        self.img_resolution, self.img_channels, self.component_dropout = img_resolution, img_channels, component_dropout
        self.input_shape, self.cond_shape = [None, k, z_dim], [None, c_dim]
        self.pos = torch.nn.Parameter(torch.rand([k - 1, w_dim])) if k > 1 else None
        self.synthesis = SynthesisNetwork(w_dim=w_dim, k=k, img_resolution=img_resolution, img_channels=img_channels,
                                          **dict(synthesis_kwargs))
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, k=k, num_broadcast=self.num_ws, **dict(mapping_kwargs))

    def forward(self, z=None, c=None, ws=None, truncation_psi=1, truncation_cutoff=None, return_img=True, return_att=False,
                return_ws=False, subnet=None, **synthesis_kwargs):
        return_tensor = False
        if subnet is not None:
            return_ws, return_img, return_att, return_tensor = (subnet == "mapping"), (subnet == "synthesis"), False, True
        _input = z if z is not None else ws
        mask = random_dp_binary([_input.shape[0], self.k - 1], self.component_dropout, self.training, _input.device)
        if ws is None:
            ws = self.mapping(z, c, pos=self.pos, mask=mask, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        torch_misc.assert_shape(ws, [None, self.k, self.num_ws, self.w_dim])
        ret = ()
        if return_img or return_att:
            img, att_maps = self.synthesis(ws, pos=self.pos, mask=mask, return_att_maps=return_att, **synthesis_kwargs)
            if return_img:
                ret += (img,)
            if return_att:
                ret += (att_maps,)
        if return_ws:
            ret += (ws,)
        return ret[0] if return_tensor else ret


def ganformer_default_kwargs(img_resolution, channel_base=32768, channel_max=512, architecture="resnet"):
    """The GANformer-default generator (reference run_network.py:61-77, :230-283): k = 16 local + 1 global components of
    32 dims, duplex attention at resolutions 4..128, resnet synthesis blocks."""
    tr = dict(num_heads=1, attention_dropout=0.12, use_pos=True, ltnt_gate=False)
    return dict(
        z_dim=32, c_dim=0, w_dim=32, k=17, img_resolution=img_resolution, img_channels=3, component_dropout=0.0,
        mapping_kwargs=dict(num_layers=8, layer_dim=None, resnet=True, shared=False, ltnt2ltnt=True, transformer=True, **tr),
        synthesis_kwargs=dict(channel_base=channel_base, channel_max=channel_max, architecture=architecture, style=True,
                              latent_stem=False, local_noise=True, transformer=True, start_res=0, end_res=8, norm="layer",
                              integration="mul", img_gate=False, iterative=False, kmeans=True, kmeans_iters=1, pos_dim=None,
                              pos_type="sinus", pos_init="uniform", pos_directions_num=2, **tr))

"""Shared helpers for the tests: seeded model construction (no reference tree needed) and case tables."""
import os
import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def randomize(G, seed=1):
    """make every term live (noise_strength is 0 and biases are 0/1 at init)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("noise_strength"):
                p.copy_(torch.randn([], generator=g) * 0.1)
            elif n.endswith("biasAct.bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return G


def build_G(res, seed=0, channel_base=32768, channel_max=512, rand_seed=1):
    from morphganformer_b200.training import networks as N
    torch.manual_seed(seed)
    G = N.Generator(**N.ganformer_default_kwargs(res, channel_base=channel_base, channel_max=channel_max))
    return randomize(G.eval().requires_grad_(False), rand_seed)


def state_dict_cpu(G):
    return {k: v.detach().cpu() for k, v in G.state_dict().items()}


def sd_checksum(sd):
    """order-independent fp64 checksum of a state dict (sum of sum(|x|) and sum(x*idx-weights))."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double().flatten()
        if v.numel():
            w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 977
            tot += float((v * w).sum()) + float(v.abs().sum())
    return tot


def build_vgg_lpips_sd(seed=4):
    """LPIPS-VGG16 parameters: torchvision default random init under `seed` (north_star: random-init only) in the
    reference's naming (net.sliceK.N.*) + the five `lin` weights shipped with the reference (copied into
    tests/golden/lpips_lin_vgg_v0.1.npz by make_golden.py; 1472 floats)."""
    import torchvision
    torch.manual_seed(seed)
    feats = torchvision.models.vgg16(weights=None).features
    sd = {}
    slices = {1: range(0, 4), 2: range(4, 9), 3: range(9, 16), 4: range(16, 23), 5: range(23, 30)}
    for s, idxs in slices.items():
        for i in idxs:
            m = feats[i]
            if isinstance(m, torch.nn.Conv2d):
                sd[f"net.slice{s}.{i}.weight"] = m.weight.detach().clone()
                sd[f"net.slice{s}.{i}.bias"] = m.bias.detach().clone()
    lin = np.load(os.path.join(GOLDEN, "lpips_lin_vgg_v0.1.npz"))
    for kk in range(5):
        sd[f"lin{kk}.model.1.weight"] = torch.from_numpy(lin[f"lin{kk}"]).reshape(1, -1, 1, 1)
    return sd


def build_lpips_sd(net, seed=4):
    """LPIPS alex / squeeze parameters: torchvision default random init under `seed` in the reference's naming + the shipped
    `lin` weights (tests/golden/lpips_lin_{alex,squeeze}_v0.1.npz, written by make_golden_lpips_nets.py)."""
    import torchvision
    if net == "vgg":
        return build_vgg_lpips_sd(seed)
    torch.manual_seed(seed)
    feats = (torchvision.models.alexnet(weights=None) if net == "alex" else torchvision.models.squeezenet1_1(weights=None)).features
    slices = {"alex": [range(0, 2), range(2, 5), range(5, 8), range(8, 10), range(10, 12)],
              "squeeze": [range(0, 2), range(2, 5), range(5, 8), range(8, 10), range(10, 11), range(11, 12), range(12, 13)]}[net]
    sd = {}
    for s, idxs in enumerate(slices, 1):
        for i in idxs:
            for k, v in feats[i].state_dict().items():
                sd[f"net.slice{s}.{i}.{k}"] = v.detach().clone()
    lin = np.load(os.path.join(GOLDEN, f"lpips_lin_{net}_v0.1.npz"))
    for kk in range(len(slices)):
        sd[f"lin{kk}.model.1.weight"] = torch.from_numpy(lin[f"lin{kk}"]).reshape(1, -1, 1, 1)
    return sd


# (name, x shape, filter taps or None, kwargs) -- the calls SURVEY 8a-7 lists plus edge cases the reference ops accept
UPFIRDN_CASES = [
    ("skip_up2", (2, 5, 8, 8), [1, 3, 3, 1], dict(up=2, padding=[2, 1, 2, 1], gain=4)),
    ("post_transpose", (1, 6, 17, 17), [1, 3, 3, 1], dict(padding=[1, 1, 1, 1], gain=4)),
    ("nearest_up2", (2, 16, 4, 4), [1, 1], dict(up=2, padding=[1, 0, 1, 0], gain=4)),
    ("down2", (2, 3, 16, 16), [1, 3, 3, 1], dict(down=2, padding=[1, 1, 1, 1])),
    ("crop_neg_pad", (1, 2, 12, 10), [1, 2, 1], dict(padding=[-1, -2, -1, 0])),
    ("flip", (1, 2, 9, 7), [1, 2, 3, 4], dict(flip_filter=True, padding=[3, 0, 0, 3], gain=2.5)),
    ("odd_sizes_up3_down2", (1, 3, 7, 5), [1, 4, 6, 4, 1], dict(up=3, down=2, padding=[2, 3, 1, 4])),
    ("identity_f_none", (1, 2, 5, 5), None, dict(padding=[1, 1, 0, 0])),
    ("separable_8tap_up8", (1, 2, 4, 4), [1] * 8, dict(up=8, padding=[7, 0, 7, 0], gain=64)),
    ("asym_updown", (1, 2, 6, 9), [1, 3, 3, 1], dict(up=[2, 1], down=[1, 2], padding=[2, 1, 1, 1])),
    ("wide_tile_edges", (1, 1, 70, 130), [1, 3, 3, 1], dict(up=2, padding=[2, 1, 2, 1], gain=4)),
]

BIAS_ACT_ACTS = ["linear", "relu", "lrelu", "tanh", "sigmoid", "elu", "selu", "softplus", "swish"]

# (name, x shape, w shape, kwargs)  -- the six branches of conv2d_resample (reference :99,:105,:111,:117,:137,:142)
RESAMPLE_CASES = [
    ("plain3x3", (2, 4, 9, 9), (6, 4, 3, 3), dict(padding=1)),
    ("plain_noflip", (2, 4, 9, 9), (6, 4, 3, 3), dict(padding=1, flip_weight=False)),
    ("up2_3x3", (2, 4, 8, 8), (6, 4, 3, 3), dict(up=2, padding=1, flip_weight=False, f=[1, 3, 3, 1])),
    ("up2_1x1", (2, 4, 8, 8), (6, 4, 1, 1), dict(up=2, padding=0, flip_weight=False, f=[1, 3, 3, 1])),
    ("down2_3x3", (2, 4, 16, 16), (6, 4, 3, 3), dict(down=2, padding=1, f=[1, 3, 3, 1])),
    ("down2_1x1", (2, 4, 16, 16), (6, 4, 1, 1), dict(down=2, padding=0, f=[1, 3, 3, 1])),
    ("grouped_up2", (1, 8, 8, 8), (12, 4, 3, 3), dict(up=2, padding=1, groups=2, flip_weight=False, f=[1, 3, 3, 1])),
    ("grouped_plain", (1, 8, 7, 7), (12, 4, 3, 3), dict(padding=1, groups=2)),
    ("asym_pad_generic", (1, 3, 8, 8), (5, 3, 3, 3), dict(padding=[2, 0, 1, 0])),
    ("updown", (1, 3, 8, 8), (5, 3, 3, 3), dict(up=2, down=2, padding=1, f=[1, 3, 3, 1])),
]


def case_tensor(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def img_abs_tol(ref):
    """Absolute max-abs image tolerance of the tcgen05 engine in its default fp16 forward storage: 1e-2 (north_star) for images of the
    nominal GAN range |img| <= 2.5; the BASELINE configurations themselves (256^2: range 2.4, 1024^2: range 4.2) are held to 1e-2 absolute in
    tests/test_fullsize_parity_gpu.py.  The narrow 64-channel toy generators of the unit tests produce images of range ~5 with twice the
    relative rounding noise (fewer channels to average over): 2e-2 absolute there (0.4 % of the range)."""
    return 1e-2 if float(ref.detach().abs().max()) <= 2.5 else 2e-2

"""TEST INFRASTRUCTURE ONLY -- imports the real reference (read-only) when it is present.

`/root/reference` exists only in the build container, never on the GPU box.  This module is
used (a) by tests/golden/make_golden.py to generate the committed golden vectors and (b) by
the container-only tests that pin the oracle restatement (oracle/ganformer.py, oracle/ops.py,
oracle/lpips_ref.py) against the real reference code.  Nothing in the product imports it.

Shims (the reference tree itself is never modified):
  1. stub modules for cosmetic imports: termcolor, seaborn (misc.py:4-5, torch_utils/misc.py:7),
     skimage / IPython (lpips/__init__.py:7, lpips/networks_basic.py:11-12, lpips/dist_model.py:13-19)
  2. class-level `TransformerLayer.dim` property, because `self.dim = dim` is commented out at
     training/networks.py:581 while :616-617 and :814 read it.
  3. LPIPS is built as PNetLin(pnet_rand=True) + the shipped lin weights lpips/weights/v0.1/vgg.pth
     (no network, torchvision default random init under a fixed seed).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MGF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "training", "networks.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a namespace with the reference modules: networks, upfirdn2d, bias_act, conv2d_resample, fma, lpips_nb."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _stub("termcolor", colored=lambda s, *a, **k: str(s))
    _stub("seaborn", color_palette=lambda *a, **k: [])
    sk = _stub("skimage")
    sk.measure = _stub("skimage.measure", compare_ssim=None)
    sk.color = _stub("skimage.color")
    sk.transform = _stub("skimage.transform")
    sk.io = _stub("skimage.io")
    _stub("IPython", embed=lambda *a, **k: None)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # the product package mirrors `torch_utils` / `training` names inside morphganformer_b200/, not at
    # top level, so there is no clash with the reference's top-level packages.
    from training import networks  # noqa
    from torch_utils.ops import upfirdn2d, bias_act, conv2d_resample, fma, conv2d_gradfix  # noqa
    networks.TransformerLayer.dim = property(lambda self: self.size_head * self.num_heads)
    from lpips import networks_basic as lpips_nb  # noqa
    _loaded.update(networks=networks, upfirdn2d=upfirdn2d, bias_act=bias_act,
                   conv2d_resample=conv2d_resample, fma=fma, conv2d_gradfix=conv2d_gradfix,
                   lpips_nb=lpips_nb, root=REF_ROOT)
    return types.SimpleNamespace(**_loaded)


def ganformer_kwargs(res, channel_base=32768, channel_max=512, architecture="resnet"):
    """GANformer-default generator kwargs (run_network.py:61-77, :230-283; loader.py:104-154)."""
    tr = dict(num_heads=1, attention_dropout=0.12, use_pos=True, ltnt_gate=False)
    return dict(
        z_dim=32, c_dim=0, w_dim=32, k=17, img_resolution=res, img_channels=3, component_dropout=0.0,
        mapping_kwargs=dict(num_layers=8, layer_dim=None, resnet=True, shared=False, ltnt2ltnt=True,
                            transformer=True, **tr),
        synthesis_kwargs=dict(channel_base=channel_base, channel_max=channel_max, architecture=architecture,
                              style=True, latent_stem=False, local_noise=True, transformer=True,
                              start_res=0, end_res=8, norm="layer", integration="mul", img_gate=False,
                              iterative=False, kmeans=True, kmeans_iters=1, pos_dim=None, pos_type="sinus",
                              pos_init="uniform", pos_directions_num=2, **tr),
    )


def build_generator(res, seed=0, **kw):
    import torch
    ref = load()
    torch.manual_seed(seed)
    G = ref.networks.Generator(**ganformer_kwargs(res, **kw)).eval().requires_grad_(False)
    return G


def build_lpips(seed=4, net_type="vgg"):
    import torch
    ref = load()
    torch.manual_seed(seed)
    net = ref.lpips_nb.PNetLin(pnet_type=net_type, pnet_rand=True, use_dropout=True, version="0.1", lpips=True).eval()
    sd = torch.load(os.path.join(REF_ROOT, "lpips", "weights", "v0.1", net_type + ".pth"), map_location="cpu")
    net.load_state_dict(sd, strict=False)
    return net.requires_grad_(False)

"""LPIPS with the AlexNet and SqueezeNet-1.1 backbones, and VGG16 in exact fp32, on the ops path.

Mirrors lpips.PerceptualLoss(model='net-lin', net='alex' | 'squeeze') of the reference (lpips/__init__.py:13-41 ->
networks_basic.py:27-92; backbones pretrained_networks.py:6-95; linear heads weights/v0.1/{alex,squeeze}.pth) and the optional
area-downsample to 256^2 the v1 scripts apply first (projection_example_v1.py:150-155).

Every convolution (11x11 stride 4, 5x5, 3x3, 1x1, stride 2) runs on the library's exact-fp32 direct-convolution kernels through
`conv2d_gradfix.conv2d` (mgf_conv2d_fwd/dgrad/wgrad_f32), bias + ReLU on `bias_act` (mgf_bias_act), so the distance is
differentiable to any order like the reference's.  Max-pooling (3x3 stride 2, `ceil_mode` for SqueezeNet) and the per-tap
normalise / difference / 1x1 `lin` / mean tail are PyTorch tensor ops -- this path is the functional drop-in for the two cheaper
backbones (AlexNet is ~10x cheaper than VGG16), not the tuned tcgen05 path `lpips_engine.LpipsEngine` provides for VGG16.
CUDA tensors only (the ops raise on CPU tensors: no fallback)."""
import torch
import torch.nn.functional as F

from .torch_utils.ops import bias_act, conv2d_gradfix

CHNS = {"alex": [64, 192, 384, 256, 256], "squeeze": [64, 128, 256, 384, 384, 512, 512], "vgg": [64, 128, 256, 512, 512]}
# VGG16 conv indices per slice (torchvision features numbering, reference pretrained_networks.py:108-117); a 2x2 max-pool opens slices 2..5
_VGG_SLICES = [(1, (0, 2)), (2, (5, 7)), (3, (10, 12, 14)), (4, (17, 19, 21)), (5, (24, 26, 28))]
_SHIFT = (-.030, -.088, -.188)
_SCALE = (.458, .448, .450)


def _conv_relu(x, w, b, stride=1, padding=0):
    return bias_act.bias_act(conv2d_gradfix.conv2d(x, w, None, stride=stride, padding=padding), b, act="relu", gain=1)    # bias_act's default relu gain is sqrt(2) (StyleGAN convention); VGG/Alex/Squeeze use plain ReLU


class LpipsNet(torch.nn.Module):
    def __init__(self, state_dict, net="alex", downsample_to=None):
        """state_dict: the reference PNetLin naming (net.sliceK.N[.squeeze|.expand1x1|.expand3x3].weight/bias, linK.model.1.weight).
        downsample_to: if set (e.g. 256), inputs larger than that are box-averaged by the integer factor first."""
        super().__init__()
        if net not in CHNS:
            raise NotImplementedError("net must be 'alex', 'squeeze' or 'vgg'")
        self.net_type, self.downsample_to = net, downsample_to
        for k, v in state_dict.items():
            if k.startswith("net.") or k.startswith("lin"):
                self.register_buffer(k.replace(".", "__"), v.detach().float().clone())
        self.register_buffer("shift", torch.tensor(_SHIFT).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor(_SCALE).view(1, 3, 1, 1))

    def _p(self, name):
        return getattr(self, name.replace(".", "__"))

    def _cv(self, h, s, i, **kw):
        return _conv_relu(h, self._p(f"net.slice{s}.{i}.weight"), self._p(f"net.slice{s}.{i}.bias"), **kw)

    def _fire(self, h, s, i):
        p = f"net.slice{s}.{i}."
        q = _conv_relu(h, self._p(p + "squeeze.weight"), self._p(p + "squeeze.bias"))
        a = _conv_relu(q, self._p(p + "expand1x1.weight"), self._p(p + "expand1x1.bias"))
        b = _conv_relu(q, self._p(p + "expand3x3.weight"), self._p(p + "expand3x3.bias"), padding=1)
        return torch.cat([a, b], 1)

    def features(self, x):
        if self.net_type == "vgg":                        # exact-fp32 VGG16 trunk (the tuned 16-bit tcgen05 path is lpips_engine.LpipsEngine)
            feats, h = [], x
            for s, idxs in _VGG_SLICES:
                if s > 1:
                    h = F.max_pool2d(h, 2, 2)
                for i in idxs:
                    h = self._cv(h, s, i, padding=1)
                feats.append(h)
            return feats
        if self.net_type == "alex":                       # torchvision alexnet.features[0:12]
            f1 = self._cv(x, 1, 0, stride=4, padding=2)
            f2 = self._cv(F.max_pool2d(f1, 3, 2), 2, 3, padding=2)
            f3 = self._cv(F.max_pool2d(f2, 3, 2), 3, 6, padding=1)
            f4 = self._cv(f3, 4, 8, padding=1)
            return [f1, f2, f3, f4, self._cv(f4, 5, 10, padding=1)]
        pool = lambda h: F.max_pool2d(h, 3, 2, ceil_mode=True)     # squeezenet1_1.features[0:13]
        f1 = self._cv(x, 1, 0, stride=2)
        f2 = self._fire(self._fire(pool(f1), 2, 3), 2, 4)
        f3 = self._fire(self._fire(pool(f2), 3, 6), 3, 7)
        f4 = self._fire(pool(f3), 4, 9)
        f5 = self._fire(f4, 5, 10)
        f6 = self._fire(f5, 6, 11)
        return [f1, f2, f3, f4, f5, f6, self._fire(f6, 7, 12)]

    def _pre(self, img):
        if self.downsample_to and img.shape[2] > self.downsample_to:
            f = img.shape[2] // self.downsample_to
            b, c, h, w = img.shape
            img = img.reshape(b, c, h // f, f, w // f, f).mean([3, 5])
        return (img.float() - self.shift) / self.scale

    def forward(self, pred, target, normalize=False):
        """[N,3,H,W] in [-1,1] (or [0,1] with normalize=True) -> [N,1,1,1]"""
        if normalize:
            pred, target = 2 * pred - 1, 2 * target - 1
        f0, f1 = self.features(self._pre(pred)), self.features(self._pre(target))
        val = 0
        for kk, (a, b) in enumerate(zip(f0, f1)):
            na = a / (a.square().sum(1, keepdim=True).sqrt() + 1e-10)
            nb = b / (b.square().sum(1, keepdim=True).sqrt() + 1e-10)
            lin = self._p(f"lin{kk}.model.1.weight").reshape(1, -1, 1, 1)
            val = val + ((na - nb).square() * lin).sum(1, keepdim=True).mean([2, 3], keepdim=True)
        return val


def PerceptualLoss(state_dict, model="net-lin", net="vgg", **kw):
    """Factory with the reference's call shape (lpips/__init__.py:13-24): VGG16 -> the tcgen05 engine, alex / squeeze -> LpipsNet."""
    if model != "net-lin":
        raise NotImplementedError("only model='net-lin' (the LPIPS linear-head variant) is built")
    if net in ("vgg", "vgg16") and not kw.get("exact_fp32"):
        from .lpips_engine import PerceptualLoss as _Vgg
        return _Vgg(state_dict, model=model, net=net)
    return LpipsNet(state_dict, net="vgg" if net == "vgg16" else net, downsample_to=kw.get("downsample_to"))

"""TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference's LPIPS-VGG16 distance (lpips/networks_basic.py:64-92 PNetLin.forward,
ScalingLayer :94-101, NetLinLayer :104-111, spatial_average :17-18; lpips/__init__.py:44-46
normalize_tensor; lpips/pretrained_networks.py:97-135 vgg16 slices) on a state_dict with the reference's
parameter names (net.sliceK.N.weight/bias, linK.model.1.weight).  Dropout is identity in eval mode.
Pinned by tests/test_oracle_vs_reference.py (container only) and tests/golden/lpips_golden.npz.
"""
import torch
import torch.nn.functional as F

# (slice, torchvision features index, pool-before?)  pretrained_networks.py:108-117
VGG_LAYERS = [
    [(1, 0), (1, 2)],
    [(2, 5), (2, 7)],
    [(3, 10), (3, 12), (3, 14)],
    [(4, 17), (4, 19), (4, 21)],
    [(5, 24), (5, 26), (5, 28)],
]
SHIFT = [-.030, -.088, -.188]
SCALE = [.458, .448, .450]


def vgg_features(sd, x):
    feats = []
    h = x
    for si, layers in enumerate(VGG_LAYERS):
        if si > 0:
            h = F.max_pool2d(h, 2, 2)
        for (s, idx) in layers:
            h = F.relu(F.conv2d(h, sd[f"net.slice{s}.{idx}.weight"].to(h.dtype), sd[f"net.slice{s}.{idx}.bias"].to(h.dtype), padding=1))
        feats.append(h)
    return feats


def normalize_tensor(t, eps=1e-10):
    return t / (torch.sqrt(torch.sum(t ** 2, dim=1, keepdim=True)) + eps)


def lpips(sd, pred, target):
    """returns [N,1,1,1] like PerceptualLoss.forward (lpips/__init__.py:26-41)."""
    shift = torch.tensor(SHIFT, dtype=pred.dtype)[None, :, None, None]
    scale = torch.tensor(SCALE, dtype=pred.dtype)[None, :, None, None]
    f0 = vgg_features(sd, (pred - shift) / scale)
    f1 = vgg_features(sd, (target - shift) / scale)
    val = 0
    for kk in range(5):
        d = (normalize_tensor(f0[kk]) - normalize_tensor(f1[kk])) ** 2
        r = F.conv2d(d, sd[f"lin{kk}.model.1.weight"].to(pred.dtype)).mean([2, 3], keepdim=True)
        val = val + r
    return val

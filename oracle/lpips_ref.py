"""TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference's LPIPS distance (VGG16; AlexNet and SqueezeNet-1.1 backbones for SURVEY.md 8f rank 2) (lpips/networks_basic.py:64-92 PNetLin.forward,
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


def alex_features(sd, x):
    """lpips/pretrained_networks.py:57-95 (torchvision alexnet.features[0:12], taps after every ReLU'd conv)."""
    cv = lambda h, s, i, **kw: F.relu(F.conv2d(h, sd[f"net.slice{s}.{i}.weight"].to(h.dtype), sd[f"net.slice{s}.{i}.bias"].to(h.dtype), **kw))
    f1 = cv(x, 1, 0, stride=4, padding=2)
    f2 = cv(F.max_pool2d(f1, 3, 2), 2, 3, padding=2)
    f3 = cv(F.max_pool2d(f2, 3, 2), 3, 6, padding=1)
    f4 = cv(f3, 4, 8, padding=1)
    f5 = cv(f4, 5, 10, padding=1)
    return [f1, f2, f3, f4, f5]


def squeeze_features(sd, x):
    """lpips/pretrained_networks.py:6-55 (torchvision squeezenet1_1.features[0:13]; Fire = squeeze 1x1 -> cat(expand 1x1, expand 3x3))."""
    def fire(h, s, i):
        p = f"net.slice{s}.{i}."
        q = F.relu(F.conv2d(h, sd[p + "squeeze.weight"].to(h.dtype), sd[p + "squeeze.bias"].to(h.dtype)))
        a = F.relu(F.conv2d(q, sd[p + "expand1x1.weight"].to(h.dtype), sd[p + "expand1x1.bias"].to(h.dtype)))
        b = F.relu(F.conv2d(q, sd[p + "expand3x3.weight"].to(h.dtype), sd[p + "expand3x3.bias"].to(h.dtype), padding=1))
        return torch.cat([a, b], 1)
    pool = lambda h: F.max_pool2d(h, 3, 2, ceil_mode=True)
    f1 = F.relu(F.conv2d(x, sd["net.slice1.0.weight"].to(x.dtype), sd["net.slice1.0.bias"].to(x.dtype), stride=2))
    f2 = fire(fire(pool(f1), 2, 3), 2, 4)
    f3 = fire(fire(pool(f2), 3, 6), 3, 7)
    f4 = fire(pool(f3), 4, 9)
    f5 = fire(f4, 5, 10)
    f6 = fire(f5, 6, 11)
    f7 = fire(f6, 7, 12)
    return [f1, f2, f3, f4, f5, f6, f7]


FEATURES = {"vgg": vgg_features, "alex": alex_features, "squeeze": squeeze_features}


def normalize_tensor(t, eps=1e-10):
    return t / (torch.sqrt(torch.sum(t ** 2, dim=1, keepdim=True)) + eps)


def lpips(sd, pred, target, net="vgg"):
    """returns [N,1,1,1] like PerceptualLoss.forward (lpips/__init__.py:26-41); net in {'vgg', 'alex', 'squeeze'}."""
    shift = torch.tensor(SHIFT, dtype=pred.dtype, device=pred.device)[None, :, None, None]
    scale = torch.tensor(SCALE, dtype=pred.dtype, device=pred.device)[None, :, None, None]
    f0 = FEATURES[net](sd, (pred - shift) / scale)
    f1 = FEATURES[net](sd, (target - shift) / scale)
    val = 0
    for kk in range(len(f0)):
        d = (normalize_tensor(f0[kk]) - normalize_tensor(f1[kk])) ** 2
        r = F.conv2d(d, sd[f"lin{kk}.model.1.weight"].to(pred.dtype)).mean([2, 3], keepdim=True)
        val = val + r
    return val

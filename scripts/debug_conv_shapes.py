"""conv2d_gradfix.conv2d (exact-fp32 direct-convolution kernels) vs torch.nn.functional.conv2d on assorted shapes (fwd, dgrad, wgrad)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morphganformer_b200.torch_utils.ops import conv2d_gradfix
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
cases = [(2, 3, 96, 96, 64, 11, 4, 2), (2, 64, 11, 11, 192, 5, 1, 2), (2, 192, 5, 5, 384, 3, 1, 1), (2, 3, 96, 96, 64, 3, 2, 0), (1, 3, 64, 64, 64, 11, 4, 2),
         (2, 3, 35, 35, 8, 11, 4, 2), (2, 3, 35, 35, 8, 7, 4, 2), (2, 3, 35, 35, 8, 7, 1, 3), (2, 5, 20, 20, 8, 5, 3, 2), (1, 1, 23, 23, 1, 11, 4, 2)]
for (n, ic, h, w, oc, k, s, p) in cases:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, ic, h, w, device="cuda", generator=g, requires_grad=True); wt = torch.randn(oc, ic, k, k, device="cuda", generator=g, requires_grad=True)
    b = torch.randn(oc, device="cuda", generator=g)
    y = conv2d_gradfix.conv2d(x, wt, b, stride=s, padding=p)
    yr = torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), stride=s, padding=p)
    dy = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, [x, wt], dy)
    gxr, gwr = torch.autograd.grad(yr, [x, wt], dy.double())
    rel = lambda a, r: ((a.double() - r).abs().max() / r.abs().max()).item()
    print("n%d ic%d %dx%d oc%d k%d s%d p%d -> out %s  fwd %.2e dgrad %.2e wgrad %.2e" % (n, ic, h, w, oc, k, s, p, tuple(y.shape[2:]), rel(y, yr), rel(gx, gxr), rel(gw, gwr)))

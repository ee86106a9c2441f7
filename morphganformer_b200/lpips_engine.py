"""LPIPS-VGG16 perceptual distance on the tcgen05 convolution kernel, forward and backward wrt the generated image.

Mirrors lpips.PerceptualLoss(model='net-lin', net='vgg') of the reference (lpips/__init__.py:13-41 -> dist_model.py:110-118 ->
networks_basic.py:64-92; VGG16 slices pretrained_networks.py:97-135): ScalingLayer, 13 conv3x3+ReLU, 4 max-pools, five taps,
channel-unit-normalisation, squared difference, 1x1 `lin` layers, spatial mean, sum over taps.  The target branch is constant
during a projection, so its normalised tap features are computed once (set_target) instead of every step as the reference does.
Also produces the MSE term of the projection loss in the same pass over the image (1024_example_percept_MSE.py:143).

Everything arithmetic is a kernel of libmgf_sm100a.so: mgf_vgg_conv1_fwd/bwd (ScalingLayer + conv1_1 + ReLU and its backward,
straight from / to the fp32 NCHW image), mgf_conv_tc (the other 12 convolutions and their input gradients, bias+ReLU and ReLU-mask
fused in the epilogue), mgf_maxpool2_*, mgf_lpips_head, mgf_lpips_tap_pool_bwd.
"""
import torch
from . import _lib, tc

# (cin, cout) per conv; 'P' = 2x2 max-pool.  Taps (relu1_2, relu2_2, relu3_3, relu4_3, relu5_3) follow conv indices 1,3,6,9,12.
VGG = [(3, 64), (64, 64), "P", (64, 128), (128, 128), "P", (128, 256), (256, 256), (256, 256), "P",
       (256, 512), (512, 512), (512, 512), "P", (512, 512), (512, 512), (512, 512)]
TAP_AFTER = {1: 0, 3: 1, 6: 2, 9: 3, 12: 4}
REF_NAMES = ["net.slice1.0", "net.slice1.2", "net.slice2.5", "net.slice2.7", "net.slice3.10", "net.slice3.12", "net.slice3.14",
             "net.slice4.17", "net.slice4.19", "net.slice4.21", "net.slice5.24", "net.slice5.26", "net.slice5.28"]
TAPS_B = [(0, 1 - ky, 1 - kx, ky * 3 + kx) for ky in range(3) for kx in range(3)]


def _L():
    return _lib.lib()


def _p(t):
    return t.data_ptr() if t is not None else None


class LpipsEngine:
    def __init__(self, state_dict, device="cuda"):
        """state_dict: reference PNetLin naming (net.sliceK.N.weight/bias, linK.model.1.weight)."""
        self.dev = torch.device(device)
        self.wf, self.wb, self.bias = [], [], []
        for i, name in enumerate(REF_NAMES):
            w = state_dict[name + ".weight"].detach().float().to(self.dev)          # [O, I, 3, 3]
            b = state_dict[name + ".bias"].detach().float().to(self.dev).contiguous()
            O, I = w.shape[:2]
            if i == 0:
                wc = torch.zeros(O, 32, device=self.dev)
                wc[:, :27] = w.permute(0, 2, 3, 1).reshape(O, 27)                    # column index = (ky*3+kx)*3 + c
                self.wf.append(wc.reshape(1, 1, O, 32).contiguous())                      # fp32 master, cast per forward dtype
                self.wb.append(wc.t().reshape(1, 1, 32, O).to(torch.bfloat16).contiguous())
            else:
                wk = w.reshape(O, I, 9)
                self.wf.append(wk.permute(2, 0, 1).reshape(1, 9, O, I).contiguous())
                self.wb.append(wk.permute(2, 1, 0).reshape(1, 9, I, O).to(torch.bfloat16).contiguous())
            self.bias.append(b)
        self.lin = [state_dict[f"lin{k}.model.1.weight"].detach().float().reshape(-1).to(self.dev).contiguous() for k in range(5)]
        self._st = {}
        self._wf16 = {}
        self.n1 = None
        self.target = None
        self.generation = 0          # bumped by every forward / set_target: identifies whose activations the buffers hold

    def _wfwd(self, ci):
        dt = _lib.forward_torch_dtype()
        w = self._wf16.get(ci)
        if w is None or w.dtype != dt:
            w = self._wf16[ci] = self.wf[ci].to(dt).contiguous()
        return w

    def _buf(self, name, shape, dtype=torch.bfloat16, fwd=False):
        if fwd:
            dtype = _lib.forward_torch_dtype()
        t = self._st.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._st[name] = t
        return t

    def _features(self, img, target, mse, tag, head_val=None):
        """Runs the VGG trunk; returns the list of conv outputs h[i] (post-ReLU, NHWC bf16) and pooled tensors.
        head_val: if given ([B] fp32, zeroed), the LPIPS head of every tap that is followed by a pool (relu1_2 .. relu4_3) is accumulated
        into it by the fused tap+pool kernel (one pass over the feature map instead of two); the caller adds the last tap."""
        B, _, R, _ = img.shape
        s = _lib.stream_ptr(self.dev)
        # ScalingLayer + conv1_1 + ReLU (+ the MSE sum) in one kernel, straight from the fp32 image (no im2col buffer)
        y0 = self._buf(f"{tag}h0", (B, R, R, 64), fwd=True)
        _lib.check(_L().mgf_vgg_conv1_fwd(_p(img), _p(target), _p(mse), _p(self.wf[0]), _p(self.bias[0]), _p(y0), B, R, s), "mgf_vgg_conv1_fwd")
        h, x, res, ci = [y0], y0, R, 1
        pooled = {}
        for item in VGG[1:]:
            if item == "P":
                y = self._buf(f"{tag}p{ci}", (B, res // 2, res // 2, x.shape[3]), fwd=True)
                if head_val is not None and (ci - 1) in TAP_AFTER:
                    k = TAP_AFTER[ci - 1]
                    stats = self._buf(f"{tag}stats{k}", (B, res, res, 2), torch.float32)      # per-pixel (|f|, g . f) for the backward kernel
                    _lib.check(_L().mgf_lpips_tap_pool_fwd(_p(x), _p(self.n1[k]), _p(self.lin[k]), _p(y), _p(head_val), _p(stats), B, res, res, x.shape[3], s),
                               "mgf_lpips_tap_pool_fwd")
                else:
                    _lib.check(_L().mgf_maxpool2_fwd(_p(x), _p(y), B, res, res, x.shape[3], s), "mgf_maxpool2_fwd")
                pooled[ci] = y
                x, res = y, res // 2
                continue
            cin, cout = item
            y = self._buf(f"{tag}h{ci}", (B, res, res, cout), fwd=True)
            tc.conv_tc([x], self._wfwd(ci), tc.TAPS_3X3, (B, res, res), 1, cout, y, bias=self.bias[ci], act=2, gain=1.0, tag="vgg.fwd")
            h.append(y)
            x = y
            ci += 1
        return h, pooled, None

    @torch.no_grad()
    def set_target(self, target):
        """target [B,3,R,R] fp32 in [-1,1]; caches the unit-normalised tap features (the reference recomputes them every step)."""
        self.generation += 1
        target = target.to(self.dev, torch.float32)
        if getattr(self, "target", None) is not None and tuple(self.target.shape) == tuple(target.shape):
            self.target.copy_(target)                   # keep addresses stable (CUDA-graph replay of the step)
        else:
            self.target = target.contiguous().clone()
            self.n1 = None
        if self.n1 is not None and self.n1[0].dtype != _lib.forward_torch_dtype():
            self.n1 = None
        target = self.target
        B = target.shape[0]
        # the trunk runs in the generated-image branch's buffers (tag "g": overwritten by every forward anyway), so a new job
        # allocates nothing once the first one has run -- addresses stay stable for graph replay and set_target costs ~6 ms
        h, _, _ = self._features(target, None, None, "g")
        fresh = self.n1 is None
        if fresh:
            self.n1 = []
        s = _lib.stream_ptr(self.dev)
        for ci, k in TAP_AFTER.items():
            f = h[ci]
            n = torch.empty_like(f) if fresh else self.n1[k]
            _lib.check(_L().mgf_lpips_head(0, _p(f), None, None, None, _p(n), None, 0, B, f.shape[1] * f.shape[2], f.shape[3], s), "mgf_lpips_head")
            if fresh:
                self.n1.append(n)

    @torch.no_grad()
    def forward(self, img, want_mse=True):
        """img [B,3,R,R] fp32 -> (lpips [B], mse_sum [B] = sum of squared differences to the target)."""
        assert self.n1 is not None, "call set_target first"
        self.generation += 1
        img = img.contiguous()
        B = img.shape[0]
        s = _lib.stream_ptr(self.dev)
        val = self._buf("val", (B,), torch.float32); val.zero_()
        mse = self._buf("mse", (B,), torch.float32); mse.zero_()
        h, pooled, col = self._features(img, self.target if want_mse else None, mse if want_mse else None, "g", head_val=val)
        self.h, self.pooled, self.img = h, pooled, img
        for ci, k in TAP_AFTER.items():
            if ci != 12:                      # taps followed by a pool were accumulated by the fused tap+pool kernel
                continue
            f = h[ci]
            _lib.check(_L().mgf_lpips_head(1, _p(f), _p(self.n1[k]), _p(self.lin[k]), None, None, _p(val), 0, B, f.shape[1] * f.shape[2], f.shape[3], s), "mgf_lpips_head")
        return val, mse

    @torch.no_grad()
    def backward(self, coef, mse_coef):
        """coef [B] fp32 = d(loss)/d(lpips_b); mse_coef = d(loss)/d(mse_sum_b) * 2 (scalar).  Returns d(loss)/d(img) fp32 NCHW."""
        h, B = self.h, self.img.shape[0]
        s = _lib.stream_ptr(self.dev)
        R = self.img.shape[2]

        def head_bwd(ci, relu_mask):
            f = h[ci]; k = TAP_AFTER[ci]
            out = self._buf(f"df{ci}", tuple(f.shape))
            _lib.check(_L().mgf_lpips_head(2, _p(f), _p(self.n1[k]), _p(self.lin[k]), _p(coef), _p(out), None, int(relu_mask), B,
                                           f.shape[1] * f.shape[2], f.shape[3], s), "mgf_lpips_head")
            return out

        # walk the trunk backwards; `g` is always the gradient wrt the PRE-activation of conv `ci` (ReLU mask already applied)
        g = head_bwd(12, True)
        ci = 12
        items = [it for it in VGG]
        # conv index -> is it directly preceded by a pool?
        pos = len(items) - 1
        while ci >= 1:
            res = h[ci].shape[1]
            cin = VGG_CIN[ci]
            prev_is_pool = items[pos - 1] == "P"
            if prev_is_pool:
                # d(pooled input) -> route through the pool to conv ci-1's output, add its tap gradient, apply its ReLU mask
                dp = self._buf(f"dp{ci}", (B, res, res, cin))
                tc.conv_tc([g], self.wb[ci], TAPS_B, (B, res, res), 1, cin, dp, tag="vgg.bwd", fwd=False)
                src = h[ci - 1]
                g2 = self._buf(f"dpre{ci - 1}", tuple(src.shape))
                if (ci - 1) in TAP_AFTER:      # tap + pool: one fused kernel (head gradient + pool routing + ReLU mask)
                    k = TAP_AFTER[ci - 1]
                    _lib.check(_L().mgf_lpips_tap_pool_bwd(_p(src), _p(self.n1[k]), _p(self.lin[k]), _p(coef), _p(dp), _p(g2), _p(self._st.get(f"gstats{k}")), B,
                                                           src.shape[1], src.shape[2], src.shape[3], s), "mgf_lpips_tap_pool_bwd")
                else:
                    _lib.check(_L().mgf_maxpool2_bwd(_p(src), _p(dp), None, _p(g2), B, src.shape[1], src.shape[2], src.shape[3], s), "mgf_maxpool2_bwd")
                g = g2
                pos -= 2
            else:
                src = h[ci - 1]
                g2 = self._buf(f"dpre{ci - 1}", tuple(src.shape))
                tc.conv_tc([g], self.wb[ci], TAPS_B, (B, res, res), 1, cin, g2, X=src, actgrad=True, ag_alpha=0.0, ag_gain=1.0, tag="vgg.bwd", fwd=False)
                g = g2
                pos -= 1
            ci -= 1
        dimg = torch.empty_like(self.img)
        # conv1_1 input gradient (64 -> 27 product + col2im in shared memory) / scale + the MSE gradient, one kernel
        _lib.check(_L().mgf_vgg_conv1_bwd(_p(g), _p(self.wf[0]), _p(self.img), _p(self.target) if mse_coef else None, float(mse_coef), _p(dimg), B, R, s),
                   "mgf_vgg_conv1_bwd")
        return dimg


VGG_CIN = [c[0] for c in VGG if c != "P"]


class PerceptualLoss(torch.nn.Module):
    """API mirror of lpips.PerceptualLoss (reference lpips/__init__.py:13-41) for model='net-lin', net='vgg':
    forward(pred, target) -> [N,1,1,1], differentiable wrt pred.  The target features are cached per target tensor."""

    def __init__(self, state_dict, model="net-lin", net="vgg", use_gpu=True, **_unused):
        super().__init__()
        if model != "net-lin" or net not in ("vgg", "vgg16"):
            raise NotImplementedError("only the LPIPS-VGG16 ('net-lin', 'vgg') variant is built (SURVEY.md 8f rank 2)")
        self.engine = LpipsEngine(state_dict)
        self._target_ref = None           # (the caller's target tensor itself, its _version): a strong reference, so the cached
                                          # features can never be matched by a NEW tensor that re-uses a freed allocation's address

    def forward(self, pred, target, normalize=False):
        cached = self._target_ref
        if cached is None or cached[0] is not target or cached[1] != target._version:
            self.engine.set_target(2 * target - 1 if normalize else target)
            self._target_ref = (target, target._version)
        if normalize:
            pred = 2 * pred - 1
        return _LpipsFn.apply(pred, self.engine).reshape(-1, 1, 1, 1)


class _LpipsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, eng):
        val, _ = eng.forward(pred.detach().float(), want_mse=False)
        ctx.eng = eng
        ctx.generation = eng.generation          # the engine keeps ONE set of activations: a later forward invalidates this graph
        return val.clone()

    @staticmethod
    def backward(ctx, dval):
        eng = ctx.eng
        if ctx.generation != eng.generation:
            raise RuntimeError("PerceptualLoss: backward of a forward pass whose activations were overwritten by a later forward / set_target "
                               "call on the same module (the engine is single-slot: call backward before the next forward, or use one module per graph)")
        dimg = eng.backward(dval.contiguous().float(), 0.0)
        return dimg, None

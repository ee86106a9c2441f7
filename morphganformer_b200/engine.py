"""bf16 tcgen05 synthesis engine: G.synthesis(ws) forward and its explicit backward wrt ws.

Host orchestration only -- every arithmetic step is a kernel of libmgf_sm100a.so called through the C ABI
(include/mgf.h): tcgen05 implicit-GEMM convolutions (mgf_conv_tc), fused duplex attention (mgf_attn_*), and the
HBM-bound helpers (style/demod, weight modulation, ToRGB, FIR up-sampling, activation backward).
PyTorch supplies device memory, the stream and the autograd hook (one torch.autograd.Function around the whole
synthesis network so the (tiny) mapping network can stay in eager PyTorch).

Dataflow (reference training/networks.py:1244-1264, :1132-1174, :1010-1042, :252-328), activations NHWC bf16:
  per layer   s = affine(w_global) ; d = rsqrt(sum (w s)^2 + 1e-8)                       mgf_style_fwd
              Wf[b] = bf16(W * s[b] * d[b])   (demodulation folded into the weight tile)  mgf_modulate_weights
              y = conv3x3(x, Wf[b])           (up-conv: 4 output phases, FIR folded in)    mgf_conv_tc  (+ noise/bias/lrelu epilogue)
              attention layers (res <= 128):  z = lrelu(xn(y) * (1 + A VM + bm) + noise + b)  mgf_attn_fwd
  resnet skip v = conv1x1(x_in) ; x_out = z1 + sqrt(.5) * FIRup2(v)                        mgf_conv_tc + mgf_upfir2_add
  last block  yl = conv_last ; img = ToRGB(yl) (fp32 NCHW)                                  mgf_conv_tc + mgf_torgb_fwd
Backward mirrors it with dgrad convs (the same kernel on transposed weights; up-conv dgrad = 36 taps over 4 strided phase
views), d(styles) reduced in the dgrad epilogue (sum_p dxs * x), d(demod) from R[b,o] = sum_p dy * y.
"""
import ctypes
import math
import numpy as np
import torch

from . import _lib, tc
from .torch_utils.ops import upfirdn2d as _upfirdn2d

SQRT2 = math.sqrt(2.0)
SQRT_HALF = math.sqrt(0.5)
import os as _os
BWD_SIDE_REDUCTIONS = _os.environ.get("MGF_BWD_SIDE", "1") != "0"      # A/B switch: d(style) reductions on the side stream
LRELU_ALPHA = 0.2
POINTWISE_SKIP = _os.environ.get("MGF_POINTWISE_SKIP", "1") != "0"     # A/B switch: resnet-skip 1x1 convolutions on the streaming pointwise kernel
TWO_STAGE_MAX_RES = int(_os.environ.get("MGF_TWO_STAGE_MAX_RES", "65536"))   # A/B switch: up-conv dgrad as FIR + 9-tap strided conv up to this output size
# A/B switches: up-conv forward as transposed conv (9 taps over four parity GEMMs) + FIR pass for output sizes in [MIN, MAX]; above MAX the
# FIR-folded four-phase form runs in ONE tile per pixel block (BN = 4 * Cout columns, activation tiles fetched once for the four phases)
# with the layer tail fused.  Measured: the folded form costs 4x the tensor work, which only matters while the layer is compute-bound
# (<= 128^2 outputs, C >= 256: two-stage saves ~0.1 ms on each of the 64^2 / 128^2 layers); at 256^2 .. 1024^2 the extra HBM round trip +
# FIR pass of the two-stage form costs more than it saves.  DEFAULT OFF (MAX = 0): the 16-bit (2h+1)^2 intermediate is one more rounding
# per up-convolution -- image error 8.4e-3 -> 1.2e-2 at 64^2 -- and the 1e-2 image tolerance has no room for it.
TWO_STAGE_FWD_MIN_RES = int(_os.environ.get("MGF_TWO_STAGE_FWD_MIN_RES", "8"))
TWO_STAGE_FWD_MAX_RES = int(_os.environ.get("MGF_TWO_STAGE_FWD_MAX_RES", "0"))
FIR4_TAPS = (ctypes.c_float * 4)(0.125, 0.375, 0.375, 0.125)      # [1,3,3,1] / 8 (symmetric: flipping is a no-op); gain 4 = up^2 passed separately


def _L():
    return _lib.lib()


def _s(dev):
    return _lib.stream_ptr(dev)


def _p(t):
    return t.data_ptr() if t is not None else None


# ------------------------------------------------------------------------------------------------ host-side constant folding
def up_phase_matrix(f1d=(1, 3, 3, 1)):
    """Cm[ph, t, k]: the up-sampling conv (conv_transpose2d stride 2 with the un-flipped 3x3 weights, then upfirdn2d(pad=1, gain=4),
    reference conv2d_resample.py:117-134 as called from SynthesisLayer with up=2, padding=1, flip_weight=False) written as
       y[2m+py, 2n+px] = sum_{t=(ty,tx)} sum_{k=(ky,kx)} Cm[ph, t, k] * w[k] * x[m+ty-1, n+tx-1],   ph = 2*py + px.
    Computed numerically from the impulse response (pure numpy, no learned data involved)."""
    f = np.asarray(f1d, dtype=np.float64)
    f2 = np.outer(f, f); f2 /= f2.sum()
    fflip = f2[::-1, ::-1] * 4.0          # upfirdn2d: true convolution (flipped taps), gain = up^2
    H = 7; c = 3                           # impulse at (c, c) of an HxH input
    Cm = np.zeros((4, 9, 9))
    for ky in range(3):
        for kx in range(3):
            ct = np.zeros((2 * H + 1, 2 * H + 1))
            ct[2 * c + ky, 2 * c + kx] = 1.0                       # conv_transpose2d, stride 2, padding 0: out[2m+ky] += x[m] w[ky]
            pad = np.pad(ct, 1)
            out = np.zeros((2 * H, 2 * H))
            for fy in range(4):
                for fx in range(4):
                    out += pad[fy:fy + 2 * H, fx:fx + 2 * H] * fflip[fy, fx]
            for py in range(2):
                for px in range(2):
                    for ty in range(3):
                        for tx in range(3):
                            m, n = c - (ty - 1), c - (tx - 1)      # output low-res position whose window tap (ty,tx) hits the impulse
                            Cm[2 * py + px, ty * 3 + tx, ky * 3 + kx] = out[2 * m + py, 2 * n + px]
            # the response must be fully covered by the 3x3 windows of the four phases
            cover = np.zeros_like(out)
            for ty in range(3):
                for tx in range(3):
                    m, n = c - (ty - 1), c - (tx - 1)
                    cover[2 * m:2 * m + 2, 2 * n:2 * n + 2] = 1
            assert np.abs(out * (1 - cover)).max() < 1e-12, "up-conv response leaves the 3x3 low-res window"
    return Cm


def _fc_gain(fc):
    return float(fc.w_gain), float(fc.b_gain)


class _Layer:
    """Frozen, pre-folded parameters of one SynthesisLayer (+ its attention) on the device."""
    pass


class SynthesisEngine:
    def __init__(self, synthesis):
        self.net = synthesis
        self.dev = next(synthesis.parameters()).device
        if self.dev.type != "cuda":
            raise _lib.MgfError("SynthesisEngine needs the module on a CUDA device (no CPU fallback)")
        if synthesis.architecture not in ("resnet", "skip", "orig"):
            raise NotImplementedError("tc engine: unknown synthesis architecture %r" % (synthesis.architecture,))
        if synthesis.k != 17:
            # VM buffers, mgf_attn_fwd/bwd and mgf_small_gemm are built for 16 local components + 1 global latent (GANformer default,
            # reference networks.py:1185-1218 with k = components_num + 1); other counts would read Kf / Sc / maskbias out of bounds
            raise NotImplementedError("tc engine: built for k = 17 latents (16 components + 1 global), got k = %d (use engine='ops')" % synthesis.k)
        self.res = synthesis.img_res
        self.k = synthesis.k
        self.num_ws = synthesis.num_ws
        self.Cm = torch.from_numpy(up_phase_matrix()).float()
        self.blocks = []
        self._states = {}
        self._side = None          # side stream for the per-layer operand preparation (forward_raw)
        self.refresh()

    # -------------------------------------------------------------------------------------------- weight folding
    @torch.no_grad()
    def refresh(self):
        """(Re)fold all weights; call again if the module's parameters change."""
        dev = self.dev
        self.blocks = []
        w_idx = 0
        self.arch = arch = self.net.architecture            # "resnet" (GANformer default) | "skip" | "orig" (reference networks.py:1070-1174)
        for r in self.net.block_resolutions:
            blk = getattr(self.net, f"b{r}")
            e = {"res": r, "stem": blk.stem, "last": blk.is_last}
            if blk.stem:
                e["const"] = blk.const.detach().float().permute(1, 2, 0).contiguous()                       # [4,4,C] fp32 master
                e["const_absmax"] = float(e["const"].abs().max())
                e["conv1"] = self._fold_layer(blk.conv1, w_idx, gain=1.0); w_idx += 1
            else:
                e["conv0"] = self._fold_layer(blk.conv0, w_idx, gain=1.0); w_idx += 1
                e["conv1"] = self._fold_layer(blk.conv1, w_idx, gain=SQRT_HALF if arch == "resnet" else 1.0); w_idx += 1
                if arch == "resnet":
                    wsk = (blk.skip.weight.detach().float() * float(blk.skip.w_gain))[:, :, 0, 0]                 # [O, I]
                    e["skip_f32"] = wsk.reshape(1, 1, *wsk.shape).contiguous()                                      # [1,1,O,I] fp32 master
                    e["skip_b"] = wsk.t().reshape(1, 1, wsk.shape[1], wsk.shape[0]).to(torch.bfloat16).contiguous()  # [1,1,I,O]
                    f1 = np.array([1, 3, 3, 1], dtype=np.float64); f1 = f1 / f1.sum()
                    e["fk4"] = (ctypes.c_float * 4)(*[float(v) for v in f1[::-1]])
                    e["skip_gain"] = 4.0 * SQRT_HALF
            if blk.is_last:
                e["conv_last"] = self._fold_layer(blk.conv_last, w_idx, gain=1.0); w_idx += 1
            if blk.is_last or arch == "skip":      # ToRGB reads the ws slot after this block's convs (shared with the next block's conv0, :1134-1174)
                tr = blk.torgb
                C = tr.weight.shape[1]
                e["rgb"] = dict(idx=w_idx, C=C, w=tr.weight.detach().float().reshape(3, C).contiguous(),
                                A=tr.affine.weight.detach().float().contiguous(), ab=(tr.affine.bias.detach().float() * float(tr.affine.b_gain)).contiguous(),
                                again=float(tr.affine.w_gain), sgain=float(tr.w_gain), bias=tr.biasAct.bias.detach().float().contiguous())
                e["fir"] = blk.resample_kernel.detach().float()            # image up-sampling filter of the 'skip' architecture
            self.blocks.append(e)
        self.pos = None

    def _fold_layer(self, m, idx, gain):
        dev = self.dev
        L = _Layer()
        L.idx, L.up, L.res = idx, m.up, m.out_res
        W = (m.weight.detach().float() * float(m.w_gain))           # [O, I, 3, 3]
        O, I = W.shape[:2]
        L.O, L.I = O, I
        L.Wsq = W.square().sum(dim=[2, 3]).contiguous()              # [O, I]
        Wk = W.reshape(O, I, 9)
        if m.up == 1:
            L.Bf = Wk.permute(2, 0, 1).contiguous()                  # [9, O, I]
            L.Bb = Wk.permute(2, 1, 0).contiguous()                  # [9, I, O]
            L.taps_f = [(0, ky - 1, kx - 1, ky * 3 + kx) for ky in range(3) for kx in range(3)]
            L.taps_b = [(0, 1 - ky, 1 - kx, ky * 3 + kx) for ky in range(3) for kx in range(3)]
            L.phases = 1
        else:
            Weff = torch.einsum("ptk,oik->ptoi", self.Cm.to(dev), Wk)                  # [4, 9, O, I]
            L.phases = 4
            # forward in the reference's own two-stage form (conv2d_resample.py:117-134): ct = conv_transpose2d(x, w, stride 2) on the
            # (2h+1) x (2w+1) grid, then the 4x4 FIR (mgf_fir4, with the layer tail fused).  The four parities of ct are four small GEMMs
            # over the low-resolution grid with 4 / 2 / 2 / 1 taps (9 in all, on the plain [9, O, I] weights) instead of the 4 x 9 taps of
            # the FIR-folded phase kernels: a quarter of the tensor work and of the L2 -> shared-memory operand traffic, for one extra
            # HBM round trip of the (2h+1)^2 tensor.  ct[2m+ey, 2n+ex] = sum_{a,b} x[m-a, n-b] w[2a+ey, 2b+ex] over the taps that exist.
            L.two_stage_fwd = TWO_STAGE_FWD_MIN_RES <= L.res <= TWO_STAGE_FWD_MAX_RES
            if L.two_stage_fwd:
                L.Bf = Wk.permute(2, 0, 1).contiguous()                                   # [9, O, I]
                L.taps_f = ([(0, -a, -b, (2 * a) * 3 + 2 * b) for a in (0, 1) for b in (0, 1)] + [(0, -a, 0, (2 * a) * 3 + 1) for a in (0, 1)]
                            + [(0, 0, -b, 3 + 2 * b) for b in (0, 1)] + [(0, 0, 0, 4)])
                L.phase_ntaps = (4, 2, 2, 1)
            else:
                L.Bf = Weff.permute(1, 0, 2, 3).reshape(9, 4 * O, I).contiguous()       # [9, 4*O, I]
                L.taps_f = [(0, ty - 1, tx - 1, ty * 3 + tx) for ty in range(3) for tx in range(3)]
            # input gradient in the reference's own two-stage form (adjoint of conv_transpose2d(stride 2) -> FIR): g = FIR4(dy) on the
            # (2h+1) x (2w+1) grid (mgf_fir4_pad), then dx[i,j] = sum_k W_k^T g[2i+ky, 2j+kx]: a stride-2 3x3 conv = 9 taps over the four
            # phase views of g (view (ky%2, kx%2), shift (ky//2, kx//2)) instead of 36 taps of the folded four-phase kernels
            # Measured (8 images): the 9-tap form halves the dgrad time at every resolution, but the FIR pass costs 0.08-0.54 ms at
            # 128^2 .. 1024^2 (instruction-bound at 1.9 TB/s), so it only wins up to 128^2 outputs; above that the folded 36-tap form stays.
            L.two_stage = L.res <= TWO_STAGE_MAX_RES
            if L.two_stage:
                L.Bb = Wk.permute(2, 1, 0).contiguous()                                   # [9, I, O]
                L.taps_b = [((ky % 2) * 2 + (kx % 2), ky // 2, kx // 2, ky * 3 + kx) for ky in range(3) for kx in range(3)]
            else:
                L.Bb = Weff.permute(0, 1, 3, 2).reshape(36, I, O).contiguous()            # [36, I, O]
                L.taps_b = [(ph, 1 - ty, 1 - tx, ph * 9 + ty * 3 + tx) for ph in range(4) for ty in range(3) for tx in range(3)]
        # low-resolution layers: per-sample weight tensors (B x 4.7-19 MB) would dwarf the activations, so modulate the ACTIVATIONS
        # instead (xs = x*s, epilogue *d; reference non-fused algebra, networks.py:314-324) and keep ONE shared weight tensor
        L.shared_w = L.res <= 128 and O * I >= 256 * 256
        L.Bb16 = L.Bb.unsqueeze(0).to(torch.bfloat16).contiguous() if L.shared_w else None     # [1, T, I, O]
        L._Bf16 = None
        if m.up == 1:
            L.two_stage = False
            L.two_stage_fwd = False
        # 32 -> 32 channel 3x3 layers (the 1024^2 block): 64-byte pixel rows halve the TMA efficiency, so view two neighbouring pixels as
        # one 64-channel super-pixel (same memory) and expand the weights to the block form [9][(po,o)][(pi,i)]
        L.superpix = (m.up == 1 and O == 32 and I == 32 and L.res >= 64 and not L.shared_w and m.transformer is None)
        if L.superpix:
            Bf2 = torch.zeros(9, 64, 64, device=W.device); Bb2 = torch.zeros(9, 64, 64, device=W.device)
            for ky in range(3):
                for ks in range(3):            # super-pixel shift dxs = ks - 1
                    for po in range(2):
                        for pi in range(2):
                            kx = 2 * (ks - 1) + pi - po + 1          # forward: x_in = x_out + kx - 1
                            if 0 <= kx < 3:
                                Bf2[ky * 3 + ks, po * 32:(po + 1) * 32, pi * 32:(pi + 1) * 32] = W[:, :, ky, kx]
                            kb = po + 1 - 2 * (ks - 1) - pi          # dgrad: dy pixel x' = x + 1 - kx
                            if 0 <= kb < 3:
                                Bb2[ky * 3 + ks, po * 32:(po + 1) * 32, pi * 32:(pi + 1) * 32] = W[:, :, ky, kb].t()
            L.Bf, L.Bb = Bf2.contiguous(), Bb2.contiguous()
            L.taps_f = [(0, ky - 1, ks - 1, ky * 3 + ks) for ky in range(3) for ks in range(3)]
            L.taps_b = [(0, 1 - ky, ks - 1, ky * 3 + ks) for ky in range(3) for ks in range(3)]
        L.A = m.affine.weight.detach().float().contiguous()          # [I, 32]
        L.ab = (m.affine.bias.detach().float() * float(m.affine.b_gain)).contiguous()
        L.again = float(m.affine.w_gain)
        L.has_noise = bool(m.local_noise)
        L.noise = m.noise_const.detach().float().contiguous() if m.local_noise else None
        L.nstr = m.noise_strength.detach().float().reshape(1).contiguous() if m.local_noise else None
        L.has_bias = m.biasAct is not None
        L.bias = (m.biasAct.bias.detach().float() * float(m.biasAct.b_gain)).contiguous() if L.has_bias else None
        L.gain = SQRT2 * gain if L.has_bias else 1.0
        L.attn = m.transformer is not None
        if L.attn:
            t = m.transformer
            if not (t.kmeans and t.parametric and t.num_heads == 1 and t.integration == "mul" and t.norm == "layer"):
                raise NotImplementedError("tc engine: attention variant outside the GANformer-default configuration")
            if t.to_len != 16 or t.centroids.shape[-2] != 16:
                raise NotImplementedError("tc engine: attention over %d components (the kernels are built for 16)" % t.to_len)
            C = O
            rs = 1.0 / math.sqrt(float(t.size_head))
            Wq = t.to_queries.weight.detach().float() * float(t.to_queries.w_gain)
            bq = t.to_queries.bias.detach().float() * float(t.to_queries.b_gain)
            aw = t.att_weight.detach().float().reshape(-1)
            cen = t.centroids.detach().float()[0, 0]                                      # [16, 2C]
            awc1, awc2 = cen[:, :C] * aw[:C], cen[:, C:] * aw[C:]
            L.Kf = ((awc1 @ Wq) * rs).contiguous()                                        # [16, C]
            gp = m.grid_pos.detach().float().reshape(-1, m.grid_pos.shape[-1])           # [HW, 32]
            P = gp @ (t.from_pos_map.weight.detach().float() * float(t.from_pos_map.w_gain)).t() + t.from_pos_map.bias.detach().float() * float(t.from_pos_map.b_gain)
            L.Sc = (((awc1 @ bq).unsqueeze(0) + P @ awc2.t()) * rs).contiguous()          # [HW, 16]
            Wv = t.to_values.weight.detach().float() * float(t.to_values.w_gain)          # [C, 32]
            bv = t.to_values.bias.detach().float() * float(t.to_values.b_gain)
            Wm = t.modulation.weight.detach().float() * float(t.modulation.w_gain)        # [C, C]
            L.WVM = (Wm @ Wv).contiguous()                                                # [C, 32]
            L.bVM = (Wm @ bv).contiguous()
            L.WVMt = L.WVM.t().contiguous()                                               # [32, C]
            L.bm = (t.modulation.bias.detach().float() * float(t.modulation.b_gain)).contiguous()
            L.att_dp = float(t.att_dp.p)
            L.tabK, L.tabK_dtype = None, None
        return L

    # -------------------------------------------------------------------------------------------- buffers
    def _buf(self, st, name, shape, dtype=torch.bfloat16, zero=False, fwd=False):
        """fwd=True: a forward-dtype tensor (bf16 or fp16 per _lib.set_forward_dtype); default bf16 = gradient tensors."""
        if fwd:
            dtype = _lib.forward_torch_dtype()
        t = st.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.dev)
            st[name] = t
        return t

    def _zbuf(self, st, name, shape):
        """fp32 reduction target (d(style), R, dVM, ...) carved out of one per-state arena that backward_raw clears with a single fill
        at its start, instead of ~50 separate fill launches per step.  Offsets are assigned on first use and never change."""
        key = "z:" + name
        t = st.get(key)
        if t is None:
            n = 1
            for s in shape:
                n *= int(s)
            off = st.get("zpool_used", 0)
            pool = st.get("zpool")
            if pool is None:
                # 4 MB per 8 samples: ~4x what a 1024^2 generator needs (every target is [B, channels] or [B, 16, channels])
                pool = st["zpool"] = torch.zeros((1 << 20) * max(1, (st.get("batch", 8) + 7) // 8), dtype=torch.float32, device=self.dev)
            if off + n > pool.numel():
                raise _lib.MgfError("engine: reduction arena exhausted (%d + %d floats)" % (off, n))
            t = st[key] = pool[off:off + n].view(*shape)
            st["zpool_used"] = off + ((n + 63) // 64) * 64
        return t

    def _state(self, B):
        st = self._states.get(B)
        if st is None:
            st = {}
            self._states[B] = st
        return st

    # -------------------------------------------------------------------------------------------- kernels
    def _styles(self, L, ws, st, B):
        s = self._buf(st, f"s{L.idx}", (B, L.I), torch.float32)
        d = self._buf(st, f"d{L.idx}", (B, L.O), torch.float32)
        wg = ws[:, -1, L.idx]                                   # [B, 32] strided view
        _lib.check(_L().mgf_style_fwd(_p(wg), wg.stride(0), _p(L.A), _p(L.ab), L.again, 1.0, _p(L.Wsq), _p(s), _p(d),
                                      B, L.I, L.O, wg.shape[1], _s(self.dev)), "mgf_style_fwd")
        return s, d

    def _modulate(self, base, rs, nmod, cs, out, B, out_fwd):
        T, NT, K = base.shape
        _lib.check(_L().mgf_modulate_weights(_p(base), _p(rs), nmod, _p(cs), _p(out), int(out_fwd), B, T, NT, K, _s(self.dev)), "mgf_modulate_weights")

    def _skip_f(self, e):
        dt = _lib.forward_torch_dtype()
        w = e.get("skip_f")
        if w is None or w.dtype != dt:
            w = e["skip_f"] = e["skip_f32"].to(dt).contiguous()
        return w

    def _prep_layer(self, L, ws, st, B):
        """Everything of a layer that depends only on ws: styles s / demod d, the per-sample modulated forward weights and (attention
        layers) VM = (Y Wv^T + bv) Wm^T.  forward_raw runs this for ALL layers on a side stream, so these ~60 small latency-bound
        launches overlap the convolutions instead of sitting between them; the main stream waits on the layer's event."""
        s, d = self._styles(L, ws, st, B)
        Wf, VM = None, None
        if L.shared_w:
            dt = _lib.forward_torch_dtype()
            if L._Bf16 is None or L._Bf16.dtype != dt:
                L._Bf16 = L.Bf.unsqueeze(0).to(dt).contiguous()                      # [1, T, NT, I]
            Wf = L._Bf16
        elif L.superpix:
            s2, d2 = s.repeat(1, 2).contiguous(), d.repeat(1, 2).contiguous()
            st[f"s2_{L.idx}"], st[f"d2_{L.idx}"] = s2, d2
            Wf = self._buf(st, f"Wf{L.idx}", (B,) + tuple(L.Bf.shape), fwd=True)
            self._modulate(L.Bf, d2, 64, s2, Wf, B, True)
        else:
            Wf = self._buf(st, f"Wf{L.idx}", (B,) + tuple(L.Bf.shape), fwd=True)
            self._modulate(L.Bf, d, L.O, s, Wf, B, True)
        if L.attn:
            VM = self._buf(st, f"VM{L.idx}", (B, 16, L.O), torch.float32)
            comps = ws[:, :-1, L.idx]                           # [B,16,32] strided
            _lib.check(_L().mgf_small_gemm(_p(comps), comps.stride(0), comps.stride(1), _p(L.WVM), _p(L.bVM), _p(VM),
                                           16 * L.O, L.O, B, 16, L.O, comps.shape[2], 0, _s(self.dev)), "mgf_small_gemm")
            # 16-bit coefficient tables of VM (per step, here on the side stream) and of Kf (once per fold and forward dtype) in the attention
            # kernels' shared-memory layout: every CTA copies them instead of rebuilding them from the fp32 constants
            fdt = _lib.forward_torch_dtype()
            tabV = self._buf(st, f"tabV{L.idx}", (B * int(_L().mgf_attn_table_bytes(1, L.O)),), torch.uint8)
            need_k = L.tabK is None or L.tabK_dtype != fdt
            if need_k:
                L.tabK = torch.empty(int(_L().mgf_attn_table_bytes(0, L.O)), dtype=torch.uint8, device=self.dev)
                L.tabK_dtype = fdt
            _lib.check(_L().mgf_attn_tables(_p(L.Kf) if need_k else None, _p(VM), _p(L.tabK) if need_k else None, _p(tabV), B, L.O, _s(self.dev)),
                       "mgf_attn_tables")
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        st[f"prep{L.idx}"] = (s, d, Wf, VM, ev)

    def _layer_fwd(self, L, x_in, ws, maskbias, st, B, noise_on, add=None):
        """x_in [B,h,w,I] bf16 -> z [B,H,W,O] bf16 (post noise/bias/act)."""
        s, d, Wf, VM, ev = st[f"prep{L.idx}"]
        torch.cuda.current_stream(self.dev).wait_event(ev)
        h, w = x_in.shape[1], x_in.shape[2]
        H, Wd = h * L.up, w * L.up
        st[f"xin{L.idx}"] = x_in
        scale_n = None
        if L.shared_w:
            xs = self._buf(st, f"xs{L.idx}", tuple(x_in.shape), fwd=True)
            _lib.check(_L().mgf_scale_channels(_p(x_in), _p(s), _p(xs), 1, B, h * w, L.I, _s(self.dev)), "mgf_scale_channels")
            a_in = xs
            scale_n = d if L.phases == 1 else d.repeat(1, L.phases).contiguous()      # [B, NT]
        elif L.superpix:
            a_in = x_in.view(B, h, w // 2, 64)
        else:
            a_in = x_in
        noise, nbs = self._noise_of(L, st, B, H, Wd, draw=True)
        nstr = L.nstr if noise is not None else None
        kw = dict(osy=L.up, osx=L.up, ofy=(0, 0, 1, 1), ofx=(0, 1, 0, 1))
        if L.two_stage_fwd:
            # transposed convolution onto the (2h+1) x (2w+1) grid (buffer padded to even sizes; the extra row / column come out as exact
            # zeros: their taps read only TMA zero fill), then the FIR pass -- fused with noise + bias + leaky-ReLU unless attention follows
            ct = self._buf(st, f"ct{L.idx}", (B, H + 2, Wd + 2, L.O), fwd=True)
            tc.conv_tc([a_in], Wf, L.taps_f, (B, h + 1, w + 1), 4, L.O, ct, scale_n=scale_n, tag="g.fwd", phase_ntaps=L.phase_ntaps,
                       alg_scale=float(h * w) / float((h + 1) * (w + 1)), **kw)
            out = self._buf(st, f"y{L.idx}" if L.attn else f"z{L.idx}", (B, H, Wd, L.O), fwd=True)
            tail = L.has_bias and not L.attn
            _lib.check(_L().mgf_fir4(_p(ct), _p(out), FIR4_TAPS, 4.0, -1, B, H + 2, Wd + 2, H, Wd, H, Wd, L.O, 1, 1,
                                     _p(noise) if not L.attn else None, _p(nstr) if not L.attn else None, nbs, _p(L.bias) if tail else None,
                                     1 if tail else 0, LRELU_ALPHA, L.gain if tail else 1.0, _s(self.dev)), "mgf_fir4")
            if not L.attn:
                return out
            y = out
        if L.attn:
            y = self._buf(st, f"y{L.idx}", (B, H, Wd, L.O), fwd=True)
            if not L.two_stage_fwd:
                tc.conv_tc([a_in], Wf, L.taps_f, (B, h, w), L.phases, L.O, y, scale_n=scale_n, alg_scale=1.0 / L.phases, tag="g.fwd", **kw)
            z = self._buf(st, f"z{L.idx}", (B, H, Wd, L.O), fwd=True)
            dmask = None
            if st.get("train"):
                # attention dropout of training mode (reference networks.py:505-513, :374-376): cell mask [B,1,HW,16] then column mask
                # [B,1,1,16], both torch dropouts of rate attention_dropout / 2 drawn right after the layer's noise plane -- the same draws,
                # in the same order, as the ops engine makes, so one torch seed gives both engines identical masks
                pdrop = float(L.att_dp)
                if pdrop > 0:
                    m1 = torch.nn.functional.dropout(torch.ones([B, 1, H * Wd, 16], device=self.dev), pdrop, True)
                    m2 = torch.nn.functional.dropout(torch.ones([B, 1, 1, 16], device=self.dev), pdrop, True)
                    dmask = (m1 * m2).reshape(B, H * Wd, 16).contiguous()
            st[f"dmask{L.idx}"] = dmask
            probs = None
            if st.get("want_probs"):                            # attention maps requested: [B, HW, 16] fp32 per attention layer, in layer order
                probs = torch.empty(B, H * Wd, 16, device=self.dev)
                st["probs"].append(probs)
            _lib.check(_L().mgf_attn_fwd(_p(y), _p(L.Kf), _p(L.Sc), _p(maskbias), _p(VM), _p(L.bm), _p(noise), _p(nstr), _p(L.bias),
                                         L.gain, LRELU_ALPHA, _p(z), _p(probs), _p(dmask), _p(L.tabK), _p(st[f"tabV{L.idx}"]),
                                         B, H * Wd, L.O, nbs, _s(self.dev)), "mgf_attn_fwd")
        elif L.superpix:
            z = self._buf(st, f"z{L.idx}", (B, H, Wd, L.O), fwd=True)
            bias2 = L.bias.repeat(2).contiguous() if L.bias is not None else None
            tc.conv_tc([a_in], Wf, L.taps_f, (B, h, w // 2), 1, 64, z.view(B, H, Wd // 2, 64), noise=noise, noise_strength=nstr, bias=bias2,
                       act=1 if L.has_bias else 0, alpha=LRELU_ALPHA, gain=L.gain, alg_scale=0.5, tag="g.fwd", superpix=True, noise_bstride=nbs)
        else:
            z = self._buf(st, f"z{L.idx}", (B, H, Wd, L.O), fwd=True)
            tc.conv_tc([a_in], Wf, L.taps_f, (B, h, w), L.phases, L.O, z, scale_n=scale_n, noise=noise, noise_strength=nstr, bias=L.bias,
                       act=1 if L.has_bias else 0, alpha=LRELU_ALPHA, gain=L.gain, add=add, alg_scale=1.0 / L.phases, tag="g.fwd", noise_bstride=nbs, **kw)
        return z

    # -------------------------------------------------------------------------------------------- forward
    @torch.no_grad()
    def _noise_of(self, L, st, B, H, Wd, draw=False):
        """(noise tensor or None, elements between per-sample planes).  'const': the layer's [H,W] buffer shared by all samples;
        'random': a fresh randn([B,1,H,W]) per layer and call, drawn in layer order like the reference (networks.py:1015-1017, so the
        same torch seed gives the same noise as the ops engine), kept in the state for the backward pass."""
        mode = st["noise_mode"]
        if not L.has_noise or mode == "none":
            return None, 0
        if mode == "const":
            return L.noise, 0
        key = f"rnoise{L.idx}"
        if draw:
            st[key] = torch.randn([B, 1, H, Wd], device=self.dev)
        return st[key], H * Wd

    def forward_raw(self, ws, mask=None, noise_mode="const", want_probs=False, train=False):
        """want_probs: also keep every attention layer's probabilities [B, HW, 16] (fp32) in self.last_probs (attention maps).
        train: training-mode forward (attention dropout masks drawn per layer; the backward pass re-uses them)."""
        if noise_mode not in ("const", "none", "random"):
            raise ValueError("noise_mode must be 'random', 'const' or 'none'")
        B = ws.shape[0]
        st = self._state(B)
        st["batch"] = B
        st["want_probs"], st["probs"] = bool(want_probs), []
        st["train"] = bool(train)
        self.last_probs = st["probs"]
        ws = ws.detach().to(torch.float32).contiguous()
        st["ws"] = ws
        if mask is None:
            mask = torch.ones(B, self.k - 1, device=self.dev)
        maskbias = ((1.0 - mask.to(torch.float32)) * -10000.0).contiguous()
        st["maskbias"] = maskbias
        st["noise_mode"] = noise_mode
        st["noise_on"] = noise_on = noise_mode != "none"
        # per-layer operands that depend only on ws, for all layers, on the side stream (fork here, per-layer events, join at the end)
        main = torch.cuda.current_stream(self.dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            for e in self.blocks:
                for key in ("conv0", "conv1", "conv_last"):
                    if e.get(key) is not None:
                        self._prep_layer(e[key], ws, st, B)
        x, img = None, None
        for e in self.blocks:
            r = e["res"]
            if e["stem"]:
                x_in = self._buf(st, "const", (B,) + tuple(e["const"].shape), fwd=True)
                if x_in.dtype == torch.float16 and e["const_absmax"] > 65504.0:
                    raise _lib.MgfError("tc engine: the learned 4x4 constant (|max| %.3g) does not fit fp16 forward storage; "
                                        "select bf16: _lib.set_forward_dtype('bf16')" % e["const_absmax"])
                x_in.copy_(e["const"].unsqueeze(0).expand(B, -1, -1, -1))
                x = self._layer_fwd(e["conv1"], x_in, ws, maskbias, st, B, noise_on)
            else:
                x_in = x
                z0 = self._layer_fwd(e["conv0"], x_in, ws, maskbias, st, B, noise_on)
                z1 = self._layer_fwd(e["conv1"], z0, ws, maskbias, st, B, noise_on)
                if self.arch == "resnet":
                    O, I = e["conv0"].O, e["conv0"].I
                    h = x_in.shape[1]
                    v = self._buf(st, f"v{r}", (B, h, h, O), fwd=True)
                    if POINTWISE_SKIP and _L().mgf_pointwise_supported(I, O):      # HBM-bound levels: streaming mma.sync kernel (pointwise.cu)
                        _lib.check(_L().mgf_pointwise(_p(x_in), _p(self._skip_f(e)), _p(v), B * h * h, I, O, 1, _s(self.dev)), "mgf_pointwise")
                    else:
                        tc.conv_tc([x_in], self._skip_f(e), [(0, 0, 0, 0)], (B, h, h), 1, O, v, tag="g.fwd")
                    x = self._buf(st, f"xout{r}", (B, r, r, O), fwd=True)
                    _lib.check(_L().mgf_upfir2_add(_p(v), _p(z1), _p(x), e["fk4"], e["skip_gain"], B, h, h, O, _s(self.dev)), "mgf_upfir2_add")
                else:
                    x = z1
            if img is not None and self.arch == "skip":               # running image of the 'skip' architecture (:1166-1167), fp32 NCHW ops kernel
                img = _upfirdn2d.upsample2d(img, e["fir"])
            if e["last"]:
                yl = self._layer_fwd(e["conv_last"], x, ws, maskbias, st, B, noise_on)
                st["yl"] = yl
            if "rgb" in e:
                rgb = e["rgb"]
                xr = yl if e["last"] else x
                st[f"xrgb{r}"] = xr
                srgb = self._buf(st, f"s_rgb{r}", (B, rgb["C"]), torch.float32)
                wg = ws[:, -1, rgb["idx"]]
                _lib.check(_L().mgf_style_fwd(_p(wg), wg.stride(0), _p(rgb["A"]), _p(rgb["ab"]), rgb["again"], rgb["sgain"], None, _p(srgb), None,
                                              B, rgb["C"], 0, wg.shape[1], _s(self.dev)), "mgf_style_fwd")
                y_rgb = torch.empty(B, 3, r, r, dtype=torch.float32, device=self.dev)
                _lib.check(_L().mgf_torgb_fwd(_p(xr), _p(rgb["w"]), _p(srgb), _p(rgb["bias"]), _p(y_rgb), B, r * r, rgb["C"], _s(self.dev)), "mgf_torgb_fwd")
                img = y_rgb if img is None else img.add_(y_rgb)
        main.wait_stream(self._side)
        return img

    # -------------------------------------------------------------------------------------------- backward
    def _dgrad(self, L, dy, st, B, out, add=None, actgrad_X=None, ag_gain=1.0):
        """dy [B,H,W,O] (gradient wrt this layer's conv output) -> out [B,h,w,I] = d x_in; accumulates d(styles)."""
        d = st[f"d{L.idx}"]; s = st[f"s{L.idx}"]
        if L.superpix:
            Wb = None            # handled below with the block-expanded operands
        elif L.shared_w:     # dy*d on the (small) gradient tensor, shared transposed weights
            dyd = self._buf(st, f"dyd{L.idx}", tuple(dy.shape))
            _lib.check(_L().mgf_scale_channels(_p(dy), _p(d), _p(dyd), 0, B, dy.shape[1] * dy.shape[2], L.O, _s(self.dev)), "mgf_scale_channels")
            dy, Wb = dyd, L.Bb16
        else:
            Wb, ev = st[f"bprep{L.idx}"]                       # modulated on the side stream at the start of backward_raw
            torch.cuda.current_stream(self.dev).wait_event(ev)
        x_in = st[f"xin{L.idx}"]
        h, w = x_in.shape[1], x_in.shape[2]
        if L.superpix:
            assert add is None
            s2, d2 = st[f"s2_{L.idx}"], st[f"d2_{L.idx}"]
            Wb, ev = st[f"bprep{L.idx}"]
            torch.cuda.current_stream(self.dev).wait_event(ev)
            ds2 = self._zbuf(st, f"ds2_{L.idx}", (B, 64))
            tc.conv_tc([dy.view(B, h, w // 2, 64)], Wb, L.taps_b, (B, h, w // 2), 1, 64, out.view(B, h, w // 2, 64), scale_n=s2, reduce_out=ds2,
                       X=x_in.view(B, h, w // 2, 64), actgrad=actgrad_X is not None, ag_alpha=LRELU_ALPHA, ag_gain=ag_gain,
                       reduce_per_sample=True, alg_scale=0.5, tag="g.bwd", fwd=False)
            dsum = self._buf(st, f"ds{L.idx}", (B, 32), torch.float32)          # persistent: consumed later on the side stream
            torch.add(ds2[:, :32], ds2[:, 32:], out=dsum)
            return dsum
        ds = self._zbuf(st, f"ds{L.idx}", (B, L.I))
        if L.up == 1:
            acts = [dy]
        elif not L.two_stage:
            acts = [tc.phase_view(dy, py, px) for (py, px) in ((0, 0), (0, 1), (1, 0), (1, 1))]
        else:
            gq = self._buf(st, f"gfir{L.idx}", (B, 2 * h + 2, 2 * w + 2, L.O))
            _lib.check(_L().mgf_fir4_pad(_p(dy), _p(gq), FIR4_TAPS, 4.0, B, 2 * h, 2 * w, L.O, _s(self.dev)), "mgf_fir4_pad")
            acts = [tc.phase_view(gq, py, px) for (py, px) in ((0, 0), (0, 1), (1, 0), (1, 1))]
        tc.conv_tc(acts, Wb, L.taps_b, (B, h, w), 1, L.I, out, scale_n=s, reduce_out=ds, X=x_in, add=add,
                   actgrad=actgrad_X is not None, ag_alpha=LRELU_ALPHA, ag_gain=ag_gain, reduce_per_sample=True,
                   alg_scale=1.0 if (L.up == 1 or L.two_stage) else 0.25, tag="g.bwd", fwd=False)
        return ds

    def _on_side(self, fn):
        """Runs fn() on the side stream after everything queued so far on the current stream (small reductions into dws that nothing on
        the critical path waits for; backward_raw joins the side stream before it returns)."""
        if not BWD_SIDE_REDUCTIONS:
            return fn()
        main = torch.cuda.current_stream(self.dev)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            fn()

    def _prep_bwd(self, st, B):
        """Backward operands that depend only on the forward's demodulation factors: Wb[b] = W * d[b] for every non-shared layer."""
        for e in self.blocks:
            for key in ("conv0", "conv1", "conv_last"):
                L = e.get(key)
                if L is None or L.shared_w:
                    continue
                Wb = self._buf(st, f"Wb{L.idx}", (B,) + tuple(L.Bb.shape))
                self._modulate(L.Bb, None, 1, st[f"d2_{L.idx}"] if L.superpix else st[f"d{L.idx}"], Wb, B, False)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.dev))
                st[f"bprep{L.idx}"] = (Wb, ev)

    def _style_bwd(self, L, ds, R, st, dws, B):
        d = st[f"d{L.idx}"]; s = st[f"s{L.idx}"]
        dwg = dws[:, -1, L.idx]
        self._on_side(lambda: _lib.check(_L().mgf_style_bwd(_p(ds), _p(R), _p(s), _p(d), _p(L.Wsq), _p(L.A), L.again, 1.0, _p(dwg), dwg.stride(0),
                                                            B, L.I, L.O, dwg.shape[1], _s(self.dev)), "mgf_style_bwd"))

    def _attn_bwd(self, L, dz, st, dws, B):
        """dz: gradient wrt the layer output z.  Returns (dy, R): gradient wrt the conv output and sum_p dy*y."""
        y = st[f"y{L.idx}"]
        H, Wd = y.shape[1], y.shape[2]
        noise, nbs = self._noise_of(L, st, B, H, Wd)
        nstr = L.nstr if noise is not None else None
        dy = self._buf(st, f"dy{L.idx}", tuple(y.shape))
        dVM = self._zbuf(st, f"dVM{L.idx}", (B, 16, L.O))
        R = self._zbuf(st, f"R{L.idx}", (B, L.O))
        _lib.check(_L().mgf_attn_bwd(_p(y), _p(dz), _p(L.Kf), _p(L.Sc), _p(st["maskbias"]), _p(st[f"VM{L.idx}"]), _p(L.bm), _p(noise), _p(nstr),
                                     _p(L.bias), L.gain, LRELU_ALPHA, _p(dy), _p(dVM), _p(R), _p(st.get(f"dmask{L.idx}")), _p(L.tabK), _p(st[f"tabV{L.idx}"]),
                                     B, H * Wd, L.O, nbs, _s(self.dev)), "mgf_attn_bwd")
        dcomp = dws[:, :-1, L.idx]                                 # [B,16,32] strided, accumulate
        self._on_side(lambda: _lib.check(_L().mgf_small_gemm(_p(dVM), 16 * L.O, L.O, _p(L.WVMt), None, _p(dcomp), dcomp.stride(0), dcomp.stride(1),
                                                             B, 16, dcomp.shape[2], L.O, 1, _s(self.dev)), "mgf_small_gemm"))
        return dy, R

    def _act_bwd(self, L, dz, z, st, B, mode, want_dy=True):
        H, Wd = z.shape[1], z.shape[2]
        noise, nbs = self._noise_of(L, st, B, H, Wd)
        nstr = L.nstr if noise is not None else None
        dy = self._buf(st, f"dy{L.idx}", tuple(z.shape)) if want_dy else None
        R = self._zbuf(st, f"R{L.idx}", (B, L.O))

        def run():
            _lib.check(_L().mgf_act_bwd(_p(dz), _p(z), _p(dy), _p(R), _p(noise), _p(nstr), _p(L.bias), LRELU_ALPHA, L.gain, mode,
                                        B, H * Wd, L.O, nbs, _s(self.dev)), "mgf_act_bwd")
        if want_dy:
            run()
        else:             # reduce-only pass (R feeds only the d(style) reduction, which already lives on the side stream): off the critical path
            self._on_side(run)
        return dy, R

    @torch.no_grad()
    def backward_raw(self, dimg):
        """dimg [B,3,R,R] fp32 -> dws [B,k,num_ws,32] fp32 (gradient wrt the ws passed to the last forward_raw)."""
        B = dimg.shape[0]
        st = self._state(B)
        ws = st["ws"]
        dws = torch.zeros_like(ws)
        if st.get("zpool") is not None:
            st["zpool"][:max(st.get("zpool_used", 0), 1)].zero_()       # every reduction target of this backward pass, one fill
        dimg = dimg.to(torch.float32).contiguous()
        main = torch.cuda.current_stream(self.dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self._prep_bwd(st, B)
        g = None
        for e in reversed(self.blocks):
            r = e["res"]
            if "rgb" in e:
                # ToRGB backward of this block: gradient wrt its input (conv_last's output in the last block, the block output otherwise)
                rgb = e["rgb"]
                xr = st[f"xrgb{r}"]
                dxr = self._buf(st, f"dxrgb{r}", tuple(xr.shape))
                ds_rgb = self._zbuf(st, f"ds_rgb{r}", (B, rgb["C"]))
                Ll = e.get("conv_last")
                R_rgb = self._zbuf(st, f"R{Ll.idx}" if e["last"] else f"Rrgb{r}", (B, rgb["C"]))       # sum_p dy*y: only conv_last's demodulation needs it
                _lib.check(_L().mgf_torgb_bwd(_p(dimg), _p(xr), _p(rgb["w"]), _p(st[f"s_rgb{r}"]), _p(dxr), _p(ds_rgb), _p(R_rgb),
                                              B, r * r, rgb["C"], _s(self.dev)), "mgf_torgb_bwd")
                dwg = dws[:, -1, rgb["idx"]]
                self._on_side(lambda ds_rgb=ds_rgb, rgb=rgb, dwg=dwg: _lib.check(_L().mgf_style_bwd(
                    _p(ds_rgb), None, None, None, None, _p(rgb["A"]), rgb["again"], rgb["sgain"], _p(dwg), dwg.stride(0), B, rgb["C"], 0, dwg.shape[1],
                    _s(self.dev)), "mgf_style_bwd"))
                if e["last"]:
                    g_last = self._buf(st, f"g{r}", tuple(st[f"xin{Ll.idx}"].shape))
                    ds = self._dgrad(Ll, dxr, st, B, g_last)
                    self._style_bwd(Ll, ds, R_rgb, st, dws, B)
                    g = g_last
                else:
                    g.add_(dxr)                                    # 'skip' architecture: the block output also feeds its own ToRGB
                if self.arch == "skip" and not e["stem"]:
                    # gradient of the running image at the previous resolution: adjoint of upsample2d through the ops kernel's autograd
                    with torch.enable_grad():
                        a = torch.zeros(B, 3, r // 2, r // 2, device=self.dev, requires_grad=True)
                        up = _upfirdn2d.upsample2d(a, e["fir"])
                    (dimg,) = torch.autograd.grad(up, [a], grad_outputs=[dimg])
            if e["stem"]:
                L1 = e["conv1"]
                dy1, R1 = self._attn_bwd(L1, g, st, dws, B) if L1.attn else self._act_bwd(L1, g, st[f"z{L1.idx}"], st, B, 0)
                scratch = self._buf(st, "gconst", tuple(st[f"xin{L1.idx}"].shape))
                ds1 = self._dgrad(L1, dy1, st, B, scratch)
                self._style_bwd(L1, ds1, R1, st, dws, B)
                continue
            L0, L1 = e["conv0"], e["conv1"]
            x_in = st[f"xin{L0.idx}"]
            h = x_in.shape[1]
            resnet = self.arch == "resnet"
            gs = None
            if resnet:
                # skip branch: d v = FIR^T(g) ; d x_in (skip part) = conv1x1^T
                dv = self._buf(st, f"dv{r}", (B, h, h, L0.O))
                g_blk, fk4, sg = g, e["fk4"], e["skip_gain"]
                self._on_side(lambda: _lib.check(_L().mgf_upfir2_bwd(_p(g_blk), _p(dv), fk4, sg, B, h, h, L0.O, _s(self.dev)), "mgf_upfir2_bwd"))
                ev_dv = torch.cuda.Event(); ev_dv.record(self._side if BWD_SIDE_REDUCTIONS else main)
                gs = self._buf(st, f"gs{r}", tuple(x_in.shape))
            # conv1
            z0 = st[f"z{L0.idx}"]
            dy1, R1 = self._attn_bwd(L1, g, st, dws, B) if L1.attn else self._act_bwd(L1, g, st[f"z{L1.idx}"], st, B, 0)
            dz0 = self._buf(st, f"dz{L0.idx}", tuple(z0.shape))
            if L0.attn:
                ds1 = self._dgrad(L1, dy1, st, B, dz0)
                dy0, R0 = self._attn_bwd(L0, dz0, st, dws, B)
            else:
                # fuse conv0's leaky-ReLU backward into conv1's dgrad epilogue (X = z0 is both conv1's input and conv0's output)
                ds1 = self._dgrad(L1, dy1, st, B, dz0, actgrad_X=z0, ag_gain=L0.gain)
                dy0 = dz0
                _, R0 = self._act_bwd(L0, dy0, z0, st, B, 1, want_dy=False)
            self._style_bwd(L1, ds1, R1, st, dws, B)
            # conv0 (up) dgrad, adding the skip-branch gradient in the epilogue -> gradient wrt the previous block's output
            gprev = self._buf(st, f"g{r // 2}", tuple(x_in.shape))
            if resnet:
                main.wait_event(ev_dv)                         # the FIR adjoint of the skip branch ran beside conv1's backward
                if POINTWISE_SKIP and _L().mgf_pointwise_supported(L0.O, L0.I):
                    _lib.check(_L().mgf_pointwise(_p(dv), _p(e["skip_b"]), _p(gs), B * h * h, L0.O, L0.I, 0, _s(self.dev)), "mgf_pointwise")
                else:
                    tc.conv_tc([dv], e["skip_b"], [(0, 0, 0, 0)], (B, h, h), 1, L0.I, gs, tag="g.bwd", fwd=False)
            ds0 = self._dgrad(L0, dy0, st, B, gprev, add=gs)
            self._style_bwd(L0, ds0, R0, st, dws, B)
            g = gprev
        main.wait_stream(self._side)
        return dws

    # -------------------------------------------------------------------------------------------- autograd entry
    def _param_state(self):
        return tuple((id(p), p._version) for p in list(self.net.parameters()) + list(self.net.buffers()))

    def __call__(self, ws, pos=None, mask=None, noise_mode="const", fused_modconv=None, want_probs=False, **_ignored):
        """Public entry (G.synthesis with engine='tc').  Training mode (reference networks.py:505-513, :1010-1017): attention dropout and,
        with noise_mode='random', fresh noise planes per call; gradients flow to ws only (the engine folds the weights: it serves projection,
        path-length style regularisers and evaluation, not weight updates -- use engine='ops' to train the weights).  The folded weights are
        rebuilt automatically when a parameter changed in place since the last call (optimizer steps, load_state_dict)."""
        state = self._param_state()
        if state != getattr(self, "_folded_state", None):
            if getattr(self, "_folded_state", None) is not None:
                self.refresh()
            self._folded_state = state
        img = _SynthesisFn.apply(ws, self, mask, noise_mode, want_probs, bool(self.net.training))
        if _lib.forward_torch_dtype() == torch.float16 and not torch.cuda.is_current_stream_capturing():
            _lib.check_fp16_overflow(self.dev, "G.synthesis (tc engine)")     # public entry: a clipped image must not pass silently
        return img


class _SynthesisFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ws, eng, mask, noise_mode, want_probs, train):
        ctx.eng = eng
        ctx.B = ws.shape[0]
        return eng.forward_raw(ws, mask=mask, noise_mode=noise_mode, want_probs=want_probs, train=train)

    @staticmethod
    def backward(ctx, dimg):
        return ctx.eng.backward_raw(dimg), None, None, None, None, None


def smoke():
    """tiny tc-engine forward+backward on cuda:0 (called from __graft_entry__.smoke): fp16-forward mode, seeded, image within 1e-2 max-abs
    (absolute, the north_star bar; measured 8.35e-3 with ws drawn directly, image range 4.1) of the exact-fp32 ops engine."""
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    import util
    G = util.build_G(64, 0, 2048, 64).cuda()
    ws = util.case_tensor((2, 17, G.num_ws, 32), 60).cuda().requires_grad_(True)
    mask = torch.ones(2, 16, device="cuda")
    _lib.set_forward_dtype("fp16")
    try:
        G.synthesis.engine = "tc"
        img, _ = G.synthesis(ws, pos=G.pos, mask=mask, noise_mode="const")
        img.square().mean().backward()
        torch.cuda.synchronize()
    finally:
        _lib.set_forward_dtype(_lib.DEFAULT_FORWARD_DTYPE)
    G.synthesis.engine = "ops"
    ref, _ = G.synthesis(ws.detach(), pos=G.pos, mask=mask, noise_mode="const", return_att_maps=False)
    err, rng = (img.detach() - ref).abs().max().item(), max(1.0, ref.abs().max().item())
    assert err < 1e-2, "tc engine deviates from the ops engine: %g max-abs (range %g)" % (err, rng)
    print("smoke ok: tc engine 64x64 fwd+bwd, max|img_tc - img_fp32| = %.3g (range %.3g), |dws| = %.3g" % (err, rng, ws.grad.abs().max().item()))

"""Per-image latent projection: Adam on the z latents [B,17,32] against lamda * LPIPS-VGG + (1 - lamda) * MSE.

Restates the loop of the reference scripts (1024_example_percept_MSE.py:113-175; get_lr :62-67, latent_noise :70-72, latent
statistics :212-216) with the gradient actually flowing through G (the reference detaches the image through NumPy, SURVEY.md
section 0-4 -- the intended loop is built here, matching oracle/projection.py).  All images of a batch are independent jobs
(per-image mean MSE, per-image LPIPS), so a batch can be sharded over GPUs with no per-step collective.

One step = mapping network (mgf_mapping_fwd; PyTorch module for non-default mapping configurations) -> tcgen05 synthesis engine forward -> LPIPS/MSE forward ->
LPIPS/MSE backward -> synthesis backward -> mapping backward -> fused Adam + next-step latent noise (mgf_adam_noise_step).
No host synchronisation inside a step: the lr / noise schedule lives in a device array indexed by a device step counter.
"""
import math
import torch
from . import _lib
from .lpips_engine import LpipsEngine


def get_lr(t, initial_lr, rampdown=0.25, rampup=0.05):
    lr_ramp = min(1.0, (1.0 - t) / rampdown)
    lr_ramp = 0.5 - 0.5 * math.cos(lr_ramp * math.pi)
    return initial_lr * lr_ramp * min(1.0, t / rampup)


def latent_stats(noise_sample):
    mean = noise_sample.mean(0)
    std = ((noise_sample - mean).pow(2).sum() / noise_sample.shape[0]) ** 0.5
    return mean, std


class Projector:
    def __init__(self, G, lpips_state_dict, batch, steps, lr=0.1, lamda=0.5, noise=0.05, noise_ramp=0.75, lr_rampdown=0.25,
                 lr_rampup=0.05, weight_decay=1e-4, latent_mean=None, latent_std=None, use_lpips=True, step_noise=None,
                 noise_seed=3, forward_dtype=None, fused_mapping=True, engine="tc", noise_mode="const"):
        """forward_dtype: None keeps the library's current setting; 'fp16' / 'bf16' select the 16-bit type of the engine's forward
        activations and operands (gradients are always bf16).  fp16 meets the 1e-2 image / 1e-3 loss parity bars; bf16 has the
        fp32 exponent range (use it for checkpoints whose activations may exceed 6.5e4).  Same speed."""
        if forward_dtype is not None:
            _lib.set_forward_dtype(forward_dtype)
        if engine not in ("tc", "ops"):
            raise ValueError("engine must be 'tc' (16-bit tcgen05 engine, the throughput path) or 'ops' (exact fp32 kernels + autograd)")
        if noise_mode not in ("const", "random", "none"):
            raise ValueError("noise_mode must be 'const', 'random' or 'none'")
        self.engine, self.noise_mode = engine, noise_mode      # 'random' is what the reference scripts get by default (networks.py:1010); 'const' is reproducible
        self.G = G
        self.dev = next(G.parameters()).device
        if self.dev.type != "cuda":
            raise _lib.MgfError("Projector needs the generator on a CUDA device (no CPU fallback)")
        G.synthesis.engine = engine
        G.eval().requires_grad_(False)
        self.B, self.steps, self.lamda, self.use_lpips = batch, steps, float(lamda), use_lpips
        self.wd = float(weight_decay)
        k, zd = G.k, G.z_dim
        if latent_mean is None:
            g = torch.Generator(device="cpu").manual_seed(1234)
            latent_mean, latent_std = latent_stats(torch.randn(10000, k, zd, generator=g))
        self.latent_mean = latent_mean.to(self.dev, torch.float32)
        self.latent_std = float(latent_std)
        # schedule: sched[i] = (lr of step i, noise strength of step i+1)
        sched = []
        for i in range(steps):
            t, t1 = i / steps, (i + 1) / steps
            ns1 = self.latent_std * noise * max(0.0, 1.0 - t1 / noise_ramp) ** 2
            sched += [get_lr(t, lr, lr_rampdown, lr_rampup), ns1]
        self.ns0 = self.latent_std * noise
        self.sched = torch.tensor(sched, dtype=torch.float32, device=self.dev)
        if step_noise is None:
            g = torch.Generator(device=self.dev).manual_seed(noise_seed)
            step_noise = torch.randn(steps, batch, k, zd, generator=g, device=self.dev)
        if step_noise.dim() != 4 or step_noise.shape[0] < steps or tuple(step_noise.shape[1:]) != (batch, k, zd):
            raise ValueError("step_noise must be [>= steps, batch, k, z_dim] = [>= %d, %d, %d, %d] (the Adam kernel indexes it by the device step counter), got %s"
                             % (steps, batch, k, zd, tuple(step_noise.shape)))
        self.step_noise = step_noise.to(self.dev, torch.float32).contiguous()
        self.lp = LpipsEngine(lpips_state_dict, self.dev) if (use_lpips and engine == "tc") else None
        self.lp32 = None
        if use_lpips and engine == "ops":          # exact-fp32 LPIPS-VGG16 on the ops kernels (autograd)
            from .lpips_nets import LpipsNet
            self.lp32 = LpipsNet(lpips_state_dict, "vgg").to(self.dev)
        # the mapping network and its backward as two kernels when G has the GANformer-default mapping (else the PyTorch module + autograd)
        from . import mapping_engine
        self.mapper = mapping_engine.MappingEngine(G) if (fused_mapping and engine == "tc" and mapping_engine.supported(G)) else None
        self.mask = torch.ones(batch, k - 1, device=self.dev)
        self.reset()

    def reset(self):
        B = self.B
        if getattr(self, "latent", None) is None:      # allocate once: addresses stay fixed (CUDA-graph replay, repeated jobs)
            self.latent = torch.empty(B, *self.latent_mean.shape, device=self.dev)
            self.m, self.v, self.latent_n = torch.empty_like(self.latent), torch.empty_like(self.latent), torch.empty_like(self.latent)
            self.best_loss = torch.empty(B, device=self.dev)
            self.best_latent = torch.empty_like(self.latent)
        self.latent.copy_(self.latent_mean.unsqueeze(0).expand(B, -1, -1))
        self.m.zero_(); self.v.zero_()
        self.latent_n.copy_(self.latent + self.step_noise[0] * self.ns0)
        if getattr(self, "step_ctr", None) is None:
            self.step_ctr = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self.step_idx = torch.zeros(1, dtype=torch.int64, device=self.dev)
            self.losses = torch.zeros(self.steps, B, device=self.dev)
            self.graph = None
        self.step_ctr.zero_(); self.losses.zero_()
        self.best_loss.fill_(float("inf"))
        self.best_latent.copy_(self.latent_n)
        self.i = 0

    def set_targets(self, target):
        """target [B,3,R,R] fp32 in [-1,1] (device or host; a pinned host tensor is uploaded asynchronously straight into the
        resident target buffer, no intermediate device allocation)."""
        assert target.shape[0] == self.B
        if getattr(self, "target", None) is not None and tuple(self.target.shape) == tuple(target.shape):
            self.target.copy_(target, non_blocking=True)   # same buffer: a captured graph keeps working for the next job
        else:
            self.target = target.to(self.dev, torch.float32).contiguous().clone()
            self.graph = None
        if self.lp is not None:
            self.lp.set_target(self.target)
        if getattr(self, "coef", None) is None:
            self.coef = torch.full((self.B,), self.lamda if self.use_lpips else 0.0, device=self.dev)

    def _loss_and_grad(self, img):
        R = img.shape[2]
        n = 3 * R * R
        if self.use_lpips:
            val, mse_sum = self.lp.forward(img)
            per_img = self.lamda * val + (1 - self.lamda) * mse_sum / n
            dimg = self.lp.backward(self.coef, (1 - self.lamda) * 2.0 / n)
        else:
            lib, s = _lib.lib(), _lib.stream_ptr(self.dev)
            mse_sum = torch.zeros(self.B, device=self.dev)
            _lib.check(lib.mgf_lpips_prep(img.data_ptr(), self.target.data_ptr(), None, mse_sum.data_ptr(), self.B, R, s), "mgf_lpips_prep")
            per_img = mse_sum / n
            dimg = torch.empty_like(img)
            _lib.check(lib.mgf_lpips_prep_bwd(None, img.data_ptr(), self.target.data_ptr(), 2.0 / n, dimg.data_ptr(), self.B, R, s), "mgf_lpips_prep_bwd")
        return per_img, dimg

    def step(self, use_graph=True):
        """One projection step for the whole batch; no host sync.  Returns the per-image loss tensor (device).
        After capture() the step is one CUDA-graph replay (use_graph=False forces the eager launch sequence)."""
        if self.i >= self.steps:
            raise RuntimeError("Projector: step %d is past the %d-step schedule this projector was built for (reset() starts a new job)" % (self.i, self.steps))
        if use_graph and self.graph is not None:
            self.graph.replay()
            self.i += 1
            return self._graph_loss
        return self._step_eager()

    def capture(self):
        """Captures one step (mapping fwd/bwd in PyTorch + every mgf kernel launch) into a CUDA graph.  The schedule and
        the loss row are indexed by the device step counter, so replays advance exactly like eager steps."""
        saved = [t.clone() for t in (self.latent, self.m, self.v, self.latent_n, self.best_loss, self.best_latent, self.step_ctr, self.losses)]
        i_saved = self.i
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(2):                     # warm-up: allocates every persistent buffer, sets kernel attributes
                self._step_eager()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.lib().mgf_launch_count()
        with torch.cuda.graph(g):
            self._graph_loss = self._step_eager()
        self.launches_per_step = int(_lib.lib().mgf_launch_count() - n0)   # mgf kernels recorded in one step
        for dst, src in zip((self.latent, self.m, self.v, self.latent_n, self.best_loss, self.best_latent, self.step_ctr, self.losses), saved):
            dst.copy_(src)
        self.i = i_saved
        self.graph = g
        return g

    def _step_ops(self):
        """The same step on the exact-fp32 path: ops-engine synthesis + fp32 LPIPS, gradients by autograd (parity reference on the GPU:
        slow, bit-for-bit the oracle's arithmetic up to summation order)."""
        G = self.G
        z = self.latent_n.detach().requires_grad_(True)
        with torch.enable_grad():
            img = G(z, noise_mode=self.noise_mode)[0]
            per_img = (img - self.target).square().mean(dim=[1, 2, 3])
            if self.use_lpips:
                per_img = self.lamda * self.lp32(img, self.target).reshape(-1) + (1 - self.lamda) * per_img
            (gz,) = torch.autograd.grad(per_img.sum(), [z])
        return per_img.detach(), gz, img.detach()

    def _step_eager(self):
        if self.engine == "ops":
            per_img, gz, img = self._step_ops()
            return self._finish_step(per_img, gz, img)
        G = self.G
        eng = self._engine()
        if self.mapper is not None:
            ws = self.mapper.forward(self.latent_n, self.mask)
        else:
            z = self.latent_n.detach().requires_grad_(True)
            with torch.enable_grad():
                ws = G.mapping(z, None, pos=G.pos, mask=self.mask)
        img = eng.forward_raw(ws, mask=self.mask, noise_mode=self.noise_mode)
        per_img, dimg = self._loss_and_grad(img)
        dws = eng.backward_raw(dimg)
        if self.mapper is not None:
            gz = self.mapper.backward(dws)
        else:
            (gz,) = torch.autograd.grad(ws, [z], grad_outputs=[dws])
        return self._finish_step(per_img, gz, img)

    def _finish_step(self, per_img, gz, img):
        # best-so-far bookkeeping (reference: keep latent_n of the lowest-loss step, :155-158)
        better = per_img < self.best_loss
        self.best_loss.copy_(torch.where(better, per_img, self.best_loss))
        self.best_latent.copy_(torch.where(better.reshape(-1, 1, 1), self.latent_n, self.best_latent))
        self.step_idx.copy_(self.step_ctr)
        self.losses.index_copy_(0, self.step_idx.clamp(max=self.steps - 1), per_img.unsqueeze(0))
        lib = _lib.lib()
        _lib.check(lib.mgf_adam_noise_step(self.latent.data_ptr(), gz.contiguous().data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                           self.step_noise.data_ptr(), self.steps, self.latent_n.data_ptr(), self.sched.data_ptr(),
                                           self.step_ctr.data_ptr(), 0.9, 0.999, 1e-8, self.wd, self.latent.numel(),
                                           _lib.stream_ptr(self.dev)), "mgf_adam_noise_step")
        self.i += 1
        self.last_img = img
        return per_img

    def _engine(self):
        syn = self.G.synthesis
        if syn._tc is None:
            from . import engine as _engine
            syn._tc = _engine.SynthesisEngine(syn)
        return syn._tc

    def run(self, steps=None):
        n = self.steps - self.i if steps is None else int(steps)
        if self.i + n > self.steps:
            raise RuntimeError("Projector.run: %d more steps from step %d exceed the %d-step schedule" % (n, self.i, self.steps))
        for _ in range(n):
            self.step()
        if self.engine == "tc" and _lib.forward_torch_dtype() == torch.float16:
            _lib.check_fp16_overflow(self.dev, "Projector.run")      # one stream sync at the end of the job, none inside the steps
        return dict(latent=self.latent, best_latent=self.best_latent, best_loss=self.best_loss, losses=self.losses)


def interpolate_pair(G, z1, z2, alpha=0.5):
    """Pair morph (projection_example_v2_percept_morph.py:356-363): W = (1-alpha) z1 + alpha z2, then one forward."""
    z = (1 - alpha) * z1 + alpha * z2
    with torch.no_grad():
        return G(z, noise_mode="const")[0]

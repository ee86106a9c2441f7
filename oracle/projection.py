"""TEST INFRASTRUCTURE ONLY (never imported by the product package).

Restatement of the reference's per-image latent projection loop (1024_example_percept_MSE.py:113-175,
get_lr :62-67, latent_noise :70-72, latent statistics :212-216) with the corrections SURVEY.md 8c lists:
  (i)  the `.cpu().detach().numpy()` round trip at :137-140 and `mse_loss.requires_grad = True` at :144 are
       dropped, so the gradient really flows through G (the north_star's intended loop);
  (ii) truncation_psi is 1 (the reference passes it positionally into `c`, networks.py:1304);
  (iii) the per-step randn_like noise and the 10000-sample latent statistics come from injected tensors;
  (iv) noise_mode='const'.
The per-image loss is lamda * LPIPS_i + (1 - lamda) * mean_i((img - target)^2), summed over the batch for
backward: per-image terms only, so results do not depend on how images are sharded over GPUs
(the reference's MSELoss(reduction='mean') over the whole batch is the same thing at batch 1, its only
batch size).  Adam is torch.optim.Adam(betas .9/.999, eps 1e-8, weight_decay 1e-4 coupled L2), :117.
parity unpinned: the reference has no golden vectors for this loop; it is anchored on the call sites above.
"""
import math
import torch
from . import ganformer, lpips_ref


def get_lr(t, initial_lr, rampdown=0.25, rampup=0.05):
    r = min(1.0, (1.0 - t) / rampdown)
    r = 0.5 - 0.5 * math.cos(r * math.pi)
    return initial_lr * r * min(1.0, t / rampup)


def latent_stats(noise_sample):
    """:212-216 -- mean over samples [k,32]; std is a single scalar sqrt(sum((x-mean)^2)/n)."""
    mean = noise_sample.mean(0)
    std = ((noise_sample - mean).pow(2).sum() / noise_sample.shape[0]) ** 0.5
    return mean, std


def noise_strength(t, latent_std, noise=0.05, noise_ramp=0.75):
    return float(latent_std) * noise * max(0.0, 1.0 - t / noise_ramp) ** 2


def project(g_sd, lpips_sd, target, latent_mean, latent_std, step_noise, res, steps, lr=0.1, lamda=0.5,
            noise=0.05, noise_ramp=0.75, rampdown=0.25, rampup=0.05, weight_decay=1e-4, use_lpips=True,
            dtype=torch.float32, total_steps=None):
    """target [B,3,R,R]; step_noise [steps,B,k,32] (unit normal).  Returns dict(latent, losses [steps,B])."""
    total_steps = total_steps or steps
    b = target.shape[0]
    latent = latent_mean.detach().clone().unsqueeze(0).repeat(b, 1, 1).to(dtype).requires_grad_(True)
    opt = torch.optim.Adam([latent], lr=lr, weight_decay=weight_decay)
    losses, lat_n_hist = [], []
    target = target.to(dtype)
    for i in range(steps):
        t = i / total_steps
        opt.param_groups[0]["lr"] = get_lr(t, lr, rampdown, rampup)
        ns = noise_strength(t, latent_std, noise, noise_ramp)
        latent_n = latent + step_noise[i].to(dtype) * ns
        img, _ = ganformer.generator(g_sd, latent_n, res, dtype=dtype)
        mse = (img - target).pow(2).mean(dim=[1, 2, 3])
        if use_lpips:
            p = lpips_ref.lpips(lpips_sd, img, target).reshape(b)
            per_img = lamda * p + (1 - lamda) * mse
        else:
            per_img = mse
        opt.zero_grad()
        per_img.sum().backward()
        opt.step()
        losses.append(per_img.detach().clone())
        lat_n_hist.append(latent_n.detach().clone())
    # latent_n [steps,B,k,32]: the noisy latent each step's loss was evaluated at (lets a test feed the same inputs step by step)
    return dict(latent=latent.detach().clone(), losses=torch.stack(losses), latent_n=torch.stack(lat_n_hist))

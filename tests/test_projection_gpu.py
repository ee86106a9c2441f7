"""GPU parity: LPIPS engine and the projection loop against the oracle restatements (CPU fp32)."""
import numpy as np
import pytest
import torch
from oracle import ganformer, lpips_ref, projection as oproj
import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fwd", ["fp16", "bf16"])
def test_lpips_engine_forward_backward(fwd):
    """LPIPS-VGG16 value and image gradient against the oracle.  fp16 forward storage (library default): value within 2e-3 relative,
    gradient within 2e-2 relative L2 (bf16 gradient tensors through 13 convolutions) -- tight enough that a wrong tap weight (each tap carries
    10-35 % of the distance) or a wrong pooling route cannot pass.  bf16 forward storage: its measured envelope."""
    from morphganformer_b200.lpips_engine import LpipsEngine
    from morphganformer_b200 import _lib
    _lib.set_forward_dtype(fwd)
    sd = util.build_vgg_lpips_sd(4)
    B, R = 2, 64
    a = torch.tanh(util.case_tensor((B, 3, R, R), 60)).requires_grad_(True)
    b = torch.tanh(util.case_tensor((B, 3, R, R), 61))
    d = lpips_ref.lpips(sd, a, b).reshape(B)
    mse = (a - b).square().mean(dim=[1, 2, 3])
    loss = 0.5 * d + 0.5 * mse
    ga, = torch.autograd.grad(loss.sum(), [a])
    eng = LpipsEngine(sd)
    eng.set_target(b.cuda())
    val, mse_sum = eng.forward(a.detach().cuda())
    n = 3 * R * R
    np.testing.assert_allclose(val.cpu().numpy(), d.detach().numpy(), rtol=2e-3 if fwd == "fp16" else 2e-2)
    np.testing.assert_allclose((mse_sum / n).cpu().numpy(), mse.detach().numpy(), rtol=1e-5)
    dimg = eng.backward(torch.full((B,), 0.5, device="cuda"), 0.5 * 2.0 / n).cpu()
    cos = torch.nn.functional.cosine_similarity(dimg.flatten(), ga.flatten(), dim=0).item()
    rel_l2 = ((dimg - ga).norm() / ga.norm()).item()
    print("lpips val", val.cpu().tolist(), d.tolist(), "grad cos", cos, "rel-L2", rel_l2, "norm ratio", (dimg.norm() / ga.norm()).item())
    if fwd == "fp16":
        assert rel_l2 < 2e-2 and cos > 0.9998, (rel_l2, cos)
    else:
        assert cos > 0.99 and abs((dimg.norm() / ga.norm()).item() - 1) < 0.03


def test_perceptual_loss_api_autograd():
    from morphganformer_b200.lpips_engine import PerceptualLoss
    sd = util.build_vgg_lpips_sd(4)
    a = torch.tanh(util.case_tensor((1, 3, 32, 32), 1)).cuda().requires_grad_(True)
    b = torch.tanh(util.case_tensor((1, 3, 32, 32), 2)).cuda()
    p = PerceptualLoss(sd)
    out = p(a, b)
    assert tuple(out.shape) == (1, 1, 1, 1)
    out.sum().backward()
    assert a.grad is not None and torch.isfinite(a.grad).all() and a.grad.abs().max() > 0


@pytest.mark.parametrize("use_lpips,fwd", [(False, "bf16"), (True, "bf16"), (True, "fp16")])
def test_projection_loss_trajectory_matches_oracle(use_lpips, fwd):
    """N-step loss trajectory with injected noise.  The bf16 engine tracks the fp32 oracle to ~1.5e-2 relative
    per-step loss on this net (bf16 rounding of every stored activation; all 12 losses share one latent so the deviation is
    common-mode); asserted bound 3e-2.  The north_star's 1e-3 is met by the exact-fp32 ops path (test_synthesis_gpu.py),
    not yet by the bf16 engine -- DESIGN.md 'precision'."""
    from morphganformer_b200.projection import Projector, latent_stats
    from morphganformer_b200 import _lib
    _lib.set_forward_dtype(fwd)
    res, cb, cm, B, steps = 64, 2048, 64, 2, 6
    G = util.build_G(res, 0, cb, cm)
    gsd = util.state_dict_cpu(G)
    lsd = util.build_vgg_lpips_sd(4)
    mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
    noise = util.case_tensor((steps, B, 17, 32), 71)
    with torch.no_grad():
        tgt, _ = ganformer.generator(gsd, util.case_tensor((B, 17, 32), 72), res)
        tgt = torch.tanh(tgt)
    ref = oproj.project(gsd, lsd, tgt, mean, std, noise, res, steps, use_lpips=use_lpips, total_steps=50)
    P = Projector(G.cuda(), lsd, B, 50, latent_mean=mean, latent_std=std, use_lpips=use_lpips, step_noise=torch.cat([noise, torch.zeros(44, B, 17, 32)]))
    P.set_targets(tgt)
    for _ in range(steps):
        P.step()
    torch.cuda.synchronize()
    got = P.losses[:steps].cpu()
    _lib.set_forward_dtype(_lib.DEFAULT_FORWARD_DTYPE)
    print("oracle", ref["losses"].flatten().tolist())
    print("engine", fwd, got.flatten().tolist())
    np.testing.assert_allclose(got.numpy(), ref["losses"].numpy(), rtol=3e-2 if fwd == "bf16" else 4e-3)
    dl = (P.latent.cpu() - ref["latent"]).abs().max().item()
    print("latent max diff after %d steps: %g" % (steps, dl))
    assert dl < 0.15


@pytest.mark.parametrize("use_lpips,steps", [(False, 40), (True, 12)])
def test_per_step_loss_within_1e3_along_the_whole_oracle_trajectory(use_lpips, steps):
    """north_star: 'per-step loss within 1e-3 relative'.  A free-running trajectory mixes two things: the error of one step and the slow
    divergence of the latents under Adam (asserted separately above).  Here every step is evaluated at the ORACLE's noisy latent of that step
    (oracle/projection.py returns them), so each of the `steps` losses is a like-for-like per-step comparison: all within 1e-3."""
    from morphganformer_b200.projection import Projector, latent_stats
    res, cb, cm, B = 64, 2048, 64, 2
    G = util.build_G(res, 0, cb, cm)
    gsd = util.state_dict_cpu(G)
    lsd = util.build_vgg_lpips_sd(4) if use_lpips else None
    mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
    noise = util.case_tensor((steps, B, 17, 32), 71)
    with torch.no_grad():
        tgt = torch.tanh(ganformer.generator(gsd, util.case_tensor((B, 17, 32), 72), res)[0])
    ref = oproj.project(gsd, lsd, tgt, mean, std, noise, res, steps, use_lpips=use_lpips, total_steps=100)
    P = Projector(G.cuda(), lsd, B, 100, latent_mean=mean, latent_std=std, use_lpips=use_lpips, step_noise=torch.zeros(100, B, 17, 32))
    P.set_targets(tgt)
    got = []
    for i in range(steps):
        P.latent_n.copy_(ref["latent_n"][i])
        got.append(P.step(use_graph=False).cpu().clone())
    got = torch.stack(got)
    rel = ((got - ref["losses"]).abs() / ref["losses"].abs())
    print("per-step loss rel err along the oracle trajectory: max %.2e, mean %.2e (loss %.4f -> %.4f)"
          % (rel.max().item(), rel.mean().item(), ref["losses"][0].mean().item(), ref["losses"][-1].mean().item()))
    assert rel.max().item() <= 1e-3


def test_1000_step_projected_latents_tc_engine_vs_oracle():
    """north_star: '1000-step projected latents within a stated tolerance', for the THROUGHPUT engine (tcgen05, fp16 forward storage, bf16
    gradients), 1000 Adam steps at 32^2 with the oracle's injected noise.  Adam divides by sqrt(v): once the gradient is small its direction
    is set by rounding, so free-running latents of two implementations separate (the exact-fp32 path drifts 8e-3 from fp32 summation order
    alone).  Stated tolerance for the 16-bit engine: latent RMS difference <= 0.1 (|z| ~ 1), max-abs <= 0.5, final loss and best loss within
    3e-2 relative of the oracle's, per-step loss within 1e-3 over the first 10 steps.  Measured values are printed."""
    from morphganformer_b200.projection import Projector, latent_stats
    res, B, steps = 32, 2, 1000
    G = util.build_G(res, 0, 1024, 32)
    gsd, lsd = util.state_dict_cpu(G), util.build_vgg_lpips_sd(4)
    mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
    noise = util.case_tensor((steps, B, 17, 32), 71)
    with torch.no_grad():
        tgt = torch.tanh(ganformer.generator(gsd, util.case_tensor((B, 17, 32), 72), res)[0])
    ref = oproj.project(gsd, lsd, tgt, mean, std, noise, res, steps)
    P = Projector(G.cuda(), lsd, B, steps, latent_mean=mean, latent_std=std, step_noise=noise)
    P.set_targets(tgt)
    P.capture()
    P.run(steps)
    torch.cuda.synchronize()
    got_l, want_l = P.losses.cpu(), ref["losses"]
    d = P.latent.cpu() - ref["latent"]
    rel = ((got_l - want_l).abs() / want_l.abs())
    best_rel = ((P.best_loss.cpu() - want_l.min(dim=0).values).abs() / want_l.min(dim=0).values).max().item()
    print("tc engine 1000 steps: latent drift max-abs %.3g rms %.3g; loss rel err first-10 max %.2e, step 99 %.2e, step 999 %.2e; best-loss rel %.2e; loss %.4f -> %.4f"
          % (d.abs().max().item(), d.square().mean().sqrt().item(), rel[:10].max().item(), rel[99].max().item(), rel[999].max().item(), best_rel,
             want_l[0].mean().item(), want_l[-1].mean().item()))
    assert rel[:10].max().item() <= 1e-3
    assert d.square().mean().sqrt().item() <= 0.1 and d.abs().max().item() <= 0.5
    assert rel[999].max().item() <= 3e-2 and best_rel <= 3e-2
    assert want_l[-1].mean() < want_l[0].mean()


def test_cuda_graph_replay_equals_eager():
    from morphganformer_b200.projection import Projector, latent_stats
    res, cb, cm, B, steps = 64, 2048, 64, 2, 5
    lsd = util.build_vgg_lpips_sd(4)
    mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
    noise = util.case_tensor((20, B, 17, 32), 71)
    tgt = torch.tanh(util.case_tensor((B, 3, res, res), 73))
    out = []
    for graph in (False, True):
        G = util.build_G(res, 0, cb, cm).cuda()
        P = Projector(G, lsd, B, 20, latent_mean=mean, latent_std=std, step_noise=noise)
        P.set_targets(tgt)
        if graph:
            P.capture()
        for _ in range(steps):
            P.step()
        torch.cuda.synchronize()
        out.append((P.losses[:steps].cpu(), P.latent.cpu(), P.best_loss.cpu()))
    np.testing.assert_allclose(out[1][0].numpy(), out[0][0].numpy(), rtol=2e-3)     # float atomics reorder between runs
    assert (out[1][1] - out[0][1]).abs().max() < 2e-2
    np.testing.assert_allclose(out[1][2].numpy(), out[0][2].numpy(), rtol=2e-3)


def test_fused_mapping_equals_pytorch_mapping_path():
    """The projection step with the mapping network on mgf_mapping_fwd/bwd equals the step that keeps the PyTorch module + autograd."""
    from morphganformer_b200.projection import Projector, latent_stats
    res, cb, cm, B, steps = 64, 2048, 64, 2, 4
    lsd = util.build_vgg_lpips_sd(4)
    mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
    noise = util.case_tensor((20, B, 17, 32), 71)
    tgt = torch.tanh(util.case_tensor((B, 3, res, res), 73))
    out = []
    for fused in (True, False):
        G = util.build_G(res, 0, cb, cm).cuda()
        P = Projector(G, lsd, B, 20, latent_mean=mean, latent_std=std, step_noise=noise, fused_mapping=fused)
        assert (P.mapper is not None) == fused
        P.set_targets(tgt)
        P.run(steps)
        torch.cuda.synchronize()
        out.append((P.losses[:steps].cpu(), P.latent.cpu()))
    np.testing.assert_allclose(out[0][0].numpy(), out[1][0].numpy(), rtol=2e-3)
    assert (out[0][1] - out[1][1]).abs().max() < 2e-2


@pytest.mark.parametrize("res,B,cb,cm", [(32, 1, 1024, 32), (128, 3, 32768, 512)])
def test_projection_first_steps_other_resolutions(res, B, cb, cm):
    """Projection (MSE + LPIPS-VGG) at the smallest supported resolution and at a default-channel 128^2 generator with an odd batch:
    the first two per-step losses against the oracle loop (fp16-forward engine)."""
    from morphganformer_b200.projection import Projector, latent_stats
    from morphganformer_b200 import _lib
    steps = 2
    G = util.build_G(res, 0, cb, cm)
    gsd, lsd = util.state_dict_cpu(G), util.build_vgg_lpips_sd(4)
    mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
    noise = util.case_tensor((steps, B, 17, 32), 71)
    with torch.no_grad():
        tgt = torch.tanh(ganformer.generator(gsd, util.case_tensor((B, 17, 32), 72), res)[0])
    ref = oproj.project(gsd, lsd, tgt, mean, std, noise, res, steps, total_steps=50)
    try:
        P = Projector(G.cuda(), lsd, B, 50, latent_mean=mean, latent_std=std, forward_dtype="fp16",
                      step_noise=torch.cat([noise, torch.zeros(48, B, 17, 32)]))
        P.set_targets(tgt)
        P.run(steps)
        torch.cuda.synchronize()
    finally:
        _lib.set_forward_dtype(_lib.DEFAULT_FORWARD_DTYPE)
    np.testing.assert_allclose(P.losses[:steps].cpu().numpy(), ref["losses"].numpy(), rtol=3e-3)


def test_1000_step_projected_latents_exact_fp32_path_vs_oracle():
    """north_star: '1000-step projected latents within a stated tolerance'.  The exact-fp32 path (Projector(engine='ops'): ops-engine
    synthesis + fp32 LPIPS-VGG16 on the direct-convolution kernels, autograd) against the oracle loop, 1000 Adam steps at 32^2 with the
    same injected noise.  Stated tolerance: latents within 2e-2 max-abs (|z| ~ 1) and final losses within 1e-3 relative; the drift is
    fp32 summation-order noise amplified by 1000 Adam updates (measured value printed)."""
    from morphganformer_b200.projection import Projector, latent_stats
    res, B, steps = 32, 2, 1000
    G = util.build_G(res, 0, 1024, 32)
    gsd, lsd = util.state_dict_cpu(G), util.build_vgg_lpips_sd(4)
    mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
    noise = util.case_tensor((steps, B, 17, 32), 71)
    with torch.no_grad():
        tgt = torch.tanh(ganformer.generator(gsd, util.case_tensor((B, 17, 32), 72), res)[0])
    ref = oproj.project(gsd, lsd, tgt, mean, std, noise, res, steps)
    P = Projector(G.cuda(), lsd, B, steps, latent_mean=mean, latent_std=std, step_noise=noise, engine="ops")
    P.set_targets(tgt)
    P.run(steps)
    torch.cuda.synchronize()
    got_l, want_l = P.losses.cpu(), ref["losses"]
    drift = (P.latent.cpu() - ref["latent"]).abs().max().item()
    rel = ((got_l - want_l).abs() / want_l.abs())
    print("1000-step latent drift (max-abs) %.3g; loss rel err: step 0 %.2e, step 99 %.2e, step 999 %.2e; loss %.4f -> %.4f"
          % (drift, rel[0].max().item(), rel[99].max().item(), rel[999].max().item(), want_l[0].mean().item(), want_l[-1].mean().item()))
    assert rel[0].max().item() < 1e-5
    assert drift < 2e-2
    assert rel[999].max().item() < 1e-3
    assert want_l[-1].mean() < want_l[0].mean()


def test_projection_with_random_synthesis_noise_runs_and_is_captured_in_a_graph():
    """noise_mode='random' (the reference scripts' default synthesis noise): per-step fresh noise planes, eager and under CUDA-graph replay."""
    from morphganformer_b200.projection import Projector, latent_stats
    res, B = 64, 2
    lsd = util.build_vgg_lpips_sd(4)
    mean, std = latent_stats(util.case_tensor((2000, 17, 32), 70))
    tgt = torch.tanh(util.case_tensor((B, 3, res, res), 73))
    G = util.build_G(res, 0, 2048, 64).cuda()
    P = Projector(G, lsd, B, 20, latent_mean=mean, latent_std=std, step_noise=util.case_tensor((20, B, 17, 32), 71), noise_mode="random")
    P.set_targets(tgt)
    torch.manual_seed(5)
    P.step(); P.step()
    P.capture()
    for _ in range(4):
        P.step()
    torch.cuda.synchronize()
    l = P.losses[:6].cpu()
    assert torch.isfinite(l).all() and (l > 0).all()
    assert len({round(float(v), 6) for v in l[:, 0]}) == 6          # every step saw different noise (replays advance the RNG too)

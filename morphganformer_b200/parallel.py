"""Image-sharded multi-GPU plumbing for projection / synthesis jobs.

Every image is an independent job (its own latents, target, Adam state, activations; generator and VGG weights are read-only
replicas built from the same seed), so ranks exchange NOTHING per step.  The only collective of a run is one all_gather of the
projected latents [B_local,k,32] and their losses at the end (2.2 KB per image).  One process per GPU; backend nccl on GPUs
(NVLink 5 / NVSwitch), gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank`; the first n_items % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(latents, losses, group=None):
    """all_gather of per-rank results with possibly different shard sizes.  latents [b_r,k,d], losses [b_r] ->
    (latents [sum b_r,k,d], losses [sum b_r]) on every rank, in rank order (== original image order for shard_range)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return latents, losses
    world = dist.get_world_size(group)
    n = torch.tensor([latents.shape[0]], device=latents.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad_lat = torch.zeros((mx,) + tuple(latents.shape[1:]), device=latents.device, dtype=latents.dtype)
    pad_los = torch.zeros((mx,), device=losses.device, dtype=losses.dtype)
    pad_lat[:latents.shape[0]] = latents
    pad_los[:losses.shape[0]] = losses
    out_lat = [torch.empty_like(pad_lat) for _ in range(world)]
    out_los = [torch.empty_like(pad_los) for _ in range(world)]
    dist.all_gather(out_lat, pad_lat, group=group)
    dist.all_gather(out_los, pad_los, group=group)
    return (torch.cat([t[:s] for t, s in zip(out_lat, sizes)]), torch.cat([t[:s] for t, s in zip(out_los, sizes)]))


def project_sharded(make_projector, targets, steps):
    """Projects `targets` [N,3,R,R] (host tensor, same on every rank) with the images split over the ranks.
    make_projector(batch) -> Projector for `batch` local images.  Returns (latents [N,k,d], losses [N]) on every rank."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(targets.shape[0], rank, world)
    P = make_projector(hi - lo)
    P.set_targets(targets[lo:hi])
    out = P.run(steps)
    return gather_results(out["best_latent"], out["best_loss"])

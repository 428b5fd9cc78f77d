"""Ray-parallel rendering across the GPUs of one box.

Rays are independent units given a replicated scene: every rank renders one contiguous
shard of the ray list and ONE all-gather of rgb|depth assembles the image on every rank.  The
reference never shards an image (its only parallelism is Lightning DDP over batches,
configs/train_dtu.yaml:73-77), so this helper is new API next to `NeRFRendererDGS.forward`.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_rays, world, rank):
    """Contiguous equal shards; the last ranks may be one padded slot short."""
    per = (n_rays + world - 1) // world
    lo = min(n_rays, rank * per)
    return lo, min(n_rays, lo + per), per


def render_sharded(render_fn, rays, group=None):
    """rays (SB, NR, 8) replicated on every rank -> (rgb (SB,NR,3), depth (SB,NR)) on every rank.

    render_fn(rays_shard) -> (rgb (SB,n,3), depth (SB,n)) is this rank's renderer, e.g.
    ``lambda r: (lambda o: (o.fine.rgb, o.fine.depth))(renderer(model, r))``.
    """
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return render_fn(rays)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    SB, NR, _ = rays.shape
    lo, hi, per = shard_bounds(NR, world, rank)
    mine = torch.zeros(SB, per, 4, device=rays.device, dtype=torch.float32)
    if hi > lo:
        rgb, depth = render_fn(rays[:, lo:hi].contiguous())
        mine[:, :hi - lo, :3] = rgb
        mine[:, :hi - lo, 3] = depth
    out = torch.empty(world, SB, per, 4, device=rays.device, dtype=torch.float32)
    dist.all_gather_into_tensor(out.view(-1), mine.view(-1), group=group)
    full = out.permute(1, 0, 2, 3).reshape(SB, world * per, 4)[:, :NR]
    return full[..., :3].contiguous(), full[..., 3].contiguous()

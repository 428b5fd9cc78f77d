"""Ray-parallel rendering across the GPUs of one box.

Rays are independent units given a replicated scene: every rank renders one contiguous
shard of the ray list and ONE all-gather of rgb|depth assembles the image on every rank.  The
reference never shards an image (its only parallelism is Lightning DDP over batches,
configs/train_dtu.yaml:73-77), so this helper is new API next to `NeRFRendererDGS.forward`.

The compositing kernel writes its packed (r,g,b,depth) output straight into this rank's slice of the
gather buffer (diner_render_rgbd) and the collective runs in place, so between the last kernel of the
render and ncclAllGather there is no pack / copy pass; for one scene per call (SB = 1, the inference
case) the gathered buffer IS the image and the returned rgb / depth are views of it.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_rays, world, rank):
    """Contiguous equal shards; the last ranks may be one padded slot short."""
    per = (n_rays + world - 1) // world
    lo = min(n_rays, rank * per)
    return lo, min(n_rays, lo + per), per


def render_sharded(render_fn, rays, group=None, packed=False, return_packed=False):
    """rays (SB, NR, 8) replicated on every rank -> (rgb (SB,NR,3), depth (SB,NR)) on every rank.

    packed=False: render_fn(rays_shard) -> (rgb (SB,n,3), depth (SB,n)), e.g.
        ``lambda r: (lambda o: (o.fine.rgb, o.fine.depth))(renderer(model, r))``.
    packed=True:  render_fn(rays_shard, out, ray_offset) writes (SB,n,4) = [r,g,b,depth] into the contiguous tensor `out`
        (``lambda r, out, off: renderer.render_packed(model, r, out=out, ray_offset=off)``) -- the zero-copy path;
        ray_offset = index of the shard's first ray, so that counter-based sampler noise does not depend on the world size.
    return_packed=True (with packed=True): returns the (SB,NR,4) image [r,g,b,depth] itself instead of the two views.
    """
    single = not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1
    SB, NR, _ = rays.shape
    if single:
        if not packed:
            return render_fn(rays)
        full = torch.empty(SB, NR, 4, device=rays.device, dtype=torch.float32)
        render_fn(rays, full, 0)
        return full if return_packed else (full[..., :3], full[..., 3])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi, per = shard_bounds(NR, world, rank)
    n = hi - lo
    out = torch.empty(world, SB, per, 4, device=rays.device, dtype=torch.float32)
    mine = out[rank]                                        # (SB, per, 4): this rank's slice, gathered in place
    if n > 0:
        shard = rays[:, lo:hi] if SB == 1 else rays[:, lo:hi].contiguous()      # SB == 1: already contiguous
        if packed and (SB == 1 or n == per):
            render_fn(shard, mine[:, :n] if n < per else mine, lo)
        elif packed:
            tmp = torch.empty(SB, n, 4, device=rays.device, dtype=torch.float32)
            render_fn(shard, tmp, lo)
            mine[:, :n] = tmp
        else:
            rgb, depth = render_fn(shard.contiguous())
            mine[:, :n, :3] = rgb
            mine[:, :n, 3] = depth
    dist.all_gather_into_tensor(out.view(-1), mine.reshape(-1), group=group)
    full = (out.view(1, world * per, 4) if SB == 1 else out.permute(1, 0, 2, 3).reshape(SB, world * per, 4))[:, :NR]
    return full if return_packed else (full[..., :3], full[..., 3])

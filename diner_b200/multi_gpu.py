"""Ray-parallel rendering across the GPUs of one box.

Rays are independent units given a replicated scene: every rank renders one contiguous
shard of the ray list and ONE all-gather of rgb|depth assembles the image on every rank.  The
reference never shards an image (its only parallelism is Lightning DDP over batches,
configs/train_dtu.yaml:73-77), so this helper is new API next to `NeRFRendererDGS.forward`.

The compositing kernel writes its packed (r,g,b,depth) output straight into this rank's slice of the
gather buffer (diner_render_rgbd) and the collective runs in place, so between the last kernel of the
render and ncclAllGather there is no pack / copy pass; for one scene per call (SB = 1, the inference
case) and equal shards the gathered buffer IS the image and the returned rgb / depth are views of it.

Shards may be weighted (`weights`, one positive number per rank) for boxes whose GPUs differ in speed;
`balance_weights` turns per-rank render times into weights and shard boundaries stay multiples of
`align` rays (whole image rows keep the 2-D tile order of the fused launch).  On the 8 x B200 box of
this project the per-GPU differences of a step turned out to be run-to-run noise under the power cap,
not persistent speed differences, so `bench.py` keeps equal shards (DINER_BALANCE=1 enables it).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_rays, world, rank, weights=None, align=1):
    """(lo, hi, per): contiguous shard [lo, hi) of this rank and the slot size `per` of the gather buffer (largest shard).
    Equal shards by default (the last ranks may be one padded slot short); `weights` -> shard sizes proportional to them,
    boundaries rounded to multiples of `align`."""
    if weights is None:
        per = (n_rays + world - 1) // world
        lo = min(n_rays, rank * per)
        return lo, min(n_rays, lo + per), per
    w = [max(float(x), 1e-9) for x in weights]
    tot, acc, cuts = sum(w), 0.0, [0]
    for r in range(world - 1):
        acc += w[r]
        c = int(round(n_rays * acc / tot / align)) * align
        cuts.append(min(n_rays, max(cuts[-1], c)))
    cuts.append(n_rays)
    per = max(b - a for a, b in zip(cuts, cuts[1:]))
    return cuts[rank], cuts[rank + 1], per


def balance_weights(seconds, group=None):
    """Per-rank render time of an equal-shard step (float, this rank) -> list of weights (1 / time) gathered from all ranks."""
    world = dist.get_world_size(group)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [1.0 / max(float(x), 1e-9) for x in out]


def render_sharded(render_fn, rays, group=None, packed=False, return_packed=False, weights=None, align=1):
    """rays (SB, NR, 8) replicated on every rank -> (rgb (SB,NR,3), depth (SB,NR)) on every rank.

    packed=False: render_fn(rays_shard) -> (rgb (SB,n,3), depth (SB,n)), e.g.
        ``lambda r: (lambda o: (o.fine.rgb, o.fine.depth))(renderer(model, r))``.
    packed=True:  render_fn(rays_shard, out, ray_offset) writes (SB,n,4) = [r,g,b,depth] into the contiguous tensor `out`
        (``lambda r, out, off: renderer.render_packed(model, r, out=out, ray_offset=off)``) -- the zero-copy path;
        ray_offset = index of the shard's first ray, so that counter-based sampler noise does not depend on the world size.
    return_packed=True (with packed=True): returns the (SB,NR,4) image [r,g,b,depth] itself instead of the two views.
    weights / align: see shard_bounds (identical on every rank).
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
    lo, hi, per = shard_bounds(NR, world, rank, weights, align)
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
    if weights is None:
        full = (out.view(1, world * per, 4) if SB == 1 else out.permute(1, 0, 2, 3).reshape(SB, world * per, 4))[:, :NR]
    else:                                                   # uneven shards: compact the padded slots (one small copy)
        sizes = [b - a for a, b in (shard_bounds(NR, world, r, weights, align)[:2] for r in range(world))]
        full = torch.cat([out[r, :, :s] for r, s in enumerate(sizes) if s > 0], dim=1)
    return full if return_packed else (full[..., :3], full[..., 3])

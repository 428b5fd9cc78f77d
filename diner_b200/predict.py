"""Host-side mirror of the reference's render loop around the hot path (src/models/diner.py:64-97):
`encode_batch` + `predict_imgs_from_batch`, with ray generation, the ray_batch_size chunk loop and the torch.cat of the
chunks folded into one libdiner_b200 call (diner_render_image)."""
import torch


def encode_batch(nerf, batch):
    """diner.py:64-70."""
    nerf.encode(images=batch["src_rgbs"], depths=batch["src_depths"], depths_std=batch["src_depth_stds"],
                extrinsics=batch["src_extrinsics"], intrinsics=batch["src_intrinsics"])


@torch.no_grad()
def predict_imgs_from_batch(nerf, renderer, batch, znear, zfar, return_depth=False, encode=True):
    """diner.py:72-97: batch with target_rgb (SB,3,H,W) [shape only], target_extrinsics (SB,4,4), target_intrinsics (SB,3,3)
    (+ the src_* entries when encode=True) -> rgb (SB,3,H,W) [, depth (SB,1,H,W)]."""
    SB, _, H, W = batch["target_rgb"].shape
    if encode:
        encode_batch(nerf, batch)
    rgb, depth = renderer.render_image(nerf, batch["target_extrinsics"], batch["target_intrinsics"], H, W, znear, zfar)
    rgb = rgb.view(SB, H, W, 3).permute(0, 3, 1, 2)
    depth = depth.view(SB, H, W, 1).permute(0, 3, 1, 2)
    return (rgb, depth) if return_depth else rgb


def calc_losses(nerf, renderer, batch, znear, zfar, ray_batch_size, generator=None, encode=True):
    """Training-step loss of the reference (src/models/diner.py:217-290) for w_vgg = w_antibias = 0: random pixels of the target
    view(s) (`torch.randint`, :232), their rays through NeRFRendererDGS.forward, MSE against the target colours (:259-266).
    Gradients flow through diner_render_backward (the autograd Function in nerf_renderer.py); the perceptual /
    anti-bias terms of the reference operate on the rendered patch afterwards and stay in PyTorch."""
    SB, _, H, W = batch["target_rgb"].shape
    if encode:
        encode_batch(nerf, batch)
    dev = batch["target_rgb"].device
    rays = nerf.context().gen_rays(batch["target_extrinsics"].float().contiguous(), batch["target_intrinsics"].float().contiguous(),
                                   H, W, znear, zfar)                                      # (SB, H*W, 8)
    pix = torch.randint(0, H * W, (SB, ray_batch_size), generator=generator).to(dev)
    bidx = torch.arange(SB, device=dev).unsqueeze(-1).expand(-1, ray_batch_size)
    rays = rays[bidx, pix].contiguous()                                                     # (SB, B, 8)
    pred = renderer(nerf, rays).fine.rgb
    gt = batch["target_rgb"].view(SB, 3, -1).permute(0, 2, 1)[bidx, pix]                    # (SB, B, 3)
    loss = torch.nn.functional.mse_loss(pred, gt, reduction="mean")
    return dict(rgb_fine=loss, vgg_fine=0.0, antibias=0.0, total=loss)


@torch.no_grad()
def cam_sweep_frames(nerf, renderer, base_batch, target_extrinsics, znear, zfar, encode=True):
    """The render loop of DINER.create_cam_sweep (src/models/diner.py:180-206) for one base sample: the scene is encoded once,
    every sweep camera `target_extrinsics[i]` (N,4,4) is rendered with the base sample's target intrinsics, and the frames are
    returned like the reference assembles them: (2N-1, 3, 2H, W) = [rgb over colour-mapped depth], forward then backward
    (ping-pong, :209-211).  Per frame the reference generates rays, loops over ray batches and moves every chunk to the host
    (:183-199); here it is one diner_render_image call per frame and everything stays on the device."""
    from .io import torch_cmap
    _, _, H, W = base_batch["target_rgb"].shape
    if encode:
        encode_batch(nerf, base_batch)
    rgbs, depths = [], []
    for i in range(target_extrinsics.shape[0]):
        rgb, depth = renderer.render_image(nerf, target_extrinsics[i:i + 1], base_batch["target_intrinsics"], H, W, znear, zfar)
        rgbs.append(rgb.view(H, W, 3).permute(2, 0, 1))
        depths.append(torch_cmap(depth.view(1, H, W)).to(rgb.dtype))
    frames = torch.cat((torch.stack(rgbs), torch.stack(depths)), dim=-2)
    n = frames.shape[0]
    idcs = torch.cat((torch.arange(n), torch.arange(n - 1, 0, -1)))
    return frames[idcs.to(frames.device)]


def save_sweep(frames, outpath, fps=5):
    """torch_helpers.save_torch_video (src/util/torch_helpers.py:78-96) when imageio is installed (.mp4), else numbered PNG frames
    next to `outpath` written by the asynchronous ImageWriter."""
    import os
    try:
        import imageio
        imageio.mimwrite(outpath, (frames.permute(0, 2, 3, 1).detach().cpu().numpy() * 255).astype("uint8"), fps=fps, quality=10)
        return [outpath]
    except ImportError:
        from .io import ImageWriter
        w = ImageWriter()
        stem = os.path.splitext(outpath)[0]
        for i, f in enumerate(frames):
            w.add("%s-%04d.png" % (stem, i), f)
        return w.close()

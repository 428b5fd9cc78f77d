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

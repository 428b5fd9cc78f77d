"""Shared builders for the tests: product model on a device from the synthetic case inputs."""
import torch

from diner_b200.nerf_renderer import NeRFRendererDGS
from diner_b200.pixelnerf import PixelNeRF
from diner_b200.scene_ops import depth2normal


def product_model(batch, latent, mlp, device, mode="fp32"):
    model = PixelNeRF(
        poscode_conf=dict(kwargs=dict(num_freqs=6, freq_factor=6.28, include_input=True)),
        encoder_conf=dict(module="src.models.image_encoder.SpatialEncoder",
                          kwargs=dict(image_padding=64, padding_pe=4, pretrained=False)),
        mlp_fine_conf=dict(module="src.models.resnetfc.ResnetFC",
                           kwargs=dict(n_blocks=5, d_hidden=512, combine_layer=3, combine_type="average")))
    model.mlp_fine.load_state_dict(mlp)
    model = model.to(device).eval()
    SB, NV = batch["src_depths"].shape[:2]
    H, W = batch["src_depths"].shape[-2:]
    K = batch["src_intrinsics"].to(device)
    dep = batch["src_depths"].to(device)
    nrm = depth2normal(dep.flatten(end_dim=1), K.flatten(end_dim=1)).reshape(SB, NV, 3, H, W)
    model.encoder.set_scene(latent.to(device), dep, batch["src_depth_stds"].to(device), nrm)
    model.set_cameras(batch["src_extrinsics"].to(device), K, W, H)
    model.mode = mode
    return model


def renderer_for(cfg, noise=None, device="cuda"):
    r = NeRFRendererDGS(n_samples=cfg["K"], n_depth_candidates=cfg["C"], n_gaussian=cfg["G"],
                        white_bkgd=cfg["white"])
    if noise is not None:
        r.noise = {k: v.to(device).contiguous() for k, v in noise.items()}
    return r

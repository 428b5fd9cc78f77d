"""Shared builders for the tests.  The product-side builders live in the package (diner_b200.synthetic); the oracle-side
helper here moves an oracle Scene to a device so that the reference algorithm can be run as eager PyTorch on cuda."""
import copy

import torch

from diner_b200.synthetic import product_model, renderer_for  # noqa: F401  (re-exported for the tests)


def oracle_scene_on(scene, device):
    """Copy of an oracle Scene (built on the host: depth2normal uses host index tensors) with every tensor on `device`."""
    s = copy.copy(scene)
    for f in ("poses", "focal", "c", "image_shape", "latent", "depths", "depths_std", "normals"):
        setattr(s, f, getattr(scene, f).to(device))
    s.mlp = {k: v.to(device) for k, v in scene.mlp.items()}
    return s


def ray_err(rgb, depth, rgb_ref, depth_ref):
    """Per-ray max |err| over rgb and depth; NaN / inf anywhere counts as +inf (never as 'good')."""
    e = torch.maximum((rgb - rgb_ref).abs().max(dim=-1).values, (depth - depth_ref).abs())
    return torch.where(torch.isfinite(e), e, torch.full_like(e, float("inf")))

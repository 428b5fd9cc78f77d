"""Deterministic synthetic scenes, weights and noise for tests / bench (no datasets, no network).

Everything is generated with an integer hash (splitmix64) + additions only, so the very same
bits come out on every machine and torch/numpy version.  The batch dict mirrors the keys the
reference datasets emit (reference src/data/dtu.py:225-239, consumed at src/models/diner.py:66-79).
"""
import math

import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def hash_uniform(shape, seed, stream=0):
    """U[0,1) float32 with 24 random bits, keyed by (seed, stream, flat index)."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        key = _splitmix64(np.uint64(seed) * np.uint64(0x1000003) + np.uint64(stream))
        bits = _splitmix64(idx ^ key)
    u = (bits >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))
    return torch.from_numpy(u.reshape(shape))


def hash_normal(shape, seed, stream=0):
    """Approximately N(0,1) float32 (Irwin-Hall of 12 uniforms; additions only => bit-stable)."""
    acc = torch.zeros(shape, dtype=torch.float32)
    for j in range(12):
        acc = acc + hash_uniform(shape, seed, stream * 16 + j + 1000)
    return acc - 6.0


# ----------------------------------------------------------------------------------------------
# cameras / geometry
# ----------------------------------------------------------------------------------------------
def _look_at_origin(center):
    """OpenCV world->cam (x right, y down, z forward) for a camera at `center` looking at 0."""
    c = np.asarray(center, dtype=np.float64)
    f = -c / np.linalg.norm(c)
    x = np.cross(np.array([0.0, 1.0, 0.0]), f)
    x /= np.linalg.norm(x)
    y = np.cross(f, x)
    R = np.stack([x, y, f], 0)
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ c
    return E


def _intrinsics(H, W):
    K = np.eye(3)
    K[0, 0] = 1.25 * W
    K[1, 1] = 1.25 * W
    K[0, 2] = W / 2.0
    K[1, 2] = H / 2.0
    return K


def _render_view(E, K, H, W, radius, plane_z, plane_r):
    """Analytic depth (cam-space z), fg mask and world hit points for sphere(0,radius) + a disc."""
    R, t = E[:3, :3], E[:3, 3]
    C = -R.T @ t
    v, u = np.meshgrid(np.arange(H) + 0.5, np.arange(W) + 0.5, indexing="ij")
    d_cam = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], -1)
    d_w = d_cam @ R  # = R^T d (rows)
    # sphere
    a = (d_w * d_w).sum(-1)
    b = 2.0 * (d_w @ C)
    c = C @ C - radius * radius
    disc = b * b - 4 * a * c
    ts = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
    ts = np.where(ts > 0, ts, np.inf)
    # disc plane: world z = plane_z, |xy| < plane_r
    tp = (plane_z - C[2]) / np.where(np.abs(d_w[..., 2]) > 1e-9, d_w[..., 2], 1e-9)
    hit = C[None, None, :] + tp[..., None] * d_w
    okp = (tp > 0) & ((hit[..., 0] ** 2 + hit[..., 1] ** 2) < plane_r ** 2)
    tp = np.where(okp, tp, np.inf)
    tt = np.minimum(ts, tp)
    fg = np.isfinite(tt)
    depth = np.where(fg, tt, 0.0)  # d_cam.z == 1 -> parameter == cam-space z
    pts = C[None, None, :] + np.where(fg, tt, 0.0)[..., None] * d_w
    return depth, fg, pts


def make_scene(H=64, W=64, NV=4, SB=1, near=1.0, far=2.5, seed=0):
    """Synthetic multi-view batch with the reference's batch-dict keys."""
    D = 0.5 * (near + far)
    radius = 0.22 * D
    plane_z, plane_r = 0.12 * D, 0.30 * D  # disc slightly behind the sphere centre (camera at -z)
    angs = np.linspace(-0.3, 0.3, NV) if NV > 1 else np.array([0.1])
    out = {k: [] for k in ("src_rgbs", "src_depths", "src_depth_stds", "src_extrinsics",
                           "src_intrinsics", "target_extrinsics", "target_intrinsics",
                           "target_rgb", "target_alpha")}
    K = _intrinsics(H, W)
    for sb in range(SB):
        rgbs, deps, stds, exts, ints = [], [], [], [], []
        for vi, a in enumerate(angs):
            a = a + 0.03 * sb
            elev = 0.05 * ((vi % 2) * 2 - 1)
            Cc = np.array([D * math.sin(a), D * elev, -D * math.cos(a)])
            E = _look_at_origin(Cc)
            depth, fg, pts = _render_view(E, K, H, W, radius, plane_z, plane_r)
            tex = 0.5 + 0.5 * np.sin(pts * (14.0 / D) + np.array([0.0, 2.0, 4.0]))
            rgb = np.where(fg[..., None], tex, 0.0)
            nz = hash_uniform((H, W), seed, 10 + sb * 64 + vi).numpy().astype(np.float64)
            std = np.where(fg, (0.005 + 0.015 * nz) * D / 1.75, 0.0)
            rgbs.append(rgb.transpose(2, 0, 1))
            deps.append(depth[None])
            stds.append(std[None])
            exts.append(E)
            ints.append(K)
        Ct = np.array([D * math.sin(0.02 + 0.03 * sb), -0.08 * D, -D * math.cos(0.02 + 0.03 * sb)])
        Et = _look_at_origin(Ct)
        dt, fgt, ptst = _render_view(Et, K, H, W, radius, plane_z, plane_r)
        text = 0.5 + 0.5 * np.sin(ptst * (14.0 / D) + np.array([0.0, 2.0, 4.0]))
        out["src_rgbs"].append(np.stack(rgbs))
        out["src_depths"].append(np.stack(deps))
        out["src_depth_stds"].append(np.stack(stds))
        out["src_extrinsics"].append(np.stack(exts))
        out["src_intrinsics"].append(np.stack(ints))
        out["target_extrinsics"].append(Et)
        out["target_intrinsics"].append(K)
        out["target_rgb"].append(np.where(fgt[..., None], text, 0.0).transpose(2, 0, 1))
        out["target_alpha"].append(fgt[None].astype(np.float64))
    batch = {k: torch.from_numpy(np.stack(v)).float().contiguous() for k, v in out.items()}
    batch["znear"] = float(near)
    batch["zfar"] = float(far)
    return batch


def make_latent(SB, NV, C, Hl, Wl, seed=0, scale=0.5):
    """Random feature maps standing in for the ResNet pyramid (reference image_encoder.py:290-291)."""
    return (hash_normal((SB, NV, C, Hl, Wl), seed, 77) * scale).contiguous()


def make_mlp_state(d_in=55, d_latent=512, d_hidden=512, d_out=4, n_blocks=5, combine_layer=3, seed=0,
                   sigma_gain=6.0, sigma_bias=2.0):
    """Non-degenerate ResnetFC weights keyed like the reference state_dict (resnetfc.py:92-118).

    The reference's default init zeroes fc_1 (resnetfc.py:47) which makes sigma identically 0 with
    random features; these weights keep alpha / rgb / depth mid-range instead.
    """
    sd = {}
    st = [0]

    def lin(name, fo, fi, gain, bias_scale=0.02):
        st[0] += 1
        sd[name + ".weight"] = hash_normal((fo, fi), seed, 200 + st[0]) * (gain * math.sqrt(2.0 / fi))
        sd[name + ".bias"] = hash_normal((fo,), seed, 400 + st[0]) * bias_scale

    lin("lin_in", d_hidden, d_in, 1.0)
    lin("lin_out", d_out, d_hidden, 0.35)
    for b in range(n_blocks):
        lin("blocks.%d.fc_0" % b, d_hidden, d_hidden, 1.0)
        lin("blocks.%d.fc_1" % b, d_hidden, d_hidden, 0.45)
    for b in range(min(combine_layer, n_blocks)):
        lin("lin_z.%d" % b, d_hidden, d_latent, 0.6)
    sd["lin_out.weight"][3] *= sigma_gain
    sd["lin_out.bias"][3] = sigma_bias
    return sd


def gen_rays(extrinsics, intrinsics, W, H, z_near, z_far):
    """(B,H,W,8) rays [origin3, dir3, near, far]; pixel centres at +0.5.

    Restates reference src/util/cam_geometry.py:5-48 (input producer of the hot path).
    """
    B = extrinsics.shape[0]
    dev = extrinsics.device
    fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
    cx, cy = intrinsics[:, 0, 2], intrinsics[:, 1, 2]
    ys, xs = torch.meshgrid(torch.arange(0.5, H, 1, device=dev), torch.arange(0.5, W, 1, device=dev),
                            indexing="ij")
    px = (xs[None] - cx.view(B, 1, 1)) / fx.view(B, 1, 1)
    py = (ys[None] - cy.view(B, 1, 1)) / fy.view(B, 1, 1)
    d_cam = torch.stack((px, py, torch.ones_like(px)), -1)
    d_cam = d_cam / d_cam.pow(2).sum(-1, keepdim=True).sqrt()
    Rt = extrinsics[:, :3, :3].permute(0, 2, 1)
    d_w = (Rt @ d_cam.view(B, -1, 3).permute(0, 2, 1)).permute(0, 2, 1).view(B, H, W, 3)
    org = (-1 * Rt @ extrinsics[:, :3, -1:]).view(B, 1, 1, 3).expand(-1, H, W, -1)
    nr = z_near.view(B, 1, 1, 1).expand(-1, H, W, -1)
    fr = z_far.view(B, 1, 1, 1).expand(-1, H, W, -1)
    return torch.cat((org, d_w, nr, fr), -1)


# ----------------------------------------------------------------------------------------------
# product modules on a device from synthetic case inputs (bench.py, __graft_entry__.smoke(), tests)
# ----------------------------------------------------------------------------------------------
def product_model(batch, latent, mlp, device, mode="fp32"):
    """PixelNeRF (reference constructor arguments of configs/train_dtu.yaml:32-50) with the given ResnetFC weights, the
    synthetic scene installed as PixelNeRF.encode would leave it (pixelnerf.py:35-53) minus the ResNet trunk: the latent
    maps are supplied.  `mode` = 'fp32' | 'parity' | 'fast' (diner_b200.pixelnerf)."""
    from .pixelnerf import PixelNeRF
    from .scene_ops import depth2normal
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
    """NeRFRendererDGS for a case dict (K, C, G, white), optionally with dense injected noise."""
    from .nerf_renderer import NeRFRendererDGS
    r = NeRFRendererDGS(n_samples=cfg["K"], n_depth_candidates=cfg["C"], n_gaussian=cfg["G"],
                        white_bkgd=cfg["white"])
    if noise is not None:
        r.noise = {k: (v.to(device).contiguous() if torch.is_tensor(v) else v) for k, v in noise.items()}
    return r

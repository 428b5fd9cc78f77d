"""CPU oracle for the DINER volumetric-rendering hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file; the product (diner_b200/, src/) never does and fails loudly without its CUDA library.

It is a plain torch-CPU fp32 restatement of the reference algorithm, written as free functions over
an explicit `Scene`, with the three random draws of the reference turned into *dense, injectable*
noise arrays so that a fused CUDA kernel can be compared on identical inputs:

    u_coarse (SB,NR,C)  <- torch.rand_like  at reference src/models/nerf_renderer.py:57
    g_noise  (SB,NR,G)  <- torch.randn_like at reference src/models/nerf_renderer.py:188 (masked rays only)
    u_fill   (SB,NR,K)  <- torch.rand_like  at reference src/models/nerf_renderer.py:390 (missing slots only;
                           indexed by the slot's column AFTER the ascending sort)

Pinning: the reference ships no tests / golden vectors, so the oracle is pinned against
the reference's own modules executed in the build container: oracle/make_golden.py runs the
unmodified reference (imported from /root/reference through oracle/ref_import.py) with the same
dense noise injected, and stores its outputs under tests/golden/; tests/test_oracle.py checks this
file against those fixtures bit-for-bit (and live against the reference when the tree is present).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# scene state  (what reference PixelNeRF.encode leaves on the modules: pixelnerf.py:35-53,
#               image_encoder.py:225-291)
# ----------------------------------------------------------------------------------------------
@dataclass
class Scene:
    poses: torch.Tensor          # (SB,NV,4,4) world->cam
    focal: torch.Tensor          # (SB,NV,2)
    c: torch.Tensor              # (SB,NV,2)
    image_shape: torch.Tensor    # (2,) = [W,H]
    latent: torch.Tensor         # (SB,NV,L,Hl,Wl)  NCHW like the reference
    depths: torch.Tensor         # (SB,NV,1,H,W)
    depths_std: torch.Tensor     # (SB,NV,1,H,W)
    normals: torch.Tensor        # (SB,NV,3,H,W)
    mlp: Dict[str, torch.Tensor]  # ResnetFC state dict (lin_in, lin_out, blocks.N.fc_{0,1}, lin_z.N)
    feature_padding: float = 32.0
    num_freqs: int = 6
    freq_factor: float = 6.28
    n_blocks: int = 5
    combine_layer: int = 3
    beta: float = 0.0            # ResnetFC(beta): > 0 -> Softplus(beta) activations (resnetfc.py:124-127)
    extra: dict = field(default_factory=dict)


# ----------------------------------------------------------------------------------------------
# leaves
# ----------------------------------------------------------------------------------------------
def positional_encoding(x, num_freqs=6, freq_factor=6.28):
    """[x, sin(phase_j + x*f_j)], j over (f0,f0,f1,f1,...) with phases (0,pi/2,...).

    Follows reference src/models/positional_encoding.py:14-53 (include_input=True).
    """
    shp = x.shape
    d = shp[-1]
    x2 = x.reshape(-1, d)
    freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)                      # :18
    f = torch.repeat_interleave(freqs, 2).view(1, -1, 1).to(x2)                   # :24-26
    ph = torch.zeros(2 * num_freqs)
    ph[1::2] = np.pi * 0.5                                                        # :29-30
    ph = ph.view(1, -1, 1).to(x2)
    e = x2.unsqueeze(1).repeat(1, 2 * num_freqs, 1)                               # :45
    e = torch.sin(torch.addcmul(ph, e, f)).view(x2.shape[0], -1)                  # :46-47
    e = torch.cat((x2, e), dim=-1)                                                # :49
    return e.reshape(*shp[:-1], e.shape[-1])


def project_uv(scene: Scene, xyz_cam):
    """cam-space points (SB,NV,B,3) -> normalised uv in [-1,1] (pixelnerf.py:105-108, nerf_renderer.py:107-110)."""
    uv = xyz_cam[..., :2] / xyz_cam[..., 2:]
    uv = uv * scene.focal.unsqueeze(-2)
    uv = uv + scene.c.unsqueeze(-2)
    return uv / scene.image_shape * 2 - 1


def world_to_cam(scene: Scene, xyz):
    """(SB,B,3) -> (SB,NV,B,3): R x + t per view (pixelnerf.py:91-93, nerf_renderer.py:99-101)."""
    NV = scene.poses.shape[1]
    x = xyz.unsqueeze(1).expand(-1, NV, -1, -1)
    rot = torch.matmul(scene.poses[:, :, :3, :3], x.transpose(-2, -1)).transpose(-2, -1)
    return rot + scene.poses[:, :, :3, -1].unsqueeze(-2)


def _flat_grid(t, uv):
    SB, NV, N, _ = uv.shape
    return t.reshape(SB * NV, *t.shape[-3:]), uv.reshape(SB * NV, N, 1, 2)


def index_latent(scene: Scene, uv):
    """Bilinear / border gather of the latent with the feature-padding rescale (image_encoder.py:97-146)."""
    SB, NV, N, _ = uv.shape
    Hl, Wl = scene.latent.shape[-2:]
    size = torch.tensor([Wl, Hl], device=uv.device)
    uv = uv * ((size - scene.feature_padding * 2) / size).view(1, 1, 1, 2)       # :113-114
    lat, g = _flat_grid(scene.latent, uv)
    s = F.grid_sample(lat, g, align_corners=False, mode="bilinear", padding_mode="border")
    return s[:, :, :, 0].view(SB, NV, -1, N)


def index_depth(scene: Scene, uv):
    """Nearest / border (image_encoder.py:148-170)."""
    SB, NV, N, _ = uv.shape
    d, g = _flat_grid(scene.depths, uv)
    s = F.grid_sample(d, g, align_corners=False, mode="nearest", padding_mode="border")
    return s[:, :, :, 0].view(SB, NV, 1, N)


def exponential_padding(img, padding, double_width):
    """Replicate-pad then scale ring r (Chebyshev distance-1 from the image) by 2^(r/double_width).

    Follows reference src/util/torch_helpers.py:99-121 (including the overwrite order of its loop:
    rows first then columns, i ascending, so a corner cell ends with max(ring_y, ring_x)).
    """
    N, C, H, W = img.shape
    base = F.pad(img, [padding] * 4, mode="replicate")
    ex = torch.zeros(N, C, H + 2 * padding, W + 2 * padding, dtype=img.dtype, device=img.device)
    for i in range(padding):
        idx = padding - (i + 1)
        ex[:, :, idx, :] = i
        ex[:, :, -(idx + 1), :] = i
        ex[:, :, :, idx] = i
        ex[:, :, :, -(idx + 1)] = i
    return base * torch.exp(ex / double_width * np.log(2))


def index_depth_std(scene: Scene, uv, pad_size=100, pad_double_width=12):
    """Nearest lookup in the exponentially padded std map, zeros outside
    (image_encoder.py:172-199 -> torch_helpers.py:124-159)."""
    SB, NV, N, _ = uv.shape
    d, g = _flat_grid(scene.depths_std, uv)
    H, W = d.shape[-2:]
    img_size = torch.tensor([W, H], dtype=torch.float, device=uv.device)
    padded = exponential_padding(d, pad_size, pad_double_width)
    g = g * (img_size / (img_size + 2 * pad_size)).view(1, 1, 1, 2)              # :157-158
    s = F.grid_sample(padded, g, mode="nearest", padding_mode="zeros", align_corners=False)
    return s[:, :, :, 0].view(SB, NV, 1, N)


def index_normal(scene: Scene, uv):
    """Nearest / zeros (image_encoder.py:201-223)."""
    SB, NV, N, _ = uv.shape
    d, g = _flat_grid(scene.normals, uv)
    s = F.grid_sample(d, g, align_corners=False, mode="nearest", padding_mode="zeros")
    return s[:, :, :, 0].view(SB, NV, 3, N)


def resnetfc(scene: Scene, zx):
    """zx (SB,NV,B,L+d_in) -> (SB,B,4).  Follows reference src/models/resnetfc.py:61-69,129-159
    (ReLU activations, mean over the view axis at `combine_layer`)."""
    m = scene.mlp
    act = (lambda t: F.softplus(t, beta=scene.beta)) if scene.beta > 0 else torch.relu      # resnetfc.py:124-127
    L = m["lin_z.0.weight"].shape[1]
    z, x = zx[..., :L], zx[..., L:]
    x = F.linear(x, m["lin_in.weight"], m["lin_in.bias"])
    for b in range(scene.n_blocks):
        if b == scene.combine_layer:
            x = torch.mean(x, dim=1)
        if b < scene.combine_layer:
            x = x + F.linear(z, m["lin_z.%d.weight" % b], m["lin_z.%d.bias" % b])
        net = F.linear(act(x), m["blocks.%d.fc_0.weight" % b], m["blocks.%d.fc_0.bias" % b])
        dx = F.linear(act(net), m["blocks.%d.fc_1.weight" % b], m["blocks.%d.fc_1.bias" % b])
        x = x + dx
    return F.linear(act(x), m["lin_out.weight"], m["lin_out.bias"])


def query(scene: Scene, xyz, viewdirs):
    """(SB,B,3) world points + dirs -> (SB,B,4) [sigmoid rgb, relu sigma] (pixelnerf.py:55-145)."""
    SB, B, _ = xyz.shape
    NV = scene.poses.shape[1]
    xc = world_to_cam(scene, xyz)
    zf = positional_encoding(xc, scene.num_freqs, scene.freq_factor)              # :96
    vd = viewdirs.unsqueeze(1).expand(-1, NV, -1, -1)
    vd = torch.matmul(scene.poses[:, :, :3, :3], vd.transpose(-1, -2)).transpose(-1, -2)
    zf = torch.cat((zf, vd), dim=-1)                                              # :102
    uv = project_uv(scene, xc)
    lat = index_latent(scene, uv).transpose(-1, -2)                               # :110-111
    dd = index_depth(scene, uv).squeeze(-2) - xc[..., -1]                         # :114-115
    df = positional_encoding(dd.unsqueeze(-1), scene.num_freqs, scene.freq_factor)
    out = resnetfc(scene, torch.cat((lat, zf, df), dim=-1)).reshape(SB, B, 4)     # :128-137
    return torch.cat((torch.sigmoid(out[..., :3]), torch.relu(out[..., 3:4])), dim=-1)


# ----------------------------------------------------------------------------------------------
# renderer stages
# ----------------------------------------------------------------------------------------------
def sample_coarse(rays, C, u_coarse):
    """near*(1-s)+far*s, s = linspace(0,1-1/C,C) + U/C  (nerf_renderer.py:39-63)."""
    shp = rays.shape
    r = rays.reshape(-1, 8)
    near, far = r[:, -2:-1], r[:, -1:]
    step = 1.0 / C
    s = torch.linspace(0, 1 - step, C, device=rays.device).unsqueeze(0).repeat(r.shape[0], 1)
    s = s + u_coarse.reshape(-1, C) * step
    return (near * (1 - s) + far * s).view(*shp[:-1], C)


def candidate_likelihood(scene: Scene, rays, z_cand, depth_diff_max=0.05):
    """Per-candidate surface likelihood, max over views (nerf_renderer.py:95-129) -> (SB,NR,C)."""
    SB, NR, C = z_cand.shape
    NV = scene.poses.shape[1]
    step = (rays[..., -1] - rays[..., -2]) / C                                    # :95
    xyz = rays[..., None, :3] + z_cand.unsqueeze(-1) * rays[..., None, 3:6]
    xc = world_to_cam(scene, xyz.reshape(SB, -1, 3))
    rd = rays[..., 3:6].unsqueeze(1).expand(SB, NV, NR, 3)
    rd = (scene.poses[:, :, :3, :3] @ rd.transpose(-2, -1)).transpose(-2, -1)     # :103
    pd = rd.repeat_interleave(C, dim=-2)
    uv = project_uv(scene, xc)
    rdep, rstd, rnrm = index_depth(scene, uv), index_depth_std(scene, uv), index_normal(scene, uv)
    rz = xc[..., 2:].permute(0, 1, 3, 2)
    st = step.repeat_interleave(C, dim=1).view(SB, 1, 1, NR * C).expand_as(rdep)
    cosd = (pd.transpose(-2, -1) * rnrm).sum(dim=-2, keepdim=True)
    mask = (rstd != 0) & ((rdep - rz).abs() < depth_diff_max) & (cosd <= 0)       # :120-123
    lik = torch.zeros_like(rdep)
    lik[mask] = 0.5 * (
        torch.special.erf((rz[mask] + st[mask] / 2 - rdep[mask]) / (rstd[mask] * np.sqrt(2))) -
        torch.special.erf((rz[mask] - st[mask] / 2 - rdep[mask]) / (rstd[mask] * np.sqrt(2)))
    ).abs()                                                                        # :125-128
    return torch.max(lik, dim=1).values.squeeze(1).reshape(SB, NR, C)              # :129-130


def sample_depthguided(scene: Scene, rays, K, C, G, u_coarse, g_noise, return_aux=False, depth_diff_max=0.05):
    """Depth-guided shortlist (nerf_renderer.py:65-190) with dense injected noise -> (SB,NR,K), 0 = empty."""
    assert K >= G
    SB, NR = rays.shape[:2]
    z_cand = sample_coarse(rays, C, u_coarse)
    lik = candidate_likelihood(scene, rays, z_cand, depth_diff_max)
    opq = lik.clone()
    opq[..., 1:] *= torch.cumprod(1. - lik, dim=-1)[..., :-1]                      # :131-132
    idx = lik.argsort(dim=-1, descending=True)[..., :K]                            # :172
    sel_l = torch.gather(lik, -1, idx)
    z = torch.gather(z_cand, -1, idx)
    z[sel_l == 0.] = 0                                                             # :176-178
    ray_mask = torch.any(opq != 0, dim=-1)
    if G > 0:
        w = opq[ray_mask]
        x = z_cand[ray_mask]
        wn = w / w.sum(dim=-1, keepdims=True)                                      # torch_helpers.py:216
        mean = (x * wn).sum(dim=-1, keepdims=True)
        std = ((x - mean).pow(2) * wn).sum(dim=-1, keepdims=True).sqrt()
        gs = torch.zeros(SB, NR, G, device=rays.device)
        gs[ray_mask] = g_noise[ray_mask] * std + mean                              # :188
        z[..., -G:] = gs                                                           # :190
    if return_aux:
        return z, dict(z_cand=z_cand, lik=lik, opaque=opq, ray_mask=ray_mask)
    return z


def fill_up_uniform(z, rays, u_fill, return_mask=False):
    """Stratified fill of the zero slots, then ascending sort (nerf_renderer.py:367-397)."""
    shp = rays.shape
    K = z.shape[-1]
    z = z.sort(dim=-1).values.reshape(-1, K).clone()
    r = rays.reshape(-1, 8)
    miss = z == 0
    iray, isamp = torch.where(miss)
    nmiss = miss.int().sum(dim=-1)[iray]
    near, far = r[iray, -2], r[iray, -1]
    step = (far - near) / nmiss
    zm = near + isamp * step
    zm = zm + u_fill.reshape(-1, K)[iray, isamp] * step
    z[iray, isamp] = zm
    z = z.view(*shp[:2], K).sort(dim=-1).values
    if return_mask:
        return z, miss.view(*shp[:2], K)
    return z


def composite(scene: Scene, rays, z, white_bkgd, eval_batch_size=100000):
    """Alpha compositing of the queried samples (nerf_renderer.py:286-365) -> weights, rgb, depth."""
    SB, B, K = z.shape
    deltas = torch.cat([z[..., 1:] - z[..., :-1], rays[..., -1:] - z[..., -1:]], -1)   # :299-301
    pts = (rays[..., None, :3] + z.unsqueeze(-1) * rays[..., None, 3:6]).reshape(SB, B * K, 3)
    vd = rays[..., None, 3:6].expand(-1, -1, K, -1).reshape(SB, B * K, 3)
    ebs = (eval_batch_size - 1) // SB + 1
    out = torch.cat([query(scene, p, d) for p, d in
                     zip(torch.split(pts, ebs, dim=1), torch.split(vd, ebs, dim=1))], dim=1)
    out = out.reshape(SB, B, K, 4)
    rgbs, sig = out[..., :3], out[..., 3]
    alphas = 1 - torch.exp(-deltas * torch.relu(sig))                              # :344
    shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-10], -1)
    T = torch.cumprod(shifted, -1)
    w = alphas * T[..., :-1]
    rgb = torch.sum(w.unsqueeze(-1) * rgbs, -2)
    depth = torch.sum(w * z, -1)
    if white_bkgd:
        rgb = rgb + 1 - w.sum(dim=-1).unsqueeze(-1)                                # :357-360
    return w, rgb, depth


def render(scene: Scene, rays, K, C, G, white_bkgd, u_coarse, g_noise, u_fill, return_z=False):
    """NeRFRendererDGS.forward (nerf_renderer.py:399-424) with injected noise."""
    z = sample_depthguided(scene, rays, K, C, G, u_coarse, g_noise)
    z = fill_up_uniform(z, rays, u_fill)
    w, rgb, depth = composite(scene, rays, z, white_bkgd)
    if return_z:
        return rgb, depth, w, z
    return rgb, depth, w


# ----------------------------------------------------------------------------------------------
# scene preparation (restated only as far as the hot path needs its inputs)
# ----------------------------------------------------------------------------------------------
def depth2normal(dmap, Kmat):
    """Normals from depth by central differences with hole clean-up.
    Restates reference src/util/depth2normal.py:6-87 (input producer for index_normal)."""
    N, _, H, W = dmap.shape
    ys, xs = torch.meshgrid(torch.arange(0.5, H, 1.), torch.arange(0.5, W, 1.), indexing="ij")
    r = torch.stack((xs, ys), -1).reshape(-1, 2).unsqueeze(0).expand(N, -1, -1).clone()
    r -= Kmat[:, [0, 1], -1].unsqueeze(-2)
    r /= Kmat[:, [0, 1], [0, 1]].unsqueeze(-2)
    r = torch.cat((r, torch.ones_like(r[..., -1:])), dim=-1)
    p = (r.view(N, H, W, 3) * dmap.view(N, H, W, 1)).permute(0, 3, 1, 2)
    p = F.pad(p, [1] * 4, mode="replicate")
    dn, up = p[:, :, 2:, 1:-1], p[:, :, :-2, 1:-1]
    rt, lf = p[:, :, 1:-1, 2:], p[:, :, 1:-1, :-2]
    nrm = torch.linalg.cross((dn - up).permute(0, 2, 3, 1), (rt - lf).permute(0, 2, 3, 1), dim=-1)
    nrm = nrm / torch.norm(nrm, p=2, dim=-1, keepdim=True)
    off = torch.zeros(N, H, W, 3, dtype=torch.long)
    off[..., 1] += -(dn[:, 0] == 0).long() + (up[:, 0] == 0).long()
    off[..., 2] += -(rt[:, 0] == 0).long() + (lf[:, 0] == 0).long()
    idx = torch.stack(torch.meshgrid(torch.arange(N), torch.arange(H), torch.arange(W), indexing="ij"), -1)
    m = torch.any(off != 0, dim=-1)
    ni = idx[m] + off[m]
    ni[:, 1] = ni[:, 1].clip(min=0, max=H - 1)
    ni[:, 2] = ni[:, 2].clip(min=0, max=W - 1)
    nrm[m] = nrm[ni[:, 0], ni[:, 1], ni[:, 2]]
    nrm[dmap[:, 0] == 0] = 0
    return nrm.permute(0, 3, 1, 2)


def make_scene_state(batch, latent, mlp, feature_padding=32.0, **kw) -> Scene:
    """Batch dict (+ externally supplied latent maps) -> Scene, as PixelNeRF.encode would leave it
    (pixelnerf.py:44-51) minus the ResNet trunk, which is out of the hot path."""
    SB, NV = batch["src_depths"].shape[:2]
    H, W = batch["src_depths"].shape[-2:]
    K = batch["src_intrinsics"]
    nrm = depth2normal(batch["src_depths"].flatten(end_dim=1), K.flatten(end_dim=1)).reshape(SB, NV, 3, H, W)
    return Scene(poses=batch["src_extrinsics"], focal=K[:, :, torch.tensor([0, 1]), torch.tensor([0, 1])],
                 c=K[:, :, :2, -1], image_shape=torch.tensor([float(W), float(H)]), latent=latent,
                 depths=batch["src_depths"], depths_std=batch["src_depth_stds"], normals=nrm, mlp=mlp,
                 feature_padding=feature_padding, **kw)


def loss_and_grads(scene: Scene, rays, z, gt_rgb, white_bkgd):
    """Training-step gradients of the hot path (groundwork for BASELINE config 3): the MSE of
    DINER.calc_losses (src/models/diner.py:259-266, `criterion = MSELoss(reduction="mean")`, :61) on the rendered colours of
    given sample depths `z` (the sampler is @torch.no_grad in the reference, nerf_renderer.py:65), back-propagated by torch
    autograd through composite -> PixelNeRF.forward -> ResnetFC to the MLP parameters and the latent maps.
    Returns (loss, {param name: grad}, d loss / d latent)."""
    leaf_mlp = {k: v.detach().clone().requires_grad_(True) for k, v in scene.mlp.items()}
    leaf_lat = scene.latent.detach().clone().requires_grad_(True)
    old_mlp, old_lat = scene.mlp, scene.latent
    scene.mlp, scene.latent = leaf_mlp, leaf_lat
    try:
        _, rgb, _ = composite(scene, rays, z, white_bkgd)
        loss = F.mse_loss(rgb, gt_rgb, reduction="mean")
        loss.backward()
    finally:
        scene.mlp, scene.latent = old_mlp, old_lat
    return loss.detach(), {k: v.grad for k, v in leaf_mlp.items()}, leaf_lat.grad


def grad_digest(t, n=256):
    """Compact, order-sensitive summary of a gradient tensor for a small fixture: L2 norm, sum, and a strided sample."""
    f = t.detach().reshape(-1).double()
    idx = (torch.arange(min(n, f.numel())) * max(1, f.numel() // n)) % f.numel()
    return dict(norm=float(f.norm()), sum=float(f.sum()), sample=f[idx].float().clone(), numel=f.numel())


def psnr(pred, gt):
    """-10 log10(mse), data_range 1 (matches reference src/evaluation/eval_suite.py:66)."""
    return float(-10.0 * torch.log10(torch.mean((pred - gt) ** 2)))

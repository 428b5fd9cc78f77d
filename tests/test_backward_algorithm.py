"""CPU check of the ALGORITHM that diner_b200/csrc/backward_simt.cu implements (the CUDA code itself needs a GPU):
the same chain -- compositing backward in closed form, forward recompute with every block input kept, data gradients
through transposed weights, weight gradients as G^T.act(A), mean-over-views backward, bilinear scatter of the latent
gradient -- written with plain torch ops and compared against torch autograd through the oracle (which is pinned to the
reference's own gradients by tests/test_oracle.py::test_oracle_gradients_match_reference_golden)."""
import torch
import torch.nn.functional as F

from oracle import diner_oracle as O
from oracle import make_golden as MG


def manual_backward(scene, rays, z, g_rgb, white):
    m, NV = scene.mlp, scene.poses.shape[1]
    SB, NR, K = z.shape
    n_pre = scene.combine_layer
    n_post = scene.n_blocks - n_pre
    pts = (rays[..., None, :3] + z.unsqueeze(-1) * rays[..., None, 3:6]).reshape(SB, NR * K, 3)
    vd = rays[..., None, 3:6].expand(-1, -1, K, -1).reshape(SB, NR * K, 3)
    # ---- features / gathered latent exactly as O.query builds them
    xc_ = O.world_to_cam(scene, pts)
    zf = O.positional_encoding(xc_, scene.num_freqs, scene.freq_factor)
    vdc = torch.matmul(scene.poses[:, :, :3, :3], vd.unsqueeze(1).expand(-1, NV, -1, -1).transpose(-1, -2)).transpose(-1, -2)
    uv = O.project_uv(scene, xc_)
    zlat = O.index_latent(scene, uv).transpose(-1, -2)                       # (SB,NV,N,L)
    dd = O.index_depth(scene, uv).squeeze(-2) - xc_[..., -1]
    xin = torch.cat((zf, vdc, O.positional_encoding(dd.unsqueeze(-1), scene.num_freqs, scene.freq_factor)), -1)
    # ---- forward recompute, keeping block inputs (rows = (SB,NV,N))
    x = F.linear(xin, m["lin_in.weight"], m["lin_in.bias"])
    xa, net = [], []
    for b in range(n_pre):
        x = x + F.linear(zlat, m["lin_z.%d.weight" % b], m["lin_z.%d.bias" % b])
        xa.append(x)
        net.append(F.linear(torch.relu(x), m["blocks.%d.fc_0.weight" % b], m["blocks.%d.fc_0.bias" % b]))
        x = x + F.linear(torch.relu(net[-1]), m["blocks.%d.fc_1.weight" % b], m["blocks.%d.fc_1.bias" % b])
    xc = x.mean(dim=1)
    xci, netc = [], []
    for b in range(n_pre, scene.n_blocks):
        xci.append(xc)
        netc.append(F.linear(torch.relu(xc), m["blocks.%d.fc_0.weight" % b], m["blocks.%d.fc_0.bias" % b]))
        xc = xc + F.linear(torch.relu(netc[-1]), m["blocks.%d.fc_1.weight" % b], m["blocks.%d.fc_1.bias" % b])
    pre = F.linear(torch.relu(xc), m["lin_out.weight"], m["lin_out.bias"]).reshape(SB, NR, K, 4)
    c, sig = torch.sigmoid(pre[..., :3]), torch.relu(pre[..., 3])
    # ---- compositing backward (closed form of composite_backward_kernel)
    deltas = torch.cat([z[..., 1:] - z[..., :-1], rays[..., -1:] - z[..., -1:]], -1)
    a = 1 - torch.exp(-deltas * sig)
    t = 1 - a + 1e-10
    T = torch.cumprod(torch.cat([torch.ones_like(a[..., :1]), t], -1), -1)[..., :-1]
    w = a * T
    G = (c * g_rgb.unsqueeze(-2)).sum(-1) - (g_rgb.sum(-1, keepdim=True) if white else 0)
    Gw = G * w
    S = torch.flip(torch.cumsum(torch.flip(Gw, [-1]), -1), [-1]) - Gw
    dA = G * T - S / t
    d_pre = torch.cat([w.unsqueeze(-1) * g_rgb.unsqueeze(-2) * c * (1 - c),
                       (dA * deltas * (1 - a) * (sig > 0)).unsqueeze(-1)], -1).reshape(SB, NR * K, 4)
    grads = {}

    def wgrad(name, Gm, A):
        grads[name + ".weight"] = Gm.reshape(-1, Gm.shape[-1]).t() @ A.reshape(-1, A.shape[-1])
        grads[name + ".bias"] = Gm.reshape(-1, Gm.shape[-1]).sum(0)

    wgrad("lin_out", d_pre, torch.relu(xc))
    gx = (d_pre @ m["lin_out.weight"]) * (xc > 0)
    for i in reversed(range(n_post)):
        b = n_pre + i
        wgrad("blocks.%d.fc_1" % b, gx, torch.relu(netc[i]))
        gnet = (gx @ m["blocks.%d.fc_1.weight" % b]) * (netc[i] > 0)
        wgrad("blocks.%d.fc_0" % b, gnet, torch.relu(xci[i]))
        gx = gx + (gnet @ m["blocks.%d.fc_0.weight" % b]) * (xci[i] > 0)
    gx = (gx / NV).unsqueeze(1).expand(-1, NV, -1, -1)
    gz = torch.zeros_like(zlat)
    for b in reversed(range(n_pre)):
        wgrad("blocks.%d.fc_1" % b, gx, torch.relu(net[b]))
        gnet = (gx @ m["blocks.%d.fc_1.weight" % b]) * (net[b] > 0)
        wgrad("blocks.%d.fc_0" % b, gnet, torch.relu(xa[b]))
        gx = gx + (gnet @ m["blocks.%d.fc_0.weight" % b]) * (xa[b] > 0)
        wgrad("lin_z.%d" % b, gx, zlat)
        gz = gz + gx @ m["lin_z.%d.weight" % b]
    wgrad("lin_in", gx, xin)
    # ---- latent scatter through the bilinear taps (latent_taps of common.cuh: border clamp, align_corners=False)
    SBn, NVn, L, Hl, Wl = scene.latent.shape
    size = torch.tensor([Wl, Hl], dtype=torch.float32)
    uvl = uv * ((size - scene.feature_padding * 2) / size)
    px = (((uvl[..., 0] + 1) * Wl - 1) / 2).clamp(0, Wl - 1)
    py = (((uvl[..., 1] + 1) * Hl - 1) / 2).clamp(0, Hl - 1)
    x0, y0 = px.floor().long(), py.floor().long()
    x1, y1 = (x0 + 1).clamp(max=Wl - 1), (y0 + 1).clamp(max=Hl - 1)
    wx, wy = px - x0, py - y0
    d_lat = torch.zeros(SBn * NVn, Hl * Wl, L)
    gzf = gz.reshape(SBn * NVn, -1, L)
    for (yy, xx, ww) in ((y0, x0, (1 - wx) * (1 - wy)), (y0, x1, wx * (1 - wy)), (y1, x0, (1 - wx) * wy), (y1, x1, wx * wy)):
        idx = (yy * Wl + xx).reshape(SBn * NVn, -1)
        d_lat.scatter_add_(1, idx.unsqueeze(-1).expand(-1, -1, L), gzf * ww.reshape(SBn * NVn, -1, 1))
    return grads, d_lat.permute(0, 2, 1).reshape(SBn, NVn, L, Hl, Wl)


import pytest


@pytest.mark.parametrize("cfg", [
    dict(H=32, W=32, NV=4, SB=1, near=1.0, far=2.5, K=16, C=100, G=6, white=True, nr=24, seed=31),
    dict(H=32, W=48, NV=2, SB=2, near=1.0, far=2.5, K=12, C=100, G=4, white=False, nr=16, seed=32),
], ids=["nv4_white", "sb2_nv2_black"])
def test_manual_backward_chain_matches_autograd(cfg):
    batch, latent, mlp, rays, noise = MG.case_inputs(cfg)
    scene = O.make_scene_state(batch, latent, mlp)
    z = O.fill_up_uniform(O.sample_depthguided(scene, rays, cfg["K"], cfg["C"], cfg["G"], noise["u_coarse"], noise["g_noise"]),
                          rays, noise["u_fill"])
    from diner_b200 import synthetic as S
    gt = S.hash_uniform((cfg["SB"], rays.shape[1], 3), cfg["seed"], 950)
    loss, g_ref, lat_ref = O.loss_and_grads(scene, rays, z, gt, cfg["white"])
    with torch.no_grad():
        _, rgb, _ = O.composite(scene, rays, z, cfg["white"])
        g_rgb = 2.0 * (rgb - gt) / rgb.numel()
        grads, d_lat = manual_backward(scene, rays, z, g_rgb, cfg["white"])
    assert set(grads) == set(g_ref)
    for k, v in g_ref.items():
        scale = float(v.abs().max().clamp_min(1e-12))
        assert float((grads[k] - v).abs().max()) / scale <= 2e-4, (k, float((grads[k] - v).abs().max()), scale)
    scale = float(lat_ref.abs().max())
    assert scale > 0 and float((d_lat - lat_ref).abs().max()) / scale <= 2e-4

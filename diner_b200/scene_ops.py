"""Once-per-scene helpers of PixelNeRF.encode that stay in PyTorch for now (SURVEY §8(f) rank 1 'next' row)."""
import torch
import torch.nn.functional as F


@torch.no_grad()
def depth2normal(dmap, K):
    """(N,1,H,W) depth + (N,3,3) intrinsics -> (N,3,H,W) unit normals, zero where depth is zero.

    Same result as the reference's src/util/depth2normal.py:6-87: back-project, central differences on the
    replicate-padded point map, then pixels next to a hole copy the normal of their in-surface neighbour.
    """
    N, _, H, W = dmap.shape
    dev = dmap.device
    ys, xs = torch.meshgrid(torch.arange(0.5, H, 1., device=dev), torch.arange(0.5, W, 1., device=dev), indexing="ij")
    pix = torch.stack((xs, ys), -1).reshape(1, H * W, 2).repeat(N, 1, 1)
    pix -= K[:, [0, 1], -1].unsqueeze(-2)
    pix /= K[:, [0, 1], [0, 1]].unsqueeze(-2)
    ray = torch.cat((pix, torch.ones_like(pix[..., :1])), dim=-1).view(N, H, W, 3)
    pts = F.pad((ray * dmap.view(N, H, W, 1)).permute(0, 3, 1, 2), [1, 1, 1, 1], mode="replicate")
    below, above = pts[:, :, 2:, 1:-1], pts[:, :, :-2, 1:-1]
    right, left = pts[:, :, 1:-1, 2:], pts[:, :, 1:-1, :-2]
    n = torch.linalg.cross((below - above).permute(0, 2, 3, 1), (right - left).permute(0, 2, 3, 1), dim=-1)
    n = n / torch.norm(n, p=2, dim=-1, keepdim=True)
    dy = (above[:, 0] == 0).long() - (below[:, 0] == 0).long()
    dx = (left[:, 0] == 0).long() - (right[:, 0] == 0).long()
    fix = (dy != 0) | (dx != 0)
    ni, yi, xi = torch.where(fix)
    n[ni, yi, xi] = n[ni, (yi + dy[fix]).clamp(0, H - 1), (xi + dx[fix]).clamp(0, W - 1)]
    n[dmap[:, 0] == 0] = 0
    return n.permute(0, 3, 1, 2)

"""GPU-box debugging: (1) do the MMA issue orders (tail_kb) give the same bits, and how deterministic is each; (2) error of the
tcgen05 modes against the fp32 CUDA-core mode: magnitude AND sign (a biased error points at the accumulator, not the operands)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import torch
from diner_b200 import synthetic as S
from diner_b200.synthetic import product_model

NV, SB, K, nr = 4, 1, 64, 4000
H = W = 32
batch = S.make_scene(H, W, NV, SB, 1.0, 2.5, 100 + NV)
latent = torch.randn(SB, NV, 512, (H + 128) // 2, (W + 128) // 2, generator=torch.Generator().manual_seed(100 + NV)) * 0.5
mlp = S.make_mlp_state(seed=100 + NV)
rays = S.gen_rays(batch["target_extrinsics"], batch["target_intrinsics"], W, H, torch.full((SB,), 1.0), torch.full((SB,), 2.5)).view(SB, H * W, 8)
rays = rays[:, torch.arange(nr) % (H * W)].contiguous().cuda()
model = product_model(batch, latent, mlp, "cuda", "parity")
ctx = model.context()
z = ctx.sample(rays, K, 200, 12, dict(seed=nr))
pts = (rays[..., None, :3] + z.unsqueeze(-1) * rays[..., None, 3:6]).reshape(SB, -1, 3).contiguous()
vd = rays[..., None, 3:6].expand(-1, -1, K, -1).reshape(SB, -1, 3).contiguous()
ref = ctx.query(pts, vd, 0)
base = None
for tail in (0, 1, 2, 3, 4, 3, 4, 0):
    ctx.set_option("tail_kb", tail)
    outs = [ctx.query(pts, vd, 1) for _ in range(3)]
    det = all(torch.equal(outs[0], o) for o in outs[1:])
    if base is None:
        base = outs[0]
    d = (outs[0] - base).abs()
    nd = int((d.max(-1).values > 0).sum())
    print("tail_kb=%d deterministic=%s  vs tail 0: samples differing %d / %d, max |diff| %.3g | vs fp32: max rgb %.3g, max rel sigma %.3g"
          % (tail, det, nd, d.shape[1], float(d.max()), float((outs[0][..., :3] - ref[..., :3]).abs().max()),
             float(((outs[0][..., 3] - ref[..., 3]).abs() / (1 + ref[..., 3].abs())).max())), flush=True)
    if nd:
        idx = (d.max(-1).values[0] > 0).nonzero().flatten()
        print("   differing sample indices (first 24):", idx[:24].tolist(), " spacing hist mod 64:", torch.bincount(idx % 64, minlength=64).tolist())
ctx.set_option("tail_kb", 3)
out = ctx.query(pts, vd, 1)
sig, sig_ref = out[..., 3], ref[..., 3]
m = sig_ref > 1e-3
rel = ((sig - sig_ref) / sig_ref)[m]
print("sigma relative error parity vs fp32: mean %.3g  std %.3g  (n=%d)  -> bias/std = %.2f" % (float(rel.mean()), float(rel.std()), int(m.sum()), float(rel.mean() / rel.std())))
fast = ctx.query(pts, vd, 2)
relf = ((fast[..., 3] - sig_ref) / sig_ref)[m]
print("sigma relative error fast vs fp32:   mean %.3g  std %.3g" % (float(relf.mean()), float(relf.std())))

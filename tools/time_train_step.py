"""GPU-box timing of the pieces of one config-3 training step (synchronising between them): where the wall time goes."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import torch
from diner_b200 import synthetic as S
from diner_b200.nerf_renderer import NeRFRendererDGS
from diner_b200.predict import calc_losses

Ht = Wt = 256
SBt, NVt, Kt, RB = 4, 4, 128, 4096
dev = torch.device("cuda", 0)
batch = S.make_scene(Ht, Wt, NVt, SBt, 1.0, 2.5, 0)
gen = torch.Generator().manual_seed(0)
latent = torch.randn(SBt, NVt, 512, (Ht + 128) // 2, (Wt + 128) // 2, generator=gen) * 0.5
model = S.product_model(batch, latent, S.make_mlp_state(seed=0), dev, "parity").train()
model.encoder.latent = model.encoder.latent.detach().clone().requires_grad_(True)
model.encoder.scene_version += 1
rend = NeRFRendererDGS(n_samples=Kt, n_depth_candidates=1000, n_gaussian=int(15 * Kt / 40), white_bkgd=True)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
b["target_rgb"] = torch.rand(SBt, 3, Ht, Wt, generator=gen).to(dev)
opt = torch.optim.Adam(list(model.mlp_fine.parameters()) + [model.encoder.latent], lr=1e-4)
g = torch.Generator().manual_seed(1)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); opt.zero_grad(set_to_none=True)
    t1 = T(); ctx = model.context()
    t2 = T(); loss = calc_losses(model, rend, b, 1.0, 2.5, RB, generator=g, encode=False)["total"]
    t3 = T(); loss.backward()
    t4 = T(); opt.step()
    t5 = T()
    print("step %d: zero_grad %.1f ms | context (weight repack, scene upload, tables) %.1f | forward+loss %.1f | backward %.1f | adam %.1f | total %.1f"
          % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t5 - t0)), flush=True)

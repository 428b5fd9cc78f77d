"""Multi-GPU check (run under torchrun on the GPU box): the ray-sharded render + in-place NCCL all-gather must reproduce the
single-GPU image BIT FOR BIT on every rank (rays are independent; counter-based sampler noise is keyed by the
logical ray index).  Prints one line on rank 0."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import torch
import torch.distributed as dist
from diner_b200 import synthetic as S
from diner_b200.multi_gpu import render_sharded
from diner_b200.nerf_renderer import NeRFRendererDGS

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H = W = 96
batch = S.make_scene(H, W, 4, 1, 1.0, 2.5, 3)
latent = torch.randn(1, 4, 512, (H + 128) // 2, (W + 128) // 2, generator=torch.Generator().manual_seed(3)) * 0.5
model = S.product_model(batch, latent, S.make_mlp_state(seed=3), dev, "parity")
rend = NeRFRendererDGS(n_samples=32, n_depth_candidates=200, n_gaussian=12, white_bkgd=True)
rend.noise = dict(seed=1234)
rays = S.gen_rays(batch["target_extrinsics"], batch["target_intrinsics"], W, H, torch.full((1,), 1.0), torch.full((1,), 2.5)).view(1, H * W, 8)
rays = rays[:, :H * W - 5].contiguous().to(dev)            # ragged: not a multiple of the world size
whole = rend.render_packed(model, rays)
img = render_sharded(lambda r, out, off: rend.render_packed(model, r, out=out, ray_offset=off), rays, packed=True, return_packed=True)
img_w = render_sharded(lambda r, out, off: rend.render_packed(model, r, out=out, ray_offset=off), rays, packed=True, return_packed=True,
                       weights=[1.0 + 0.07 * ((r * 5) % world) for r in range(world)])          # speed-weighted (uneven) shards
ok = torch.tensor([int(torch.equal(img, whole) and torch.equal(img_w, whole))], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded render over %d GPUs == single-GPU render, bit for bit, on every rank: %s  (%d rays, rgb mean %.4f)" % (
        world, bool(ok.item()), rays.shape[1], float(whole[..., :3].mean())))
dist.destroy_process_group()
sys.exit(0 if ok.item() else 1)

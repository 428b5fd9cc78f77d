"""GPU-box profiling driver: renders a slab of the bench workload a few times (for ncu / timing experiments).
usage: python tools/profile_run.py <mode> <n_rays> [reps] [workload]"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import torch
import bench
from tests.common import product_model
from diner_b200.nerf_renderer import NeRFRendererDGS

mode, n_rays = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if len(sys.argv) > 4:
    bench.select_workload(sys.argv[4])                 # e.g. stress1024 (BASELINE configs[4]); default: the headline workload
batch, latent, mlp, rays = bench.build_inputs()
model = product_model(batch, latent, mlp, "cuda", mode)
rend = NeRFRendererDGS(n_samples=bench.K, n_depth_candidates=bench.C, n_gaussian=bench.G, white_bkgd=False)
start = (bench.H // 2) * bench.W - n_rays // 2        # rows around the image centre (foreground)
r = rays[:, start:start + n_rays].contiguous().cuda()
ctx = model.context()
if os.environ.get("DINER_RAY_IMAGE_WIDTH") is None:
    ctx.set_option("ray_image_width", bench.W)        # the slab is whole rows of the row-major image
ctx.set_timing(True)
for i in range(reps):
    with torch.no_grad():
        rend(model, r)
    torch.cuda.synchronize()
    print("rep %d:" % i, {k: round(v, 3) for k, v in ctx.last_stage_ms().items()}, flush=True)

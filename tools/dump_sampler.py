"""Debug helper (GPU box): dumps the CUDA sampler's outputs for the golden cases into gpurun_out/."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import torch
from oracle import make_golden as MG
from tests.common import product_model

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for name in sys.argv[1:] or list(MG.CASES):
    cfg = MG.CASES[name]
    batch, latent, mlp, rays, noise = MG.case_inputs(cfg)
    model = product_model(batch, latent, mlp, "cuda")
    nz = {k: v.cuda().contiguous() for k, v in noise.items()}
    z, zd = model.context().sample(rays.cuda(), cfg["K"], cfg["C"], cfg["G"], nz, want_dgs=True)
    torch.save(dict(z=z.cpu(), zd=zd.cpu()), os.path.join(ROOT, "gpurun_out", "sampler_%s.pt" % name))
    print(name, "dumped")

"""GPU-box check of the tcgen05 path against the fp32 CUDA-core path (same library), per mode / cluster size.
usage: python tools/tc_check.py <mode> <cluster> [n_rays] [K]"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import torch
from oracle import make_golden as MG
from tests.common import product_model, renderer_for

mode, cl = sys.argv[1], int(sys.argv[2])
nr = int(sys.argv[3]) if len(sys.argv) > 3 else 96
K = int(sys.argv[4]) if len(sys.argv) > 4 else 64
os.environ["DINER_TC_CLUSTER"] = str(cl)
cfg = dict(H=64, W=64, NV=4, SB=1, near=0.3211, far=1.2041, K=K, C=1000, G=int(15 * K / 40), white=False, nr=nr, seed=2)
batch, latent, mlp, rays, noise = MG.case_inputs(cfg)
ref_model = product_model(batch, latent, mlp, "cuda", "fp32")
rend = renderer_for(cfg, noise)
rays = rays.cuda()
with torch.no_grad():
    z = ref_model.context().sample(rays, cfg["K"], cfg["C"], cfg["G"], rend.noise)
    w0, rgb0, d0 = rend.composite(ref_model, rays, z)
    pts = (rays[..., None, :3] + z.unsqueeze(-1) * rays[..., None, 3:6]).reshape(1, -1, 3).contiguous()
    vd = rays[..., None, 3:6].expand(-1, -1, K, -1).reshape(1, -1, 3).contiguous()
    out0 = ref_model(pts, vd)
    model = product_model(batch, latent, mlp, "cuda", mode)
    t0 = time.time()
    out1 = model(pts, vd)
    model.context().debug_sync()
    t1 = time.time()
    w1, rgb1, d1 = rend.composite(model, rays, z)
    torch.cuda.synchronize()
print("mode=%s cl=%d  samples=%d  query max|d rgb|=%.3g  max rel|d sigma|=%.3g   composite max|d rgb|=%.3g |d depth|=%.3g  (first call %.1f ms)" % (
    mode, cl, pts.shape[1], (out1[..., :3] - out0[..., :3]).abs().max(),
    ((out1[..., 3] - out0[..., 3]).abs() / (1 + out0[..., 3].abs())).max(),
    (rgb1 - rgb0).abs().max(), (d1 - d0).abs().max(), (t1 - t0) * 1e3))
bad = (out1 - out0).abs().max(-1).values[0] > 1e-2
if bad.any():
    idx = bad.nonzero()[:, 0]
    print("  bad samples: %d of %d; first idx %s; idx%%64 %s" % (bad.sum(), bad.numel(), idx[:12].tolist(), (idx[:12] % 64).tolist()))
    print("  ref", out0[0, idx[0]].tolist(), "got", out1[0, idx[0]].tolist())

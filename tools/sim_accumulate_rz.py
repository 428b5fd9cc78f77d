"""CPU simulation of the tcgen05 parity-mode arithmetic (fp16 hi/lo split operands, W_SCALE-scaled weights, K = 16 per MMA, three
MMAs per product) with the accumulator either rounding to nearest or TRUNCATING (round toward zero) at every MMA step, on a golden
case: `python tools/sim_accumulate_rz.py nogauss_nv8`.  The truncating model reproduces the error level measured on the B200
(rgb ~4e-6, depth ~1.6e-5 on that case); round-to-nearest accumulation would give ~8e-7.  Test infrastructure (uses oracle/)."""
import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from oracle import diner_oracle as O
from oracle import make_golden as MG
torch.set_num_threads(8)
def split(x, dt, n):
    parts = []; r = x.clone()
    for _ in range(n):
        p = r.to(dt).to(torch.float32); parts.append(p); r = r - p
    return parts
def rz32(t64):
    r = t64.float()
    over = r.double().abs() > t64.abs()
    r2 = torch.nextafter(r, torch.zeros_like(r))
    return torch.where(over, r2, r)
MODE = None
def lin(x, w, b=None):
    shp = x.shape; x2 = x.reshape(-1, shp[-1])
    Kd = x2.shape[1]
    ah, al = split(x2, torch.float16, 2); wh, wl = split(w*64, torch.float16, 2)
    pad = (-Kd) % 16
    def P(t): return F.pad(t, (0, pad)).double()
    ah, al, wh, wl = P(ah), P(al), P(wh), P(wl)
    acc = torch.zeros(x2.shape[0], w.shape[0])
    nchunk = ah.shape[1] // 16
    for c in range(nchunk):
        sl = slice(16*c, 16*c+16)
        for (a_, w_) in ((ah, wh), (al, wh), (ah, wl)):
            chunk = a_[:, sl] @ w_[:, sl].T
            t = acc.double() + chunk
            acc = rz32(t) if MODE == "rz" else t.float()
    y = acc / 64
    if b is not None: y = y + b
    return y.reshape(*shp[:-1], w.shape[0])
_lin = F.linear
name = sys.argv[1]
g = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "%s.pt" % name)); cfg = g["cfg"]
batch, latent, mlp, rays, noise = MG.case_inputs(cfg)
scene = O.make_scene_state(batch, latent, mlp)
for m in ("rn", "rz"):
    MODE = m
    O.F.linear = lin
    with torch.no_grad():
        w, rgb, depth = O.composite(scene, rays, g["z_filled"], cfg["white"])
    O.F.linear = _lin
    print("%s fp16x3 accumulate-%s per 16-wide MMA: rgb %.3g depth %.3g w %.3g" % (name, m, (rgb-g["rgb"]).abs().max(), (depth-g["depth"]).abs().max(), (w-g["weights"]).abs().max()), flush=True)

"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported read-only from
/root/reference via oracle/ref_import.py) on deterministic synthetic inputs with injected noise.

Run in the build container only:  python oracle/make_golden.py
Test infrastructure; never imported by the product path.

How the reference's data-dependent random draws are made reproducible without editing it: the
three calls (nerf_renderer.py:57,188,390) go through torch.rand_like / torch.randn_like, which are
swapped for functions that return slices of dense hash noise.  The masks needed to gather the dense
noise for calls 2 and 3 (rays with non-zero opaque likelihood; missing slots after the sort) are
predicted with oracle/diner_oracle.py -- if the oracle predicted them wrongly the shapes would not
match or the outputs stored here would disagree with the oracle in tests/test_oracle.py.
"""
import os
import sys
import types
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import diner_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402
from diner_b200 import synthetic as S  # noqa: E402

CASES = {
    # name: H, W, NV, SB, near, far, K, C, G, white_bkgd, n_rays(subset), seed
    "cfg1_face64": dict(H=64, W=64, NV=4, SB=1, near=1.0, far=2.5, K=32, C=1000, G=12, white=True, nr=192, seed=1),
    "cfg2_dtu64": dict(H=64, W=64, NV=4, SB=1, near=0.3211, far=1.2041, K=64, C=1000, G=24, white=False, nr=96, seed=2),
    "sb2_nv2_rect": dict(H=48, W=64, NV=2, SB=2, near=1.0, far=2.5, K=40, C=250, G=15, white=True, nr=64, seed=3),
    "nogauss_nv8": dict(H=32, W=32, NV=8, SB=1, near=1.0, far=2.5, K=16, C=128, G=0, white=False, nr=64, seed=4),
}


# Oracle-only pins at the sample counts of BASELINE configs[2] / configs[4] (not in CASES: the GPU parity tests iterate CASES,
# and a new GPU case has to be validated on hardware before it may gate a round)
EXTRA_CASES = {
    "k128_sb2_face": dict(H=64, W=64, NV=4, SB=2, near=1.0, far=2.5, K=128, C=1000, G=48, white=True, nr=24, seed=6),
    "k256_nv8": dict(H=64, W=64, NV=8, SB=1, near=1.0, far=2.5, K=256, C=1000, G=96, white=False, nr=24, seed=7),
}

_CASE_CACHE = {}


def case_inputs(cfg):
    """Everything both the reference and the oracle / CUDA path consume, regenerated from the seed (memoised per
    process: the hash-based latent takes seconds to build and many tests share a case; callers must not modify it in place)."""
    key = tuple(sorted(cfg.items()))
    if key not in _CASE_CACHE:
        _CASE_CACHE[key] = _case_inputs(cfg)
    batch, latent, mlp, rays, noise = _CASE_CACHE[key]
    return dict(batch), latent, dict(mlp), rays, dict(noise)


def _case_inputs(cfg):
    batch = S.make_scene(cfg["H"], cfg["W"], cfg["NV"], cfg["SB"], cfg["near"], cfg["far"], cfg["seed"])
    Hl, Wl = (cfg["H"] + 128) // 2, (cfg["W"] + 128) // 2
    latent = S.make_latent(cfg["SB"], cfg["NV"], 512, Hl, Wl, cfg["seed"])
    mlp = S.make_mlp_state(seed=cfg["seed"])
    SB, H, W = cfg["SB"], cfg["H"], cfg["W"]
    rays = S.gen_rays(batch["target_extrinsics"], batch["target_intrinsics"], W, H,
                      torch.full((SB,), cfg["near"]), torch.full((SB,), cfg["far"])).view(SB, H * W, 8)
    pick = (torch.arange(cfg["nr"]) * (H * W // cfg["nr"]) + 7) % (H * W)
    rays = rays[:, pick].contiguous()
    NR = rays.shape[1]
    noise = dict(u_coarse=S.hash_uniform((SB, NR, cfg["C"]), cfg["seed"], 901),
                 g_noise=S.hash_normal((SB, NR, max(cfg["G"], 1)), cfg["seed"], 902)[..., :cfg["G"]],
                 u_fill=S.hash_uniform((SB, NR, cfg["K"]), cfg["seed"], 903))
    return batch, latent, mlp, rays, noise


def build_reference_model(ns, batch, latent, mlp):
    sys.modules["diner_ref_image_encoder"] = ns.image_encoder
    sys.modules["diner_ref_resnetfc"] = ns.resnetfc
    D = ref_import._DotMap
    model = ns.pixelnerf.PixelNeRF(
        poscode_conf=D(kwargs=dict(num_freqs=6, freq_factor=6.28, include_input=True)),
        encoder_conf=D(module="diner_ref_image_encoder.SpatialEncoder",
                       kwargs=dict(image_padding=64, padding_pe=4, pretrained=False)),
        mlp_fine_conf=D(module="diner_ref_resnetfc.ResnetFC",
                        kwargs=dict(n_blocks=5, d_hidden=512, combine_layer=3, combine_type="average")))
    model.mlp_fine.load_state_dict(mlp)
    SB, NV = batch["src_depths"].shape[:2]
    H, W = batch["src_depths"].shape[-2:]
    K = batch["src_intrinsics"]
    enc = model.encoder
    # what PixelNeRF.encode / SpatialEncoder.forward leave behind (pixelnerf.py:44-51, image_encoder.py:232-237,290-291),
    # with the ResNet trunk's output replaced by the supplied latent maps
    enc.depths, enc.depths_std = batch["src_depths"], batch["src_depth_stds"]
    enc.normals = ns.depth2normal.depth2normal(batch["src_depths"].flatten(end_dim=1),
                                               K.flatten(end_dim=1)).reshape(SB, NV, 3, H, W)
    enc.nviews, enc.nobjects = NV, SB
    enc.latent = latent
    model.poses = batch["src_extrinsics"]
    model.c = K[:, :, :2, -1]
    model.focal = K[:, :, torch.tensor([0, 1]), torch.tensor([0, 1])]
    model.image_shape[0] = W
    model.image_shape[1] = H
    return model.eval()


class InjectNoise:
    """Swap torch.rand_like / randn_like for a scripted sequence of tensors."""

    def __init__(self, seq):
        self.seq = list(seq)

    def __enter__(self):
        self._r, self._n = torch.rand_like, torch.randn_like

        def take(t, *a, **k):
            out = self.seq.pop(0)
            assert out.shape == t.shape, (out.shape, t.shape)
            return out.clone()

        torch.rand_like = take
        torch.randn_like = take
        return self

    def __exit__(self, *a):
        torch.rand_like, torch.randn_like = self._r, self._n


def run_reference(ns, cfg, batch, latent, mlp, rays, noise):
    model = build_reference_model(ns, batch, latent, mlp)
    rend = ns.nerf_renderer.NeRFRendererDGS(n_samples=cfg["K"], n_depth_candidates=cfg["C"],
                                            n_gaussian=cfg["G"], white_bkgd=cfg["white"])
    scene = O.make_scene_state(batch, latent, mlp)
    # masks predicted by the oracle, used ONLY to gather the dense noise for the reference's masked draws
    z_or, aux = O.sample_depthguided(scene, rays, cfg["K"], cfg["C"], cfg["G"], noise["u_coarse"],
                                     noise["g_noise"], return_aux=True)
    _, miss = O.fill_up_uniform(z_or, rays, noise["u_fill"], return_mask=True)
    seq = [noise["u_coarse"].reshape(-1, cfg["C"])]
    if cfg["G"] > 0:
        seq.append(noise["g_noise"][aux["ray_mask"]])
    with torch.no_grad():
        with InjectNoise(seq):
            z0 = rend.sample_depthguided(rays, model, n_samples=cfg["K"], n_candidates=cfg["C"],
                                         n_gaussian=cfg["G"])
        with InjectNoise([noise["u_fill"].reshape(-1, cfg["K"])[miss.reshape(-1, cfg["K"])]]):
            z1 = rend.fill_up_uniform_samples(z0.clone(), rays)
        w, rgb, depth = rend.composite(model, rays, z1)
        # per-sample network output on the final sample positions (stage-wise parity, no RNG involved)
        pts = (rays[..., None, :3] + z1.unsqueeze(-1) * rays[..., None, 3:6]).reshape(cfg["SB"], -1, 3)
        vd = rays[..., None, 3:6].expand(-1, -1, cfg["K"], -1).reshape(cfg["SB"], -1, 3)
        net = model(pts, viewdirs=vd)
    return dict(z_depthguided=z0, z_filled=z1, weights=w, rgb=rgb, depth=depth, net_out=net,
                normals=model.encoder.normals.clone())


GRAD_CASE = dict(H=64, W=64, NV=4, SB=1, near=1.0, far=2.5, K=32, C=1000, G=12, white=True, nr=48, seed=1)   # cfg1 inputs, 48 rays


def grad_case_inputs():
    cfg = GRAD_CASE
    batch, latent, mlp, rays, noise = case_inputs(cfg)
    gt = S.hash_uniform((cfg["SB"], rays.shape[1], 3), cfg["seed"], 950)
    return cfg, batch, latent, mlp, rays, noise, gt


def make_grad_golden(ns, outdir):
    """Loss and gradients of the UNMODIFIED reference (autograd through NeRFRendererDGS.composite / PixelNeRF.forward /
    ResnetFC with the MSE of diner.py:61,266) on the reference's own sample depths: the pin for the backward row."""
    cfg, batch, latent, mlp, rays, noise, gt = grad_case_inputs()
    ref = run_reference(ns, cfg, batch, latent, mlp, rays, noise)
    z = ref["z_filled"]
    model = build_reference_model(ns, batch, latent, mlp)
    lat = latent.clone().requires_grad_(True)
    model.encoder.latent = lat
    for p in model.parameters():
        p.requires_grad_(True)
    rend = ns.nerf_renderer.NeRFRendererDGS(n_samples=cfg["K"], n_depth_candidates=cfg["C"], n_gaussian=cfg["G"],
                                            white_bkgd=cfg["white"])
    _, rgb, _ = rend.composite(model, rays, z)
    loss = torch.nn.MSELoss(reduction="mean")(rgb, gt)
    loss.backward()
    grads = {k: O.grad_digest(p.grad) for k, p in model.mlp_fine.named_parameters()}
    torch.save(dict(cfg=cfg, z=z.contiguous(), loss=float(loss), grads=grads, latent_grad=O.grad_digest(lat.grad, 4096)),
               os.path.join(outdir, "grads_cfg1_face64.pt"))
    print("grad golden: loss %.6f, |dL/dlatent| %.4g, |dL/dW(lin_out)| %.4g" % (
        float(loss), float(lat.grad.norm()), float(model.mlp_fine.lin_out.weight.grad.norm())))


GEN_RAYS_CASES = {   # name: H, W, SB, near, far, seed  (caller side of the hot path: src/util/cam_geometry.py:5-48)
    "gen_rays_24x16": dict(H=16, W=24, SB=2, near=1.0, far=2.5, seed=3),
    "gen_rays_64x96": dict(H=64, W=96, SB=2, near=0.3211, far=1.2041, seed=3),
}


def gen_rays_inputs(cfg):
    batch = S.make_scene(cfg["H"], cfg["W"], 4, cfg["SB"], cfg["near"], cfg["far"], cfg["seed"])
    return batch["target_extrinsics"], batch["target_intrinsics"]


def make_gen_rays_golden(ns, outdir):
    """Rays of the UNMODIFIED reference gen_rays for the synthetic target cameras (pins diner_b200.synthetic.gen_rays,
    the restatement the tests feed to every render, and through it diner_gen_rays / diner_render_image)."""
    out = {}
    for name, cfg in GEN_RAYS_CASES.items():
        E, K = gen_rays_inputs(cfg)
        rays = ns.cam_geometry.gen_rays(E, K, cfg["W"], cfg["H"], torch.full((cfg["SB"],), cfg["near"]),
                                        torch.full((cfg["SB"],), cfg["far"]))
        out[name] = dict(cfg=cfg, rays=rays.contiguous())
    torch.save(out, os.path.join(outdir, "gen_rays.pt"))


def make_softplus_golden(ns, outdir):
    """Output of the UNMODIFIED reference ResnetFC with Softplus activations (beta = 2, resnetfc.py:124-127) on hashed inputs:
    pins the oracle's (and through it the fp32 kernels') Softplus branch."""
    sd = S.make_mlp_state(d_in=55, d_latent=32, d_hidden=64, seed=2)
    net = ns.resnetfc.ResnetFC(55, 4, 5, 32, 64, beta=2.0, combine_layer=3)
    net.load_state_dict(sd)
    zx = S.hash_normal((2, 4, 9, 32 + 55), 3)
    with torch.no_grad():
        out = net(zx, combine_dim=1)
    torch.save(dict(beta=2.0, mlp_args=dict(d_in=55, d_latent=32, d_hidden=64, seed=2), zx_args=((2, 4, 9, 87), 3), out=out.contiguous()),
               os.path.join(outdir, "softplus_resnetfc.pt"))


def main():
    ns = ref_import.load()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    make_gen_rays_golden(ns, outdir)
    make_softplus_golden(ns, outdir)
    if "--rays-only" in sys.argv:
        return
    if "--extra-only" not in sys.argv:
        make_grad_golden(ns, outdir)
    if "--grads-only" in sys.argv:
        return
    todo = dict(EXTRA_CASES) if "--extra-only" in sys.argv else dict(CASES, **EXTRA_CASES)
    for name, cfg in todo.items():
        batch, latent, mlp, rays, noise = case_inputs(cfg)
        ref = run_reference(ns, cfg, batch, latent, mlp, rays, noise)
        pix_alpha = ref["weights"].sum(-1)
        print("%-14s rgb[%.3f,%.3f] mean %.3f  alpha mean %.3f (min %.3f max %.3f)  depth mean %.3f  "
              "empty-after-dgs %.2f" % (name, ref["rgb"].min(), ref["rgb"].max(), ref["rgb"].mean(),
                                        pix_alpha.mean(), pix_alpha.min(), pix_alpha.max(),
                                        ref["depth"].mean(), (ref["z_depthguided"] == 0).float().mean()))
        torch.save(dict(cfg=cfg, **{k: v.contiguous() for k, v in ref.items()}),
                   os.path.join(outdir, name + ".pt"))


if __name__ == "__main__":
    main()

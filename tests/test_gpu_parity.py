"""GPU parity tests proper: the CUDA path (through the C ABI) against the committed reference
goldens (tests/golden, produced by the unmodified reference on torch-CPU), against the CPU oracle,
and against the same oracle executed as eager PyTorch ON THE GPU (the reference's own deployment
target: src/models runs on cuda in python_scripts/train.py / create_prediction_folder.py).

Tolerances: north_star asks for 1e-4 abs on rgb / depth (TOL).  The tcgen05 parity mode (fp16 hi/lo
split operands, fp32 accumulation in TMEM) is held to PAR_TOL = 5e-5, half the bar, wherever the sample
depths are given; measured: <= 1.3e-5 at the headline size, <= 3.7e-5 in the worst small case.  What is
left is not operand precision (the split keeps 22 bits) but the tensor core's accumulator, which truncates
instead of rounding at each of the 96 MMA steps of a 512-wide layer (tools/sim_accumulate_rz.py reproduces
the measured error level on the CPU with exactly that model, and 8e-7 with round-to-nearest accumulation).
Stage-wise comparisons (same sample depths in) must meet the bar on EVERY ray, NaN counting as a
failure.  End-to-end comparisons additionally go through the sampler's discontinuous decisions
(nearest-pixel lookups, "likelihood != 0" membership of the shortlist).  Those decisions hinge on the
last ulp of erf(), on which torch-CPU (Sleef) and torch-CUDA (CUDA libm erff) disagree -- i.e. the
reference disagrees with itself across its two backends on a few ill-conditioned rays.  The CUDA path
must therefore (a) match the reference-on-cuda on ALL rays, and (b) differ from the CPU goldens only
on rays where the reference-on-cuda differs from them as well.
"""
import os

import pytest
import torch

from oracle import diner_oracle as O
from oracle import make_golden as MG
from tests.common import product_model, renderer_for, oracle_scene_on, ray_err

pytestmark = pytest.mark.gpu
CASES = list(MG.CASES)
ALL_CASES = list(MG.CASES) + list(MG.EXTRA_CASES)
TOL = 1e-4          # north_star bar
PAR_TOL = 5e-5      # what the fp16x3 parity mode is held to on given sample depths (half the bar; see the module docstring)
Z_TOL = 1e-5        # sample depths: a different shortlist decision moves a sample by >= one candidate step (>= 1e-3)


def _bad(e, tol):
    """Rays beyond tol; non-finite errors are bad (NaN compares False against everything)."""
    return ~(e <= tol)


def _load(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name + ".pt"))
    cfg = g["cfg"]
    batch, latent, mlp, rays, noise = MG.case_inputs(cfg)
    return g, cfg, batch, latent, mlp, rays, noise


@pytest.mark.parametrize("name", ALL_CASES)
def test_sampler_matches_reference(golden_dir, name):
    """sample_depthguided + fill_up_uniform_samples (nerf_renderer.py:65-190,367-397) with the reference's noise injected:
    (a) against the reference algorithm run on cuda: every ray; (b) against the CPU goldens: every ray on which the
    reference's two backends agree with each other."""
    g, cfg, batch, latent, mlp, rays, noise = _load(golden_dir, name)
    model = product_model(batch, latent, mlp, "cuda")
    nz = {k: v.cuda().contiguous() for k, v in noise.items()}
    z, zd = model.context().sample(rays.cuda(), cfg["K"], cfg["C"], cfg["G"], nz, want_dgs=True)
    assert bool(torch.isfinite(z).all()) and bool(torch.isfinite(zd).all())
    # (a) reference algorithm as eager PyTorch on this GPU
    sc = oracle_scene_on(O.make_scene_state(batch, latent, mlp), "cuda")
    with torch.no_grad():
        zc_dgs = O.sample_depthguided(sc, rays.cuda(), cfg["K"], cfg["C"], cfg["G"], nz["u_coarse"], nz["g_noise"])
        zc = O.fill_up_uniform(zc_dgs, rays.cuda(), nz["u_fill"])
    dev_dgs = (zd - zc_dgs.sort(dim=-1).values).abs().max(dim=-1).values.cpu()
    dev_fill = (z - zc).abs().max(dim=-1).values.cpu()
    # (b) CPU goldens of the unmodified reference
    cpu_dgs = (zd.cpu() - g["z_depthguided"].sort(dim=-1).values).abs().max(dim=-1).values
    cpu_fill = (z.cpu() - g["z_filled"]).abs().max(dim=-1).values
    ref_split = _bad((zc.cpu() - g["z_filled"]).abs().max(dim=-1).values, Z_TOL)      # reference(cuda) != reference(cpu)
    print("%s: vs reference-on-cuda: rays exact %.4f (filled %.4f) | vs CPU golden: %.4f (filled %.4f) | rays where the "
          "reference's own backends differ: %.4f" % (name, (dev_dgs <= Z_TOL).float().mean(), (dev_fill <= Z_TOL).float().mean(),
                                                     (cpu_dgs <= Z_TOL).float().mean(), (cpu_fill <= Z_TOL).float().mean(),
                                                     ref_split.float().mean()))
    assert not bool(_bad(dev_dgs, Z_TOL).any()) and not bool(_bad(dev_fill, Z_TOL).any()), "differs from the reference on cuda"
    assert not bool((_bad(cpu_fill, Z_TOL) & ~ref_split).any()), "differs from the CPU golden on a ray where the reference agrees with itself"
    assert bool((z[..., 1:] >= z[..., :-1]).all()), "samples must be sorted ascending"


@pytest.mark.parametrize("name", ALL_CASES)
@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("parity", PAR_TOL)])
def test_query_stagewise(golden_dir, name, mode, tol):
    """PixelNeRF.forward on the reference's own sample positions (no RNG, no sampler decisions)."""
    g, cfg, batch, latent, mlp, rays, noise = _load(golden_dir, name)
    model = product_model(batch, latent, mlp, "cuda", mode)
    z = g["z_filled"]
    pts = (rays[..., None, :3] + z.unsqueeze(-1) * rays[..., None, 3:6]).reshape(cfg["SB"], -1, 3)
    vd = rays[..., None, 3:6].expand(-1, -1, cfg["K"], -1).reshape(cfg["SB"], -1, 3)
    with torch.no_grad():
        out = model(pts.cuda(), vd.cuda().contiguous()).cpu()
    ref = g["net_out"].reshape(out.shape)
    assert bool(torch.isfinite(out).all())
    err_rgb = (out[..., :3] - ref[..., :3]).abs().max()
    rel_sig = ((out[..., 3] - ref[..., 3]).abs() / (1.0 + ref[..., 3].abs())).max()
    print("%s/%s: max|d rgb| %.3g  max rel|d sigma| %.3g" % (name, mode, err_rgb, rel_sig))
    # sigma enters the image only through alpha = 1 - exp(-delta * sigma) with delta ~ 1e-2: relative 1e-4 on it is far inside the
    # bar on the rendered values, which test_composite_stagewise / test_render_end_to_end assert directly
    assert err_rgb <= tol and rel_sig <= (1e-4 if mode == "parity" else tol)


@pytest.mark.parametrize("name", ALL_CASES)
@pytest.mark.parametrize("mode", ["fp32", "parity"])
def test_composite_stagewise(golden_dir, name, mode):
    g, cfg, batch, latent, mlp, rays, noise = _load(golden_dir, name)
    model = product_model(batch, latent, mlp, "cuda", mode)
    rend = renderer_for(cfg)
    with torch.no_grad():
        w, rgb, depth = rend.composite(model, rays.cuda(), g["z_filled"].cuda())
    e_rgb = (rgb.cpu() - g["rgb"]).abs().max()
    e_d = (depth.cpu() - g["depth"]).abs().max()
    e_w = (w.cpu() - g["weights"]).abs().max()
    print("%s/%s: max|d rgb| %.3g |d depth| %.3g |d w| %.3g" % (name, mode, e_rgb, e_d, e_w))
    assert bool(torch.isfinite(rgb).all()) and bool(torch.isfinite(depth).all()) and bool(torch.isfinite(w).all())
    assert e_rgb <= PAR_TOL and e_d <= PAR_TOL and e_w <= PAR_TOL


@pytest.mark.parametrize("name", ALL_CASES)
@pytest.mark.parametrize("mode", ["fp32", "parity"])
def test_render_end_to_end(golden_dir, name, mode):
    """NeRFRendererDGS.forward with the reference's noise injected: (a) vs the reference algorithm on cuda: 1e-4 on EVERY ray
    (rgb, depth and the compositing weights); (b) vs the CPU goldens of the unmodified reference: 1e-4 on every ray on which
    the reference-on-cuda agrees with the reference-on-cpu."""
    g, cfg, batch, latent, mlp, rays, noise = _load(golden_dir, name)
    model = product_model(batch, latent, mlp, "cuda", mode)
    rend = renderer_for(cfg, noise)
    with torch.no_grad():
        out = rend(model, rays.cuda(), want_weights=True)
    assert out.fine.weights.shape == g["weights"].shape
    sc = oracle_scene_on(O.make_scene_state(batch, latent, mlp), "cuda")
    nz = rend.noise
    with torch.no_grad():
        rgb_c, depth_c, w_c = O.render(sc, rays.cuda(), cfg["K"], cfg["C"], cfg["G"], cfg["white"], nz["u_coarse"], nz["g_noise"], nz["u_fill"])
    e_dev = ray_err(out.fine.rgb, out.fine.depth, rgb_c, depth_c).cpu()
    e_w = (out.fine.weights - w_c).abs().max(dim=-1).values.cpu()
    e_cpu = ray_err(out.fine.rgb.cpu(), out.fine.depth.cpu(), g["rgb"], g["depth"])
    ref_split = _bad(ray_err(rgb_c.cpu(), depth_c.cpu(), g["rgb"], g["depth"]), TOL)
    print("%s/%s: vs reference-on-cuda max |err| %.3g (weights %.3g) | vs CPU golden: rays beyond 1e-4 %.4f, of which the "
          "reference's own backends differ on %.4f; max elsewhere %.3g" % (
              name, mode, e_dev.max(), e_w.max(), _bad(e_cpu, TOL).float().mean(), (_bad(e_cpu, TOL) & ref_split).float().mean(),
              e_cpu[~ref_split].max() if bool((~ref_split).any()) else 0.0))
    assert not bool(_bad(e_dev, TOL).any()) and not bool(_bad(e_w, TOL).any())
    assert not bool((_bad(e_cpu, TOL) & ~ref_split).any())


def test_fast_mode_psnr(golden_dir):
    g, cfg, batch, latent, mlp, rays, noise = _load(golden_dir, "cfg2_dtu64")
    model = product_model(batch, latent, mlp, "cuda", "fast")
    rend = renderer_for(cfg)
    with torch.no_grad():
        w, rgb, depth = rend.composite(model, rays.cuda(), g["z_filled"].cuda())
    p = O.psnr(rgb.cpu(), g["rgb"])
    print("fast mode PSNR vs reference render: %.1f dB, max|d rgb| %.3g" % (p, (rgb.cpu() - g["rgb"]).abs().max()))
    assert p > 40.0


def test_oracle_live_vs_cuda_fresh_seed():
    """Not a fixture: CPU oracle, oracle-on-cuda and the CUDA path on a fresh seeded case (same contract as
    test_render_end_to_end)."""
    cfg = dict(H=32, W=32, NV=4, SB=1, near=1.0, far=2.5, K=24, C=300, G=9, white=True, nr=64, seed=21)
    batch, latent, mlp, rays, noise = MG.case_inputs(cfg)
    scene = O.make_scene_state(batch, latent, mlp)
    rgb_o, depth_o, w_o = O.render(scene, rays, cfg["K"], cfg["C"], cfg["G"], cfg["white"],
                                   noise["u_coarse"], noise["g_noise"], noise["u_fill"])
    sc = oracle_scene_on(scene, "cuda")
    nz = {k: v.cuda() for k, v in noise.items()}
    with torch.no_grad():
        rgb_c, depth_c, _ = O.render(sc, rays.cuda(), cfg["K"], cfg["C"], cfg["G"], cfg["white"], nz["u_coarse"], nz["g_noise"], nz["u_fill"])
    ref_split = _bad(ray_err(rgb_c.cpu(), depth_c.cpu(), rgb_o, depth_o), TOL)
    for mode in ("fp32", "parity"):
        model = product_model(batch, latent, mlp, "cuda", mode)
        rend = renderer_for(cfg, noise)
        with torch.no_grad():
            out = rend(model, rays.cuda())
        assert not bool(_bad(ray_err(out.fine.rgb, out.fine.depth, rgb_c, depth_c), TOL).any())
        e = ray_err(out.fine.rgb.cpu(), out.fine.depth.cpu(), rgb_o, depth_o)
        assert not bool((_bad(e, TOL) & ~ref_split).any())


def test_properties_and_edges():
    """Size-independent properties + reference edge behaviour."""
    cfg = dict(H=32, W=32, NV=4, SB=1, near=1.0, far=2.5, K=32, C=200, G=12, white=True, nr=128, seed=5)
    batch, latent, mlp, rays, noise = MG.case_inputs(cfg)
    model = product_model(batch, latent, mlp, "cuda", "fp32")
    rend = renderer_for(cfg, noise)
    rays = rays.cuda()
    with torch.no_grad():
        full = rend(model, rays, want_weights=True)
        # (1) rays are independent: rendering two halves == rendering everything (the multi-GPU contract)
        rend_a = renderer_for(cfg, {k: v[:, :64] for k, v in noise.items()})
        rend_b = renderer_for(cfg, {k: v[:, 64:] for k, v in noise.items()})
        a, b = rend_a(model, rays[:, :64].contiguous()), rend_b(model, rays[:, 64:].contiguous())
        assert torch.equal(torch.cat((a.fine.rgb, b.fine.rgb), 1), full.fine.rgb)
        assert torch.equal(torch.cat((a.fine.depth, b.fine.depth), 1), full.fine.depth)
        # (2) determinism
        again = rend(model, rays)
        assert torch.equal(again.fine.rgb, full.fine.rgb)
        # (3) white vs black background differ by exactly 1 - sum(weights) (nerf_renderer.py:357-360)
        rend.white_bkgd = False
        black = rend(model, rays)
        acc = full.fine.weights.sum(-1, keepdim=True)
        assert (full.fine.rgb - (black.fine.rgb + 1 - acc)).abs().max() <= 2e-6
        assert bool((full.fine.weights >= 0).all()) and float(acc.max()) <= 1.0 + 1e-5
        # (4) empty ray batch
        rend.noise = None
        empty = rend(model, rays[:, :0].contiguous())
        assert empty.fine.rgb.shape == (1, 0, 3)
        # (4b) counter-based noise is keyed by the logical ray index: shards rendered with their ray_offset (what the multi-GPU
        #      render does) reproduce the single-call image bit for bit, for any split
        rend.noise = dict(seed=99)
        whole = rend.render_packed(model, rays)
        parts = [rend.render_packed(model, rays[:, lo:hi].contiguous(), ray_offset=lo) for lo, hi in ((0, 40), (40, 41), (41, 128))]
        assert torch.equal(torch.cat(parts, 1), whole)
        o = rend(model, rays)
        assert torch.equal(whole[..., :3], o.fine.rgb) and torch.equal(whole[..., 3], o.fine.depth)
    # (5) loud failures instead of fallbacks
    with pytest.raises(RuntimeError):
        model.context().render(rays.cpu(), 32, 200, 12, True, 0)
    with pytest.raises(AssertionError):
        rend(model, rays[0])                      # reference asserts 3-D rays (nerf_renderer.py:412)
    with pytest.raises(RuntimeError):
        model.context().render(rays, 16, 200, 32, True, 0)   # n_gaussian > n_samples


def test_full_size_properties():
    """BASELINE.json configs[1] at full size (512x512, 4 views, 64 samples/ray, 262 144 rays): size-independent
    properties (determinism, shard-vs-whole equality, white/black background relation, weights in [0,1]) and, on a
    strided subset of 2048 rays, the tcgen05 parity mode against the fp32 CUDA-core mode AND against the oracle
    (torch-CPU, and eager torch on cuda) on the same sample depths."""
    import bench
    batch, latent, mlp, rays = bench.build_inputs()
    model = product_model(batch, latent, mlp, "cuda", "parity")
    cfg = dict(K=bench.K, C=bench.C, G=bench.G, white=False)
    rend = renderer_for(cfg)
    rend.noise = dict(seed=1234)
    rays = rays.cuda()
    n = rays.shape[1]
    with torch.no_grad():
        full = rend(model, rays)
        again = rend(model, rays)
        assert torch.equal(full.fine.rgb, again.fine.rgb) and torch.equal(full.fine.depth, again.fine.depth)
        assert bool(torch.isfinite(full.fine.rgb).all()) and bool(torch.isfinite(full.fine.depth).all())
        # a strided subset rendered on its own, with the noise of the same logical rays, must reproduce the image
        # (counter-based noise is keyed by ray index within the call, so compare through explicit sample depths)
        ctx = model.context()
        sub = torch.arange(0, n, 97, device="cuda")[:2048]
        r_sub = rays[:, sub].contiguous()
        z_sub = ctx.sample(r_sub, cfg["K"], cfg["C"], cfg["G"], dict(seed=7))
        assert bool((z_sub[..., 1:] >= z_sub[..., :-1]).all())
        w_p, rgb_p, d_p = ctx.composite(r_sub, z_sub, False, 1)
        w_f, rgb_f, d_f = ctx.composite(r_sub, z_sub, False, 0)           # fp32 CUDA-core arithmetic
        e = max(float((rgb_p - rgb_f).abs().max()), float((d_p - d_f).abs().max()))
        print("full-size: parity vs fp32 on 2048 strided rays: max |err| %.3g" % e)
        assert e <= PAR_TOL
        # ... and against the ORACLE on the same rays and sample depths: on the host cores (the CPU reference path) and on cuda
        scene = O.make_scene_state(batch, latent, mlp)
        _, rgb_o, d_o = O.composite(scene, r_sub.cpu(), z_sub.cpu(), False)
        sc = oracle_scene_on(scene, "cuda")
        _, rgb_c, d_c = O.composite(sc, r_sub, z_sub, False)
        e_o = float(ray_err(rgb_p.cpu(), d_p.cpu(), rgb_o, d_o).max())
        e_c = float(ray_err(rgb_p, d_p, rgb_c, d_c).max())
        print("full-size: parity vs CPU oracle %.3g, vs oracle-on-cuda %.3g (2048 strided rays, 512x512 workload, 320x320 latent); "
              "PSNR vs CPU oracle %.1f dB" % (e_o, e_c, O.psnr(rgb_p.cpu(), rgb_o)))
        assert e_o <= PAR_TOL and e_c <= PAR_TOL
        del sc
        assert float(w_p.min()) >= 0.0 and float(w_p.sum(-1).max()) <= 1.0 + 1e-5
        w_w, rgb_w, _ = ctx.composite(r_sub, z_sub, True, 1)
        assert (rgb_w - (rgb_p + 1 - w_p.sum(-1, keepdim=True))).abs().max() <= 2e-6
        # two halves of the image == the whole image for the MLP + compositing stages (ray independence)
        z_all = ctx.sample(rays, cfg["K"], cfg["C"], cfg["G"], dict(seed=7))
        _, rgb_all, d_all = ctx.composite(rays, z_all, False, 1, want_weights=False)
        h = n // 2
        _, rgb_a, d_a = ctx.composite(rays[:, :h].contiguous(), z_all[:, :h].contiguous(), False, 1, want_weights=False)
        _, rgb_b, d_b = ctx.composite(rays[:, h:].contiguous(), z_all[:, h:].contiguous(), False, 1, want_weights=False)
        assert torch.equal(torch.cat((rgb_a, rgb_b), 1), rgb_all) and torch.equal(torch.cat((d_a, d_b), 1), d_all)
        fg = float((full.fine.depth > 0).float().mean())
        print("full-size: %.1f %% of rays hit geometry, rgb mean %.3f" % (100 * fg, float(full.fine.rgb.mean())))
        assert fg > 0.05


def test_whole_image_entry_matches_ray_batches():
    """diner_render_image (in-library gen_rays + whole-image launch) against the reference flow:
    gen_rays (cam_geometry.py:5-48, restated in diner_b200.synthetic) + renderer.forward per ray batch of 4096 + cat."""
    from diner_b200 import synthetic as S
    from diner_b200.predict import predict_imgs_from_batch
    cfg = dict(H=64, W=96, NV=4, SB=2, near=1.0, far=2.5, K=32, C=200, G=12, white=True, nr=16, seed=3)
    batch, latent, mlp, _, _ = MG.case_inputs(cfg)
    model = product_model(batch, latent, mlp, "cuda", "parity")
    rend = renderer_for(cfg)
    rend.noise = dict(seed=77)
    ext, intr = batch["target_extrinsics"].cuda(), batch["target_intrinsics"].cuda()
    SB, H, W = cfg["SB"], cfg["H"], cfg["W"]
    rays_ref = S.gen_rays(batch["target_extrinsics"], batch["target_intrinsics"], W, H, torch.full((SB,), cfg["near"]),
                          torch.full((SB,), cfg["far"])).view(SB, H * W, 8)
    rays_dev = model.context().gen_rays(ext, intr, H, W, cfg["near"], cfg["far"])
    err = (rays_dev.cpu() - rays_ref).abs().max()
    exact = (rays_dev.cpu() == rays_ref).float().mean()
    print("gen_rays: max |err| %.3g, bit-exact fraction %.4f" % (err, exact))
    assert err <= 2.5e-7 and exact >= 0.5       # 1-ulp differences in the K=3 products of a minority of the rays
    with torch.no_grad():
        whole = rend(model, rays_dev)                                   # one call over all rays, same seed
        b = dict(batch, target_rgb=torch.empty(SB, 3, H, W))
        b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
        rgb, depth = predict_imgs_from_batch(model, rend, b, cfg["near"], cfg["far"], return_depth=True, encode=False)
    assert rgb.shape == (SB, 3, H, W) and depth.shape == (SB, 1, H, W)
    assert torch.equal(rgb.permute(0, 2, 3, 1).reshape(SB, H * W, 3), whole.fine.rgb)
    assert torch.equal(depth.reshape(SB, H * W), whole.fine.depth)
    # and against the CPU oracle on the reference's rays (stage-wise: same sample depths)
    ctx = model.context()
    z = ctx.sample(rays_dev, cfg["K"], cfg["C"], cfg["G"], dict(seed=77))
    scene = O.make_scene_state(batch, latent, mlp)
    sub = slice(0, 512)
    _, rgb_o, depth_o = O.composite(scene, rays_ref[:, sub], z[:, sub].cpu(), cfg["white"])
    _, rgb_p, depth_p = ctx.composite(rays_dev[:, sub].contiguous(), z[:, sub].contiguous(), cfg["white"], 1)
    e = max(float((rgb_p.cpu() - rgb_o).abs().max()), float((depth_p.cpu() - depth_o).abs().max()))
    print("whole-image entry: composite vs oracle on 512 rays/scene: max |err| %.3g" % e)
    assert e <= PAR_TOL


@pytest.mark.parametrize("H,W,seed", [(64, 64, 1), (48, 64, 3), (96, 128, 5)])
def test_depth2normal_kernel_matches_oracle(H, W, seed):
    """diner_depth2normal (scene prepare) against the CPU oracle's restatement of
    src/util/depth2normal.py:6-87 -- bit-exact (the op order was pinned by emulation), NaN pattern included."""
    from diner_b200 import synthetic as S
    from diner_b200.capi import Context
    batch = S.make_scene(H, W, 4, 2, 1.0, 2.5, seed)
    dm = batch["src_depths"].flatten(end_dim=1)
    K = batch["src_intrinsics"].flatten(end_dim=1)
    ref = O.depth2normal(dm, K)
    out = Context("cuda").depth2normal(dm.cuda().contiguous(), K.cuda().contiguous()).cpu()
    same = (out == ref) | (torch.isnan(out) & torch.isnan(ref))
    print("depth2normal %dx%d: bit-exact fraction %.5f, max |err| %.3g" % (
        H, W, same.float().mean(), torch.nan_to_num(out - ref).abs().max()))
    assert bool(same.all())


@pytest.mark.parametrize("name,cfg", [
    # BASELINE.json configs[2] shape (Facescape: 256x256, 4 views, K=128, SB=4, white background), forward only
    ("cfg3_face256", dict(H=256, W=256, NV=4, SB=4, near=1.0, far=2.5, K=128, C=1000, G=48, white=True, nr=384, seed=11)),
    # configs[4] shape (stress: 8 views, K=256) on a 256x256 scene
    ("cfg5_nv8_k256", dict(H=256, W=256, NV=8, SB=1, near=0.3211, far=1.2041, K=256, C=1000, G=96, white=False, nr=256, seed=12)),
])
def test_other_config_shapes_forward(name, cfg):
    """The other BASELINE.json configs as parity-test cases: tcgen05 parity mode against the fp32 CUDA-core
    mode (reference arithmetic, itself pinned to the goldens) at those shapes, plus the CPU oracle on a few rays."""
    from diner_b200 import synthetic as S
    H, W, SB, NV = cfg["H"], cfg["W"], cfg["SB"], cfg["NV"]
    batch = S.make_scene(H, W, NV, SB, cfg["near"], cfg["far"], cfg["seed"])
    gen = torch.Generator().manual_seed(cfg["seed"])         # (the hash-based fixture latent is too slow to build at this size)
    latent = torch.randn(SB, NV, 512, (H + 128) // 2, (W + 128) // 2, generator=gen) * 0.5
    mlp = S.make_mlp_state(seed=cfg["seed"])
    rays = S.gen_rays(batch["target_extrinsics"], batch["target_intrinsics"], W, H, torch.full((SB,), cfg["near"]),
                      torch.full((SB,), cfg["far"])).view(SB, H * W, 8)
    rays = rays[:, (torch.arange(cfg["nr"]) * (H * W // cfg["nr"]) + 7) % (H * W)].contiguous()
    model = product_model(batch, latent, mlp, "cuda", "parity")
    ctx = model.context()
    r = rays.cuda()
    nz = dict(seed=5)
    z = ctx.sample(r, cfg["K"], cfg["C"], cfg["G"], nz)
    assert bool((z[..., 1:] >= z[..., :-1]).all())
    w_p, rgb_p, d_p = ctx.composite(r, z, cfg["white"], 1)
    w_f, rgb_f, d_f = ctx.composite(r, z, cfg["white"], 0)
    e = max(float((rgb_p - rgb_f).abs().max()), float((d_p - d_f).abs().max()))
    scene = O.make_scene_state(batch, latent, mlp)
    sub = slice(0, 24)
    _, rgb_o, d_o = O.composite(scene, rays[:, sub], z[:, sub].cpu(), cfg["white"])
    eo = max(float((rgb_p[:, sub].cpu() - rgb_o).abs().max()), float((d_p[:, sub].cpu() - d_o).abs().max()))
    print("%s: parity vs fp32 max |err| %.3g; parity vs CPU oracle (24 rays/scene) %.3g" % (name, e, eo))
    assert e <= PAR_TOL and eo <= PAR_TOL
    assert float(w_p.min()) >= 0.0 and float(w_p.sum(-1).max()) <= 1.0 + 1e-5


@pytest.mark.parametrize("NV,SB,K,nr", [(1, 1, 8, 1), (1, 2, 16, 37), (2, 1, 40, 300), (4, 1, 64, 1000), (4, 2, 24, 2500),
                                        (8, 1, 16, 1200), (8, 1, 64, 3000), (2, 2, 64, 4100), (4, 1, 128, 4096),
                                        (3, 1, 20, 700), (6, 2, 16, 900),       # view counts that are not powers of two: padded rows
                                        (1, 1, 64, 1500)])                      # one view, many rounds per CTA: every tile is a cold one
def test_pair_kernel_pipeline_shape_sweep(NV, SB, K, nr):
    """The software pipeline of the CTA-pair kernel (per-K-block release barriers, cross-tile hand-off, helper warps)
    over many tile counts -- 1 tile to >10 rounds per CTA, live and padded last rounds, 1..8 views: parity mode must agree
    with the fp32 CUDA-core mode (1e-4) and be deterministic.  A protocol bug shows up here as a watchdog error."""
    from diner_b200 import synthetic as S
    H = W = 32
    batch = S.make_scene(H, W, NV, SB, 1.0, 2.5, 100 + NV)
    latent = torch.randn(SB, NV, 512, (H + 128) // 2, (W + 128) // 2, generator=torch.Generator().manual_seed(100 + NV)) * 0.5
    mlp = S.make_mlp_state(seed=100 + NV)
    rays = S.gen_rays(batch["target_extrinsics"], batch["target_intrinsics"], W, H, torch.full((SB,), 1.0),
                      torch.full((SB,), 2.5)).view(SB, H * W, 8)
    rays = rays[:, torch.arange(nr) % (H * W)].contiguous().cuda()
    model = product_model(batch, latent, mlp, "cuda", "parity")
    ctx = model.context()
    z = ctx.sample(rays, K, 200, min(K // 3, 12), dict(seed=nr))
    _, rgb_p, d_p = ctx.composite(rays, z, True, 1, want_weights=False)
    _, rgb_q, d_q = ctx.composite(rays, z, True, 1, want_weights=False)
    assert torch.equal(rgb_p, rgb_q) and torch.equal(d_p, d_q)
    _, rgb_f, d_f = ctx.composite(rays, z, True, 0, want_weights=False)
    e_rgb, e_d = float((rgb_p - rgb_f).abs().max()), float((d_p - d_f).abs().max())
    print("NV=%d SB=%d K=%d rays=%d: parity vs fp32 max |err| rgb %.3g depth %.3g" % (NV, SB, K, nr, e_rgb, e_d))
    # (round 1 measured 1.09e-4 on depth at NV=4 / SB=2 / K=24 with bf16 hi/lo operands -- few, widely spaced samples amplify
    # the pre-activation error; the fp16 hi/lo split keeps 22 significant bits and is held to PAR_TOL here like everywhere)
    # These random-weight cases with as few as 8 widely spaced samples per ray are the worst conditioned ones of the suite (depth
    # is a sigma-weighted sum over samples up to 1.5 apart): they are held to the north_star bar itself; measured <= 6.4e-5.
    assert bool(torch.isfinite(rgb_p).all()) and bool(torch.isfinite(d_p).all())
    assert e_rgb <= PAR_TOL and e_d <= TOL
    _, rgb_s, _ = ctx.composite(rays, z, True, 2, want_weights=False)       # fast mode runs the same protocol with other timings
    assert bool(torch.isfinite(rgb_s).all()) and float((rgb_s - rgb_f).abs().max()) < 1e-2
    ctx.set_option("fused", 0)                                              # PRE / POST launches with the HBM scratch: same arithmetic
    _, rgb_u, d_u = ctx.composite(rays, z, True, 1, want_weights=False)
    assert torch.equal(rgb_u, rgb_p) and torch.equal(d_u, d_p), "fused and two-kernel paths differ"
    ctx.set_option("fused", 1)
    for pts in (2, 4):                                                      # rounds of 2 / 4 POST tiles (default 1)
        ctx.set_option("post_tiles", pts)
        _, rgb_u, d_u = ctx.composite(rays, z, True, 1, want_weights=False)
        assert torch.equal(rgb_u, rgb_p) and torch.equal(d_u, d_p), "post_tiles=%d changes the result" % pts
    ctx.set_option("post_tiles", 1)
    for tail in (0, 4):                                                     # the other MMA issue orders must give the same bits
        ctx.set_option("tail_kb", tail)
        _, rgb_t, d_t = ctx.composite(rays, z, True, 1, want_weights=False)
        assert torch.equal(rgb_t, rgb_p) and torch.equal(d_t, d_p), "tail_kb=%d changes the result" % tail
    ctx.set_option("tail_kb", 3)
    for early, fused, warm in ((1, 1, 0), (1, 0, 0), (0, 0, 0), (0, 1, 0), (0, 1, 1), (1, 1, 1)):
        ctx.set_option("early_lin", early)                                  # next tile's lin_in behind the last fc_1 (TMEM half ping-pong)
        ctx.set_option("fused", fused)
        ctx.set_option("warm_rounds", warm)                                 # next round's first PRE tile prepared under the POST tile
        for pts in ((1, 2) if warm else (1,)):
            ctx.set_option("post_tiles", pts)
            _, rgb_t, d_t = ctx.composite(rays, z, True, 1, want_weights=False)
            assert torch.equal(rgb_t, rgb_p) and torch.equal(d_t, d_p), \
                "early_lin=%d fused=%d warm_rounds=%d post_tiles=%d changes the result" % (early, fused, warm, pts)
        ctx.set_option("post_tiles", 1)
    _, rgb_s2, _ = ctx.composite(rays, z, True, 2, want_weights=False)      # fast mode through the warm-round protocol
    assert torch.equal(rgb_s2, rgb_s)


@pytest.mark.parametrize("backward_tc", [1, 0], ids=["tcgen05", "fp32_cuda_cores"])
def test_backward_matches_reference_gradients(golden_dir, backward_tc):
    """BASELINE config 3: diner_render_backward against loss / gradient digests of the UNMODIFIED reference
    (tests/golden/grads_cfg1_face64.pt, autograd through composite / PixelNeRF.forward / ResnetFC with the MSE of diner.py:266),
    on both arithmetic paths of the backward: tcgen05 GEMMs (gemm_tc3.cu, default) and fp32 CUDA cores."""
    from diner_b200.nerf_renderer import mlp_param_order
    g = torch.load(os.path.join(golden_dir, "grads_cfg1_face64.pt"))
    cfg, batch, latent, mlp, rays, noise, gt = MG.grad_case_inputs()
    model = product_model(batch, latent, mlp, "cuda", "fp32")
    ctx = model.context()
    ctx.set_option("backward_tc", backward_tc)
    r, z = rays.cuda(), g["z"].cuda().contiguous()
    _, rgb, _ = ctx.composite(r, z, cfg["white"], 0, want_weights=False)
    loss = float(((rgb.cpu() - gt) ** 2).mean())
    assert abs(loss - g["loss"]) <= 1e-5
    g_rgb = (2.0 * (rgb - gt.cuda()) / rgb.numel()).contiguous()          # d MSE(mean) / d rgb
    gp, dl = ctx.render_backward(r, z, cfg["white"], g_rgb, None, True, tuple(latent.shape))
    named = dict(model.mlp_fine.named_parameters())
    off = 0

    def check(name, t, ref, n):
        d = O.grad_digest(t.cpu(), n)
        assert d["numel"] == ref["numel"], name
        assert abs(d["norm"] - ref["norm"]) <= 2e-3 * ref["norm"], (name, d["norm"], ref["norm"])
        scale = float(ref["sample"].abs().max().clamp_min(1e-12))
        assert float((d["sample"] - ref["sample"]).abs().max()) / scale <= 5e-3, name

    for k in mlp_param_order(model.mlp_fine):
        n = named[k].numel()
        check(k, gp[off:off + n], g["grads"][k], 256)
        off += n
    assert off == gp.numel()
    check("latent", dl, g["latent_grad"], 4096)


@pytest.mark.parametrize("cfg", [
    dict(H=32, W=48, NV=2, SB=2, near=1.0, far=2.5, K=12, C=100, G=4, white=False, nr=40, seed=32),
    dict(H=32, W=32, NV=8, SB=1, near=1.0, far=2.5, K=16, C=100, G=5, white=True, nr=64, seed=33),
], ids=["sb2_nv2_black", "nv8_white"])
@pytest.mark.parametrize("backward_tc", [1, 0], ids=["tcgen05", "fp32_cuda_cores"])
def test_backward_other_shapes_vs_oracle_autograd(cfg, backward_tc):
    """diner_render_backward (with a depth term as well) against torch autograd through the CPU oracle (pinned to the reference's
    gradients by tests/test_oracle.py) on shapes the golden does not cover: SB > 1, NV != 4, black background.  The upstream
    gradients are random-signed per ray, so every parameter gradient is a heavily cancelling sum: the worst case for the
    tcgen05 path, whose accumulator truncates (see the module docstring).  Bars: every element within 0.5 % (fp32 CUDA cores) /
    2 % (tcgen05; measured 0.63 %) of the largest reference entry of its tensor, and 1e-3 relative in the Frobenius norm."""
    from diner_b200 import synthetic as S
    from diner_b200.nerf_renderer import mlp_param_order
    batch, latent, mlp, rays, noise = MG.case_inputs(cfg)
    scene = O.make_scene_state(batch, latent, mlp)
    z = O.fill_up_uniform(O.sample_depthguided(scene, rays, cfg["K"], cfg["C"], cfg["G"], noise["u_coarse"], noise["g_noise"]),
                          rays, noise["u_fill"])
    g_rgb = S.hash_normal((cfg["SB"], rays.shape[1], 3), cfg["seed"], 951) * 0.01
    g_dep = S.hash_normal((cfg["SB"], rays.shape[1]), cfg["seed"], 952) * 0.01
    # reference gradients of L = sum(g_rgb * rgb) + sum(g_dep * depth) by autograd through the oracle
    leaf_mlp = {k: v.clone().requires_grad_(True) for k, v in mlp.items()}
    leaf_lat = latent.clone().requires_grad_(True)
    scene.mlp, scene.latent = leaf_mlp, leaf_lat
    _, rgb, depth = O.composite(scene, rays, z, cfg["white"])
    ((rgb * g_rgb).sum() + (depth * g_dep).sum()).backward()
    model = product_model(batch, latent, mlp, "cuda", "fp32")
    model.context().set_option("backward_tc", backward_tc)
    gp, dl = model.context().render_backward(rays.cuda(), z.cuda().contiguous(), cfg["white"], g_rgb.cuda().contiguous(),
                                             g_dep.cuda().contiguous(), True, tuple(latent.shape))
    assert bool(torch.isfinite(gp).all()) and bool(torch.isfinite(dl).all())
    off, worst, worst_f = 0, ("", 0.0), ("", 0.0)
    items = []
    for k in mlp_param_order(model.mlp_fine):
        ref = leaf_mlp[k].grad
        items.append((k, gp[off:off + ref.numel()].view(ref.shape).cpu(), ref))
        off += ref.numel()
    assert off == gp.numel()
    items.append(("latent", dl.cpu(), leaf_lat.grad))
    for k, got, ref in items:
        scale = float(ref.abs().max().clamp_min(1e-12))
        e = float((got - ref).abs().max()) / scale
        f = float((got - ref).double().norm() / ref.double().norm().clamp_min(1e-30))
        worst = max(worst, (k, e), key=lambda x: x[1])
        worst_f = max(worst_f, (k, f), key=lambda x: x[1])
    print("backward (%s): worst element error / max|ref| %.3g (%s); worst Frobenius-relative error %.3g (%s)" % (
        "tcgen05" if backward_tc else "fp32 CUDA cores", worst[1], worst[0], worst_f[1], worst_f[0]))
    assert worst[1] <= (2e-2 if backward_tc else 5e-3) and worst_f[1] <= (2e-3 if backward_tc else 1e-3)


def test_training_step_through_module_api(golden_dir):
    """The module-level path a training step takes (diner.py:257-266): NeRFRendererDGS.forward with grad enabled ->
    autograd Function -> loss.backward() fills .grad of the ResnetFC parameters and of encoder.latent."""
    g = torch.load(os.path.join(golden_dir, "grads_cfg1_face64.pt"))
    cfg, batch, latent, mlp, rays, noise, gt = MG.grad_case_inputs()
    model = product_model(batch, latent, mlp, "cuda", "fp32").train()
    model.encoder.latent = model.encoder.latent.detach().clone().requires_grad_(True)
    model.encoder.scene_version += 1
    rend = renderer_for(cfg, noise)
    out = rend(model, rays.cuda())
    loss = torch.nn.functional.mse_loss(out.fine.rgb, gt.cuda())
    loss.backward()
    # same noise -> same sample depths as the golden (well-conditioned rays; a flipped shortlist entry would only move the
    # loss slightly), so the loss and the gradient norms must be close to the reference's
    assert abs(float(loss) - g["loss"]) <= 2e-3
    for k, p in model.mlp_fine.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
        ref = g["grads"][k]["norm"]
        assert abs(float(p.grad.double().norm()) - ref) <= 0.05 * ref + 1e-9, (k, float(p.grad.norm()), ref)
    lg = model.encoder.latent.grad
    assert lg is not None and abs(float(lg.double().norm()) - g["latent_grad"]["norm"]) <= 0.05 * g["latent_grad"]["norm"]


def test_training_step_at_config3_size():
    """BASELINE.json configs[2] at full size (Facescape-shaped: SB = 4 scenes of 256 x 256, 4 views, 128 samples/ray, 4 096 rays per
    scene = 2.1 M samples per step): loss + backward through the module API on both backward paths; the tcgen05 gradients must
    agree with the fp32 CUDA-core ones (2e-3 in the Frobenius norm per tensor) -- the chunked chain, the split-K weight gradients
    and the loss scaling at the size the small reference-gradient tests do not reach."""
    from diner_b200 import synthetic as S
    from diner_b200.predict import calc_losses
    from diner_b200.nerf_renderer import NeRFRendererDGS
    Ht = Wt = 256
    SBt, NVt, Kt, RB = 4, 4, 128, 4096
    batch = S.make_scene(Ht, Wt, NVt, SBt, 1.0, 2.5, 0)
    gen = torch.Generator().manual_seed(0)
    latent = torch.randn(SBt, NVt, 512, (Ht + 128) // 2, (Wt + 128) // 2, generator=gen) * 0.5
    model = product_model(batch, latent, S.make_mlp_state(seed=0), "cuda", "parity").train()
    model.encoder.latent = model.encoder.latent.detach().clone().requires_grad_(True)
    model.encoder.scene_version += 1
    rend = NeRFRendererDGS(n_samples=Kt, n_depth_candidates=1000, n_gaussian=int(15 * Kt / 40), white_bkgd=True)
    rend.noise = dict(seed=17)
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    b["target_rgb"] = torch.rand(SBt, 3, Ht, Wt, generator=gen).cuda()
    grads = {}
    for path in (1, 0):
        model.context().set_option("backward_tc", path)
        model.zero_grad(set_to_none=True)
        model.encoder.latent.grad = None
        loss = calc_losses(model, rend, b, 1.0, 2.5, RB, generator=torch.Generator().manual_seed(5), encode=False)["total"]
        loss.backward()
        grads[path] = {k: p.grad.detach().clone() for k, p in model.mlp_fine.named_parameters()}
        grads[path]["latent"] = model.encoder.latent.grad.detach().clone()
        assert all(bool(torch.isfinite(g).all()) for g in grads[path].values()) and 0.0 < float(loss) < 1.0
    rel = sorted(((float((grads[1][k] - grads[0][k]).double().norm() / grads[0][k].double().norm().clamp_min(1e-30)), k) for k in grads[0]),
                 reverse=True)
    print("config-3 size: tcgen05 vs fp32 backward, Frobenius-relative difference per tensor, largest first: " +
          ", ".join("%s %.2g" % (k, v) for v, k in rel[:4]))
    # The two paths also differ in the FORWARD arithmetic they recompute with, so the ReLU masks of pre-activations within ~1e-5 of
    # zero flip between them (about one mask element in 1e5; a flipped element changes its row's downstream gradient by a few
    # per cent): every tensor differs by 1e-3 .. 4e-3 in the Frobenius norm -- measured: latent 3.9e-3, lin_z weights 2.3e-3, the
    # others below.  The same holds between any two arithmetic paths through a ReLU network (torch fp32 on CPU vs on cuda);
    # agreement with the reference's own gradients is what test_backward_matches_reference_gradients / bench's grad_check pin
    # (1.1e-5 on every gradient norm).
    assert all(v <= (1e-2 if k == "latent" else 5e-3) for v, k in rel)


def test_training_steps_reduce_the_loss():
    """A few Adam steps (diner.py:333) through calc_losses on a fixed tiny scene must reduce the MSE."""
    from diner_b200 import synthetic as S
    from diner_b200.predict import calc_losses
    cfg = dict(H=32, W=32, NV=4, SB=1, near=1.0, far=2.5, K=16, C=100, G=6, white=True, nr=16, seed=41)
    batch, latent, mlp, _, _ = MG.case_inputs(cfg)
    model = product_model(batch, latent, mlp, "cuda", "fp32").train()
    rend = renderer_for(cfg)
    rend.noise = dict(seed=3)
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    b["target_rgb"] = S.hash_uniform((1, 3, 32, 32), 41, 960).cuda()
    opt = torch.optim.Adam(model.mlp_fine.parameters(), lr=1e-4)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        out = calc_losses(model, rend, b, cfg["near"], cfg["far"], 256, generator=torch.Generator().manual_seed(0), encode=False)
        out["total"].backward()
        opt.step()
        losses.append(float(out["total"]))
    print("training losses:", ["%.5f" % v for v in losses])
    assert losses[-1] < losses[0]


def test_tile_order_of_rays_does_not_change_the_image():
    """diner_set_option("ray_image_width"): the fused launch walks a row-major image in 16 x 16 pixel tiles (L2 locality of the
    gathered feature-map lines).  Only the order of the work may change: bit-identical output, also for several scenes per call
    and for K that is neither a multiple nor a divisor of the 64-sample round."""
    from diner_b200 import synthetic as S
    for SB, Hh, Ww, K in ((1, 32, 48, 64), (2, 16, 32, 24), (1, 32, 32, 160)):
        batch = S.make_scene(Hh, Ww, 4, SB, 1.0, 2.5, 61)
        latent = torch.randn(SB, 4, 512, (Hh + 128) // 2, (Ww + 128) // 2, generator=torch.Generator().manual_seed(61)) * 0.5
        model = product_model(batch, latent, S.make_mlp_state(seed=61), "cuda", "parity")
        rays = S.gen_rays(batch["target_extrinsics"], batch["target_intrinsics"], Ww, Hh, torch.full((SB,), 1.0),
                          torch.full((SB,), 2.5)).view(SB, Hh * Ww, 8).contiguous().cuda()
        ctx = model.context()
        a = ctx.render(rays, K, 100, 8, True, 1, dict(seed=3), want_weights=True)
        ctx.set_option("ray_image_width", Ww)
        b = ctx.render(rays, K, 100, 8, True, 1, dict(seed=3), want_weights=True)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), (SB, Hh, Ww, K)
        ctx.set_option("ray_image_width", Ww - 16)                     # does not divide the ray count: silently the caller's order
        c = ctx.render(rays, K, 100, 8, True, 1, dict(seed=3))
        assert torch.equal(a[0], c[0])


def test_render_host_entry_matches_device_entry():
    """diner_render_host (plain host buffers in and out, copies and the stream sync inside the call -- the entry a non-torch
    caller binds) must return exactly what diner_render returns for the same rays and seed."""
    cfg = dict(H=32, W=32, NV=4, SB=2, near=1.0, far=2.5, K=32, C=200, G=12, white=True, nr=300, seed=8)
    batch, latent, mlp, rays, _ = MG.case_inputs(cfg)
    model = product_model(batch, latent, mlp, "cuda", "parity")
    ctx = model.context()
    SB, NR = rays.shape[:2]
    rays_h = rays.contiguous()                                    # ordinary pageable host memory
    rgb_h, dep_h = torch.full((SB, NR, 3), float("nan")), torch.full((SB, NR), float("nan"))
    ctx.render_host(rays_h, cfg["K"], cfg["C"], cfg["G"], cfg["white"], 1, 4242, rgb_h, dep_h)
    rgb_d, dep_d, _, _ = ctx.render(rays.cuda(), cfg["K"], cfg["C"], cfg["G"], cfg["white"], 1, dict(seed=4242))
    assert torch.equal(rgb_h, rgb_d.cpu()) and torch.equal(dep_h, dep_d.cpu())
    assert bool(torch.isfinite(rgb_h).all())


def test_mismatched_scene_and_mlp_are_rejected_before_any_launch():
    """latent channel count / positional-code width that do not match the MLP must fail with DINER_E_INVALID in
    every mode (the fp32 path used to size its workspace from d_latent and write with the scene's stride)."""
    from diner_b200 import synthetic as S
    cfg = dict(H=32, W=32, NV=2, SB=1, near=1.0, far=2.5, K=8, C=50, G=2, white=True, nr=8, seed=2)
    batch, latent, mlp, rays, _ = MG.case_inputs(cfg)
    model = product_model(batch, latent, mlp, "cuda", "fp32")
    enc = model.encoder
    enc.set_scene(torch.cat((enc.latent, enc.latent), dim=2).contiguous(), enc.depths, enc.depths_std, enc.normals)   # 1024 channels
    for mode in ("fp32", "parity", "fast"):
        model.mode = mode
        with pytest.raises(RuntimeError, match="latent channels"):
            with torch.no_grad():
                renderer_for(cfg)(model, rays.cuda())


def test_softplus_network_and_depth_diff_max(golden_dir):
    """Reference arguments off the shipped configs: ResnetFC(beta > 0) -> Softplus activations (resnetfc.py:124-127), served by
    the fp32 kernels whatever `mode` says; sample_depthguided(depth_diff_max != 0.05) (nerf_renderer.py:66,121)."""
    g, cfg, batch, latent, mlp, rays, noise = _load(golden_dir, "sb2_nv2_rect")
    model = product_model(batch, latent, mlp, "cuda", "parity")
    model.mlp_fine.beta = 1.5
    scene = O.make_scene_state(batch, latent, mlp)
    scene.beta = 1.5
    rend = renderer_for(cfg)
    with torch.no_grad():
        w, rgb, depth = rend.composite(model, rays.cuda(), g["z_filled"].cuda())
        w_o, rgb_o, depth_o = O.composite(scene, rays, g["z_filled"], cfg["white"])
    e = float(ray_err(rgb.cpu(), depth.cpu(), rgb_o, depth_o).max())
    print("softplus(beta=1.5) network, fp32 kernels vs oracle: max |err| %.3g" % e)
    assert e <= 2e-5 and float((rgb_o - g["rgb"]).abs().max()) > 1e-3          # and it really is a different network
    with pytest.raises(RuntimeError, match="Softplus"):
        model.context().composite(rays.cuda(), g["z_filled"].cuda().contiguous(), cfg["white"], 1)
    # depth_diff_max through the reference's method signature
    model.mlp_fine.beta = 0.0
    nz = {k: v.cuda().contiguous() for k, v in noise.items()}
    rend = renderer_for(cfg, noise)
    sc = oracle_scene_on(O.make_scene_state(batch, latent, mlp), "cuda")
    for ddm in (0.01, 0.2):
        zd = rend.sample_depthguided(rays.cuda(), model, cfg["K"], cfg["C"], depth_diff_max=ddm)
        with torch.no_grad():
            zo = O.sample_depthguided(sc, rays.cuda(), cfg["K"], cfg["C"], cfg["G"], nz["u_coarse"], nz["g_noise"], depth_diff_max=ddm)
        d = (zd - zo.sort(dim=-1).values).abs().max(dim=-1).values
        assert not bool(_bad(d, Z_TOL).any()), "depth_diff_max=%g" % ddm
    z_default = rend.sample_depthguided(rays.cuda(), model, cfg["K"], cfg["C"])
    assert not torch.equal(z_default, zd)


def test_encode_path_borrows_channels_last_latent():
    """PixelNeRF.encode (ResNet-34 trunk on cuDNN, random weights) -> channels-last latent borrowed by the library (no re-layout
    pass) must render exactly what the reference-layout (NCHW, transposed once) hand-over renders."""
    from diner_b200 import synthetic as S
    cfg = dict(H=64, W=64, NV=3, SB=2, near=1.0, far=2.5, K=24, C=200, G=8, white=True, nr=500, seed=13)
    batch, latent, mlp, rays, _ = MG.case_inputs(cfg)
    torch.manual_seed(1)
    model = product_model(batch, latent, mlp, "cuda", "parity")
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    with torch.no_grad():
        model.encode(b["src_rgbs"], b["src_depths"], b["src_depth_stds"], b["src_extrinsics"], b["src_intrinsics"])
        lat = model.encoder.latent
        assert not lat.is_contiguous() and lat.permute(0, 1, 3, 4, 2).is_contiguous()
        rend = renderer_for(cfg)
        rend.noise = dict(seed=5)
        a = rend(model, rays.cuda())
        l0 = model.context().launch_count()
        model.encoder.set_scene(lat.contiguous(), model.encoder.depths, model.encoder.depths_std, model.encoder.normals)   # NCHW copy
        c = rend(model, rays.cuda())
    assert torch.equal(a.fine.rgb, c.fine.rgb) and torch.equal(a.fine.depth, c.fine.depth)
    assert bool(torch.isfinite(a.fine.rgb).all()) and float(a.fine.rgb.std()) > 0


def test_cam_sweep_loop_and_output_side(tmp_path):
    """Callers around the hot path: the camera-sweep render loop (diner.py:180-211) and the output side (torch_cmap + image
    files, diner.py:120-133) on the device."""
    from diner_b200 import io as IO
    from diner_b200.predict import cam_sweep_frames, predict_imgs_from_batch, save_sweep
    cfg = dict(H=32, W=48, NV=4, SB=1, near=1.0, far=2.5, K=16, C=100, G=6, white=True, nr=8, seed=17)
    batch, latent, mlp, _, _ = MG.case_inputs(cfg)
    model = product_model(batch, latent, mlp, "cuda", "parity")
    rend = renderer_for(cfg)
    rend.noise = dict(seed=11)
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    b["target_rgb"] = torch.zeros(1, 3, cfg["H"], cfg["W"], device="cuda")
    ext = b["target_extrinsics"].repeat(3, 1, 1)
    ext[1, 0, 3] += 0.05
    ext[2, 0, 3] += 0.10
    frames = cam_sweep_frames(model, rend, b, ext, cfg["near"], cfg["far"], encode=False)
    # diner.py:209-211: frame order [0, 1, ..., N-1, N-1, ..., 1]
    assert frames.shape == (5, 3, 2 * cfg["H"], cfg["W"]) and torch.equal(frames[2], frames[3]) and torch.equal(frames[1], frames[4])
    assert not torch.equal(frames[0], frames[2]) and bool(torch.isfinite(frames).all())
    rgb, depth = predict_imgs_from_batch(model, rend, b, cfg["near"], cfg["far"], return_depth=True, encode=False)
    assert torch.equal(rgb[0], frames[0][:, :cfg["H"]])                     # same seed -> the sweep's first frame is the plain prediction
    assert torch.equal(IO.torch_cmap(depth)[0].float(), frames[0][:, cfg["H"]:])
    files = save_sweep(frames, str(tmp_path / "sweep.mp4"))
    assert files and all(os.path.exists(f) for f in files)
    w = IO.ImageWriter()
    IO.write_prediction_images(w, str(tmp_path), ["s0"], rgb, depth, b["src_rgbs"], b["target_rgb"])
    assert len(w.close()) == 4


def test_edge_geometry_vs_oracle():
    """Geometry the fixtures do not reach (Appendix A of SURVEY.md: no guard for p.z <= 0, border / zeros padding, exponential std
    padding far outside the image): rays from a target camera placed BETWEEN the source cameras and the object and from one looking
    away, so samples project behind source cameras, far outside their images and onto the padding rings; odd sizes (K = 17, C = 77,
    G = 5, 3 views).  Sampler and render vs the oracle on cuda on every ray, query / compositing vs the CPU oracle stage-wise."""
    from diner_b200 import synthetic as S
    cfg = dict(H=40, W=56, NV=3, SB=1, near=0.2, far=3.5, K=17, C=77, G=5, white=False, nr=8, seed=44)
    batch, latent, mlp, _, _ = MG.case_inputs(cfg)
    H, W = cfg["H"], cfg["W"]
    E = batch["target_extrinsics"].clone()                       # (1,4,4) world->cam
    E2 = E.clone()
    E2[0, :3, 3] += torch.tensor([0.35, -0.2, -1.1])             # moved forward / sideways: the volume straddles the source cameras
    E3 = E.clone()                                               # same camera centre, turned by 180 degrees: looking away from the object
    centre = -E[0, :3, :3].T @ E[0, :3, 3]
    E3[0, :3, :3] = torch.diag(torch.tensor([-1.0, 1.0, -1.0])) @ E[0, :3, :3]
    E3[0, :3, 3] = -E3[0, :3, :3] @ centre
    rays = torch.cat([S.gen_rays(e, batch["target_intrinsics"], W, H, torch.full((1,), cfg["near"]), torch.full((1,), cfg["far"])).view(1, H * W, 8)
                      [:, (torch.arange(160) * 13) % (H * W)] for e in (E2, E3)], dim=1).contiguous()
    NR = rays.shape[1]
    noise = dict(u_coarse=S.hash_uniform((1, NR, cfg["C"]), 44, 901), g_noise=S.hash_normal((1, NR, cfg["G"]), 44, 902),
                 u_fill=S.hash_uniform((1, NR, cfg["K"]), 44, 903))
    scene = O.make_scene_state(batch, latent, mlp)
    sc = oracle_scene_on(scene, "cuda")
    nz = {k: v.cuda().contiguous() for k, v in noise.items()}
    with torch.no_grad():
        zc = O.fill_up_uniform(O.sample_depthguided(sc, rays.cuda(), cfg["K"], cfg["C"], cfg["G"], nz["u_coarse"], nz["g_noise"]), rays.cuda(), nz["u_fill"])
        w_c, rgb_c, dep_c = O.composite(sc, rays.cuda(), zc, cfg["white"])
        w_o, rgb_o, dep_o = O.composite(scene, rays, zc.cpu(), cfg["white"])
    # how much of the edge behaviour the case really exercises (computed with the oracle's own projection)
    xyz = (rays[..., None, :3] + zc.cpu().unsqueeze(-1) * rays[..., None, 3:6]).reshape(1, -1, 3)
    xc = O.world_to_cam(scene, xyz)
    uv = O.project_uv(scene, xc)
    behind = float((xc[..., 2] <= 0).float().mean())
    outside = float(((uv.abs() > 1).any(-1)).float().mean())
    print("edge geometry: %.1f %% of the sample-views behind a source camera, %.1f %% outside its image" % (100 * behind, 100 * outside))
    assert behind > 0.02 and outside > 0.2
    for mode, tol in (("fp32", 2e-5), ("parity", PAR_TOL)):
        model = product_model(batch, latent, mlp, "cuda", mode)
        ctx = model.context()
        z = ctx.sample(rays.cuda(), cfg["K"], cfg["C"], cfg["G"], nz)
        assert not bool(_bad((z - zc).abs().max(dim=-1).values, Z_TOL).any()), "sampler differs from the reference on cuda"
        w, rgb, dep = ctx.composite(rays.cuda(), zc.contiguous(), cfg["white"], model.mode_id())
        finite = torch.isfinite(rgb_o).all(-1) & torch.isfinite(dep_o)                 # the reference itself yields NaN where 0/0 projections reach the MLP
        same_nan = (torch.isfinite(rgb.cpu()).all(-1) & torch.isfinite(dep.cpu())) == finite
        e = ray_err(rgb.cpu()[finite], dep.cpu()[finite], rgb_o[finite], dep_o[finite])
        print("edge geometry / %s: max |err| vs CPU oracle %.3g on %d finite rays (%d rays non-finite in the reference too)" % (
            mode, float(e.max()) if e.numel() else 0.0, int(finite.sum()), int((~finite).sum())))
        assert bool(same_nan.all()) and not bool(_bad(e, tol).any())
        out = renderer_for(cfg, noise)(model, rays.cuda())
        fin_c = torch.isfinite(rgb_c).all(-1) & torch.isfinite(dep_c)
        assert not bool(_bad(ray_err(out.fine.rgb[fin_c], out.fine.depth[fin_c], rgb_c[fin_c], dep_c[fin_c]), TOL).any())

"""Benchmark of the DINER render hot path (BASELINE.json metric: rays/s at 512x512, 4 source views,
64 samples/ray).  One "step" = NeRFRendererDGS.forward over the whole 512x512 target image
(262 144 rays; depth-guided sampling -> feature projection/fusion -> positional encoding -> MLP ->
compositing), scene encode excluded.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode parity|fast|fp32] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU): the rays of the image are split into N contiguous
shards (strong scaling, BASELINE.json configs[3]), every rank renders its shard on a replicated scene
and ONE NCCL all-gather of rgb|depth ends the step.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402

H = W = 512
NV, K, C, G = 4, 64, 1000, int(15 * 64 / 40)
NEAR, FAR = 0.3211, 1.2041                     # DTU: 400*s, 1500*s with s = 0.7/872
WHITE = False
SEED = 0
FLOP_PER_SAMPLE = 4774912 * NV + 2101248       # algorithmic MLP FLOPs per sample
FLOP_PRE_PER_SAMPLE = 4774912 * NV             # ... of which in the per-sample-view (PRE) kernel
WORKLOAD = "DTU-shaped synthetic 512x512, 4 src views, 64 samples/ray, 1000 depth candidates, 24 gaussian"
RAYS_LIMIT = None                              # --rays: render only the first N rays of the image (per-GPU share of a bigger job)


def select_workload(name, rays=None):
    """Default = BASELINE.json configs[1] (the headline).  'stress1024' = configs[4]: 1024x1024, 8 source views, 256 samples/ray;
    with --rays 131072 it is the per-GPU share of that image on 8 GPUs.  Not part of the driver contract (default unchanged)."""
    global H, W, NV, K, G, FLOP_PER_SAMPLE, FLOP_PRE_PER_SAMPLE, WORKLOAD, RAYS_LIMIT
    if name == "stress1024":
        H = W = 1024
        NV, K = 8, 256
        G = int(15 * K / 40)
        WORKLOAD = "stress: DTU-shaped synthetic 1024x1024, 8 src views, 256 samples/ray, 1000 depth candidates, %d gaussian" % G
    elif name != "dtu512":
        raise SystemExit("unknown --workload %s" % name)
    FLOP_PER_SAMPLE = 4774912 * NV + 2101248
    FLOP_PRE_PER_SAMPLE = 4774912 * NV
    RAYS_LIMIT = rays
    if rays:
        WORKLOAD += " (first %d rays of the image)" % rays


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def build_inputs():
    from diner_b200 import synthetic as S
    batch = S.make_scene(H, W, NV, 1, NEAR, FAR, SEED)
    Hl, Wl = (H + 128) // 2, (W + 128) // 2
    gen = torch.Generator().manual_seed(SEED)
    latent = torch.randn(1, NV, 512, Hl, Wl, generator=gen) * 0.5          # 839 MB fp32 (stand-in for the ResNet pyramid)
    mlp = S.make_mlp_state(seed=SEED)
    rays = S.gen_rays(batch["target_extrinsics"], batch["target_intrinsics"], W, H,
                      torch.full((1,), NEAR), torch.full((1,), FAR)).view(1, H * W, 8).contiguous()
    if RAYS_LIMIT:
        start = max(0, (H // 2) * W - RAYS_LIMIT // 2)          # rows around the image centre
        rays = rays[:, start:start + RAYS_LIMIT].contiguous()
    return batch, latent, mlp, rays


def strided_pick(n_rays):
    """Indices of a bounded, strided sample of the image's rays (every value < H*W for any n_rays <= H*W)."""
    return (torch.arange(n_rays) * (H * W // n_rays) + 131) % (H * W)


def cpu_baseline(batch, latent, mlp, rays, n_rays=12288, warm=256, check=None):
    """The oracle port (torch-CPU restatement pinned to the reference) on this box's host cores, bounded sample.
    `check(rays_sample, noise, rgb, depth, z)` (optional) is handed the sample and the oracle's outputs for it: this leg is the
    one place of the bench where the oracle runs, so the parity numbers of the line are taken from the same run."""
    from oracle import diner_oracle as O
    from diner_b200 import synthetic as S
    scene = O.make_scene_state(batch, latent, mlp)
    pick = strided_pick(n_rays)
    r = rays[:, pick].contiguous()

    def run(rr):
        n = rr.shape[1]
        noise = dict(u_coarse=S.hash_uniform((1, n, C), 1, 1), g_noise=S.hash_normal((1, n, G), 1, 2), u_fill=S.hash_uniform((1, n, K), 1, 3))
        return noise, O.render(scene, rr, K, C, G, WHITE, noise["u_coarse"], noise["g_noise"], noise["u_fill"], return_z=True)
    with torch.no_grad():
        run(r[:, :warm])
        t0 = time.time()
        noise, (rgb_o, dep_o, _, z_o) = run(r)
        dt = time.time() - t0
    out = dict(value=n_rays / dt, unit="rays/s", cores=torch.get_num_threads(), kind="port",
               sample="%d rays of the 512x512 workload (strided), oracle/diner_oracle.py on torch-CPU fp32, %.1f s" % (n_rays, dt))
    if check is not None:
        out["_parity"] = check(r, noise, rgb_o, dep_o, z_o)
    return out


def gpu_eager_baseline(batch, latent, mlp, rays, dev, n_rays=4096):
    """The reference algorithm as eager PyTorch ON THE GPU (oracle port moved to cuda): the stand-in for the
    reference's own CUDA execution (SURVEY §8(d) 'GPU reference baseline'), one ray batch of 4096 = diner.py:57."""
    from oracle import diner_oracle as O
    from diner_b200 import synthetic as S
    scene = O.make_scene_state(batch, latent, mlp)               # built on the host, then moved
    for f in ("poses", "focal", "c", "image_shape", "latent", "depths", "depths_std", "normals"):
        setattr(scene, f, getattr(scene, f).to(dev))
    scene.mlp = {k: v.to(dev) for k, v in mlp.items()}
    start = (H // 2) * W - n_rays // 2
    r = rays[:, start:start + n_rays].contiguous().to(dev)
    noise = [S.hash_uniform((1, n_rays, C), 1, 1).to(dev), S.hash_normal((1, n_rays, G), 1, 2).to(dev), S.hash_uniform((1, n_rays, K), 1, 3).to(dev)]
    with torch.no_grad():
        for _ in range(2):
            O.render(scene, r, K, C, G, WHITE, *noise)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            O.render(scene, r, K, C, G, WHITE, *noise)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return dict(value=n_rays / (ms * 1e-3), unit="rays/s", kind="reference algorithm as eager fp32 PyTorch on this GPU (oracle port on cuda)",
                sample="%d-ray batch (image centre), %.1f ms" % (n_rays, ms))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the reference arm is a CPU measurement on all host cores
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    batch, latent, mlp, rays = build_inputs()
    n_rays = 4096
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(batch, latent, mlp, rays, n_rays=n_rays, warm=64 if i == 0 else 8)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = sum(vals) / len(vals)
    cb["value"] = v
    cb["sample"] = "each step = %d strided rays of the workload through oracle/diner_oracle.py (torch-CPU fp32 port of the reference; " \
                   "the Python reference itself cannot travel to this box)" % n_rays
    print(json.dumps({"impl": "reference", "metric": "rays_per_sec", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_rays / v,
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": WORKLOAD},
                      "cpu_baseline": cb,
                      "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def train_bench(args):
    """BASELINE.json configs[2] (NOT part of the driver contract): Facescape-shaped
    training step -- SB=4 scenes of 256x256, 4 source views, 128 samples/ray, 4096 random rays per scene (ray_batch_size,
    diner.py:57), MSE loss, backward through the renderer to the ResnetFC parameters and the latent maps, Adam step
    (diner.py:333).  Forward in --mode; backward through diner_render_backward (tcgen05 GEMMs of csrc/gemm_tc3.cu for the 512-wide
    layers; DINER_B200_BACKWARD_TC=0 selects the fp32 CUDA-core path).  Single GPU."""
    from diner_b200 import synthetic as S
    from diner_b200.nerf_renderer import NeRFRendererDGS
    from diner_b200.predict import calc_losses
    from diner_b200.synthetic import product_model
    Ht = Wt = 256
    SBt, NVt, Kt, RB = 4, 4, 128, 4096
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    batch = S.make_scene(Ht, Wt, NVt, SBt, 1.0, 2.5, SEED)
    gen = torch.Generator().manual_seed(SEED)
    latent = torch.randn(SBt, NVt, 512, (Ht + 128) // 2, (Wt + 128) // 2, generator=gen) * 0.5
    model = product_model(batch, latent, S.make_mlp_state(seed=SEED), dev, args.mode).train()
    model.encoder.latent = model.encoder.latent.detach().clone().requires_grad_(True)
    model.encoder.scene_version += 1
    rend = NeRFRendererDGS(n_samples=Kt, n_depth_candidates=C, n_gaussian=int(15 * Kt / 40), white_bkgd=True)
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    b["target_rgb"] = torch.rand(SBt, 3, Ht, Wt, generator=gen).to(dev)
    opt = torch.optim.Adam(list(model.mlp_fine.parameters()) + [model.encoder.latent], lr=1e-4)
    g = torch.Generator().manual_seed(1)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = calc_losses(model, rend, b, 1.0, 2.5, RB, generator=g, encode=False)["total"]
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # gradient check (outside the timed region): diner_render_backward against the loss / gradient digests of the UNMODIFIED
    # reference's autograd (tests/golden/grads_cfg1_face64.pt, produced by oracle/make_golden.py)
    grad_check = None
    try:
        from oracle import diner_oracle as O
        from oracle import make_golden as MG
        from diner_b200.nerf_renderer import mlp_param_order
        gg = torch.load(os.path.join(ROOT, "tests", "golden", "grads_cfg1_face64.pt"))
        cfg_g, batch_g, latent_g, mlp_g, rays_g, _, gt_g = MG.grad_case_inputs()
        mg = product_model(batch_g, latent_g, mlp_g, dev, args.mode)
        cg = mg.context()
        rg, zg = rays_g.to(dev), gg["z"].to(dev).contiguous()
        _, rgb_g, _ = cg.composite(rg, zg, cfg_g["white"], mg.mode_id(), want_weights=False)
        g_rgb = (2.0 * (rgb_g - gt_g.to(dev)) / rgb_g.numel()).contiguous()
        gp, dl = cg.render_backward(rg, zg, cfg_g["white"], g_rgb, None, True, tuple(latent_g.shape))
        worst, off = 0.0, 0
        named = dict(mg.mlp_fine.named_parameters())
        for k in mlp_param_order(mg.mlp_fine):
            n = named[k].numel()
            d = O.grad_digest(gp[off:off + n].cpu(), 256)
            worst = max(worst, abs(d["norm"] - gg["grads"][k]["norm"]) / max(gg["grads"][k]["norm"], 1e-30))
            off += n
        dlat = O.grad_digest(dl.cpu(), 4096)
        grad_check = {"reference": "unmodified reference autograd (tests/golden/grads_cfg1_face64.pt)",
                      "loss_abs_err": abs(float(((rgb_g.cpu() - gt_g) ** 2).mean()) - gg["loss"]),
                      "worst_param_grad_norm_rel_err": worst,
                      "latent_grad_norm_rel_err": abs(dlat["norm"] - gg["latent_grad"]["norm"]) / gg["latent_grad"]["norm"]}
    except Exception as e:                       # reported extra, never allowed to break the line
        grad_check = {"unavailable": repr(e)[:200]}
    rays_step = SBt * RB
    flop = 3 * rays_step * Kt * (4774912 * NVt + 2101248)          # forward + ~2x for dgrad + wgrad
    print(json.dumps({"metric": "train_rays_per_sec", "value": rays_step / (ms * 1e-3), "unit": "rays/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": ("forward %s (fused tcgen05 kernel), backward: tcgen05 GEMMs (f16x3 operands, f32 accumulate) for the 512-wide layers"
                                if os.environ.get("DINER_B200_BACKWARD_TC", "1") != "0" else "forward %s, backward f32 CUDA cores") % args.mode,
                      "data": "synthetic",
                      "config": {"workload": "Facescape-shaped synthetic training step: SB=4 x 256x256, 4 src views, 128 samples/ray, "
                                             "4096 rays/scene, MSE + backward + Adam", "mode": args.mode},
                      "loss": float(loss), "algorithmic_tflops": flop / (ms * 1e-3) / 1e12, "grad_check": grad_check}))


def latest_traffic(workload="dtu512"):
    """dram bytes per launch from the newest committed ncu --set full summary of this workload's shape (profiles/*_traffic.json,
    written by tools/ncu_summary.py from the .ncu-rep of the same kernel; `workload` / `samples` say what was captured); None when
    there is none."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            d = json.load(open(f))
            if d.get("workload", "dtu512") == workload:
                best = dict(d, file=os.path.relpath(f, ROOT))
        except Exception:
            pass
    return best


def parity_block(model, dev):
    """-> check callback for cpu_baseline: the CUDA path on the SAME rays and noise as the oracle run of the cpu_baseline leg,
    outside every timed region.  Stage-wise on the oracle's own sample depths (max |err| over all rays: the 1e-4 bar) and end to end
    with the same injected noise (PSNR; the rays beyond 1e-4 there are erf-ulp shortlist flips between torch-CPU and CUDA, which
    tests/test_gpu_parity.py resolves against the oracle run on cuda)."""
    def check(r, noise, rgb_o, dep_o, z_o):
        from oracle.diner_oracle import psnr
        with torch.no_grad():
            ctx = model.context()
            _, rgb_s, dep_s = ctx.composite(r.to(dev), z_o.to(dev).contiguous(), WHITE, model.mode_id(), want_weights=False)
            rgb_e, dep_e, _, _ = ctx.render(r.to(dev), K, C, G, WHITE, model.mode_id(), {k: v.to(dev).contiguous() for k, v in noise.items()})
        e2e = torch.maximum((rgb_e.cpu() - rgb_o).abs().max(-1).values, (dep_e.cpu() - dep_o).abs())
        return {"rays": int(r.shape[1]), "reference": "oracle/diner_oracle.py on torch-CPU (pinned to the reference's goldens), the cpu_baseline run",
                "max_abs_err_rgb": float((rgb_s.cpu() - rgb_o).abs().max()), "max_abs_err_depth": float((dep_s.cpu() - dep_o).abs().max()),
                "psnr_vs_ref_db": psnr(rgb_s.cpu(), rgb_o), "stage": "given the oracle's sample depths (every ray)",
                "end_to_end": {"psnr_vs_ref_db": psnr(rgb_e.cpu(), rgb_o), "median_abs_err": float(e2e.median()),
                               "frac_rays_beyond_1e-4": float((~(e2e <= 1e-4)).float().mean())}}
    return check


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", default=os.environ.get("DINER_B200_MODE", "parity"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="dtu512", help="dtu512 (default, BASELINE configs[1]) | stress1024 (configs[4]) | train256 (configs[2])")
    ap.add_argument("--rays", type=int, default=0, help="render only N rays of the image (0 = all)")
    args = ap.parse_args()
    if args.workload == "train256":
        return train_bench(args)
    select_workload(args.workload, args.rays or None)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        return reference_arm(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus

    from diner_b200.synthetic import product_model
    from diner_b200.nerf_renderer import NeRFRendererDGS
    batch, latent, mlp, rays = build_inputs()
    model = product_model(batch, latent, mlp, dev, args.mode)
    rend = NeRFRendererDGS(n_samples=K, n_depth_candidates=C, n_gaussian=G, white_bkgd=WHITE)
    n_total = rays.shape[1]
    per = (n_total + world - 1) // world
    lo, hi = rank * per, min(n_total, (rank + 1) * per)
    from diner_b200.multi_gpu import render_sharded
    rays_host = rays.contiguous().pin_memory()          # every rank holds the ray list; it renders rays[lo:hi]
    rays_dev = rays_host.to(dev)
    ctx = model.context()
    if os.environ.get("DINER_RAY_IMAGE_WIDTH") is None:
        ctx.set_option("ray_image_width", W)            # the ray list is gen_rays' row-major image (every shard = whole rows of it)

    def packed_render(r, out, ray_offset):
        rend.render_packed(model, r, out=out, ray_offset=ray_offset)    # compositing kernel writes rgb|depth into the gather slice

    weights = None

    def step(src):
        # one in-place NCCL all-gather per image when world > 1; shards are whole image rows, weighted by the measured speed of each GPU
        return render_sharded(packed_render, src, packed=True, return_packed=True, weights=weights, align=W)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(rays_dev)
    sync()
    if world > 1 and os.environ.get("DINER_BALANCE", "0") == "1":
        # optional (off by default: measured on 8 B200s the per-GPU time differences of one step are noise, not persistent speed
        # differences -- balanced shards 91.6 ms vs equal shards 89.4 ms per image): time this rank's equal shard once more, gather the
        # times and re-cut the shards proportionally (multi_gpu.balance_weights)
        from diner_b200.multi_gpu import balance_weights
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tmp = torch.empty(1, hi - lo, 4, device=dev)
        a0.record()
        packed_render(rays_dev[:, lo:hi], tmp, lo)
        a1.record()
        torch.cuda.synchronize()
        weights = balance_weights(a0.elapsed_time(a1) * 1e-3)
        from diner_b200.multi_gpu import shard_bounds
        lo, hi, _ = shard_bounds(n_total, world, rank, weights, W)       # this rank's shard from here on (roofline bookkeeping below)
        step(rays_dev)
        sync()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(rays_dev)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    # end-to-end: host rays in, rgb + depth back on the host, every step.  One GPU: through the C-ABI host-buffer entry
    # (diner_render_host: H2D copy, render, D2H copies and the stream sync inside the call).  N GPUs: pinned host rays -> device,
    # sharded render + all-gather, packed image back to the host.
    rgb_h = torch.empty(1, n_total, 3).pin_memory()
    dep_h = torch.empty(1, n_total).pin_memory()
    img_h = torch.empty(1, n_total, 4).pin_memory()
    sync()
    t0 = time.perf_counter()
    for i in range(args.steps):
        if world == 1:
            ctx.render_host(rays_host, K, C, G, WHITE, model.mode_id(), 1000 + i, rgb_h, dep_h)
        else:
            img_h.copy_(step(rays_host.to(dev, non_blocking=True)), non_blocking=True)
            torch.cuda.synchronize()
    sync()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms, ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # per-kernel breakdown for the roofline: a separate short pass with CUDA-event timing inside the library (events are
    # recorded on the stream the kernels are launched on)
    ctx.set_timing(True)
    stage = {"sampler": 0.0, "mlp_pre": 0.0, "mlp_post": 0.0, "composite": 0.0}
    nt = 2
    ctx.set_option("rebuild_maps", 1)                   # time the once-per-scene build of the hoisted lin_z maps as well
    scene_prepare_ms = None
    for i_ in range(nt):
        step(rays_dev)
        torch.cuda.synchronize()
        for k_, v_ in ctx.last_stage_ms().items():
            if k_ == "lin_z_maps":
                scene_prepare_ms = v_ if i_ == 0 else scene_prepare_ms
            else:
                stage[k_] += v_ / nt
    ctx.set_timing(False)

    if rank == 0:
        pk = peaks()
        rays_per_s = n_total * args.steps / (ms * 1e-3)
        n_samp_rank = (hi - lo) * K
        fused = stage["mlp_post"] == 0.0                # FUSED launch: one kernel runs the per sample-view AND the per sample layers
        n_pre_launches = 1 if fused else -(-n_samp_rank // 524288)
        pre_ms_per_launch = stage["mlp_pre"] / n_pre_launches if stage["mlp_pre"] > 0 else None
        pre_flops_per_launch = (FLOP_PER_SAMPLE if fused else FLOP_PRE_PER_SAMPLE) * n_samp_rank / n_pre_launches
        achieved = pre_flops_per_launch / (pre_ms_per_launch * 1e-3) / 1e12 if pre_ms_per_launch else 0.0
        # FLOPs the kernel actually issues to the tensor pipe per sample: NV x (lin_in + 3 x (fc_0 + fc_1)) (lin_z is hoisted into
        # the once-per-scene Y maps) [+ 2 x (fc_0 + fc_1) in the fused kernel; lin_out runs on the CUDA cores], x3 MMAs per product in parity mode
        passes = 3 if args.mode == "parity" else 1
        exec_per_sample = NV * (2 * 64 * 512 + 6 * 2 * 512 * 512) + (4 * 2 * 512 * 512 if fused else 0)
        executed = passes * exec_per_sample * n_samp_rank / n_pre_launches / (pre_ms_per_launch * 1e-3) / 1e12 if pre_ms_per_launch else 0.0
        tr = latest_traffic(args.workload) if args.mode == "parity" else None                    # captured on this workload's shape
        e2e_d2h = (rgb_h.numel() + dep_h.numel()) * 4 if world == 1 else img_h.numel() * 4
        line = {
            "metric": "rays_per_sec", "value": rays_per_s, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": {"parity": "f16x3 (fp16 hi/lo split operands, f32 accumulate: the 1e-4 parity mode)", "fast": "f16, f32 accumulate",
                      "fp32": "f32"}[args.mode],
            "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "run": {"mode": args.mode, "rays_per_step": n_total, "sharding": "rays/%d" % world,
                    "l2": "inputs larger than L2 (%.1f GB fp32 lin_z maps gathered per sample-view, %.0f MB of sample depths + per-sample outputs per step)"
                          % (3 * NV * ((H + 128) // 2) * ((W + 128) // 2) * 512 * 4 / 1e9, n_total * K * 20 / 1e6),
                    "shard_weights": [round(x / max(weights), 4) for x in weights] if weights else None,
                    "e2e_path": "diner_render_host (C ABI, host buffers)" if world == 1 else "pinned host rays -> sharded render + all-gather -> host image"},
            "e2e": {"value": n_total * args.steps / (ms_e2e * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": rays_host.numel() * 4, "d2h_bytes_per_step": e2e_d2h},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": ("tc2::mlp_pair_kernel<FUSED> (the whole ResnetFC: per sample-view layers, view combine, per sample layers; "
                                                       "algorithmic FLOPs per SURVEY 8(d) = 21.2 MFLOP per sample, incl. the hoisted lin_z)") if fused else
                                                      "tc2::mlp_pair_kernel<PRE> (per sample-view ResnetFC layers; algorithmic FLOPs per SURVEY 8(d), incl. the hoisted lin_z)",
                         "launches_per_step": n_pre_launches,
                         "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["sustained"], "peak_source": pk["src"] + " bf16 dense sustained (fp16 runs at the same rate)",
                         "achieved_executed_mma": executed,
                         "executed_note": "tensor-pipe FLOP/s the kernel really issues (lin_z hoisted out, x3 passes in parity mode); burst peak %.1f" % pk["burst"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel from the newest ncu --set full capture.  The capture
                         # launches the kernel on a 524 288-sample slab of this workload (ncu replays every launch ~40 times); this run's
                         # launch covers n_samp_rank samples, so the captured bytes are scaled by the sample ratio
                         "traffic": tr["dram_bytes_per_launch"] * n_samp_rank / n_pre_launches / float(tr.get("samples", 524288)) if tr else None,
                         "traffic_captured": {"bytes": tr["dram_bytes_per_launch"], "samples": tr.get("samples", 524288)} if tr else None,
                         "traffic_source": ("ncu --set full, %s" % tr["file"]) if tr else None,
                         "whole_step_frac": rays_per_s / world * K * FLOP_PER_SAMPLE / 1e12 / pk["sustained"],
                         "stage_ms_per_step": stage,
                         "once_per_scene_ms": {"lin_z_maps (hoisted lin_z over all latent pixels, excluded from the step like the scene encode)": scene_prepare_ms}},
        }
        if not args.no_cpu_baseline and world == 1 and args.workload == "dtu512" and not args.rays:
            line["cpu_baseline"] = cpu_baseline(batch, latent, mlp, rays, check=parity_block(model, dev))
            line["parity"] = line["cpu_baseline"].pop("_parity")
            try:
                del model
                torch.cuda.empty_cache()
                line["gpu_eager_baseline"] = gpu_eager_baseline(batch, latent, mlp, rays, dev)
            except Exception as e:      # reported extra, never allowed to break the contract line
                line["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

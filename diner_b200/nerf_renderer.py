"""NeRFRendererDGS with the reference's constructor, mutable attributes and forward() contract
(src/models/nerf_renderer.py:23-37,399-430), served by libdiner_b200.

`n_samples` / `n_gaussian` are read at call time because the reference CLI reassigns them on the live
module (python_scripts/create_prediction_folder.py:44-47)."""
import torch

try:                                    # the reference returns dotmap.DotMap; use it when installed
    from dotmap import DotMap
except Exception:                       # same attribute-access contract without the dependency
    class DotMap(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v


def mlp_param_order(mlp):
    """Names of the ResnetFC parameters in libdiner_b200's canonical order (the argument order of diner_set_mlp)."""
    names = ["lin_in.weight", "lin_in.bias", "lin_out.weight", "lin_out.bias"]
    for b in range(mlp.n_blocks):
        names += ["blocks.%d.fc_0.weight" % b, "blocks.%d.fc_0.bias" % b, "blocks.%d.fc_1.weight" % b, "blocks.%d.fc_1.bias" % b]
    for b in range(min(mlp.combine_layer, mlp.n_blocks)):
        names += ["lin_z.%d.weight" % b, "lin_z.%d.bias" % b]
    return names


class _RenderWithGrad(torch.autograd.Function):
    """Training-step path (src/models/diner.py:257-266): forward through the fused kernels, backward through
    diner_render_backward (fp32 CUDA cores; validated against the reference's autograd gradients in tests/test_gpu_parity.py).
    Gradients reach the ResnetFC parameters and encoder.latent; the sampler is @torch.no_grad in the reference too
    (nerf_renderer.py:65)."""

    @staticmethod
    def forward(ctx, renderer, model, rays, latent, *params):
        c = model.context()
        rgb, depth, _, z = c.render(rays, int(renderer.n_samples), int(renderer.n_depth_candidates), int(renderer.n_gaussian),
                                    renderer.white_bkgd, model.mode_id(), renderer._noise_for_call(), want_z=True)
        ctx.save_for_backward(rays, z)
        ctx.model, ctx.white, ctx.latent_shape = model, renderer.white_bkgd, tuple(latent.shape)
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.need_latent = latent.requires_grad
        return rgb, depth

    @staticmethod
    def backward(ctx, g_rgb, g_depth):
        rays, z = ctx.saved_tensors
        c = ctx.model.context()
        gp, dl = c.render_backward(rays, z, ctx.white, g_rgb.float().contiguous(),
                                   g_depth.float().contiguous() if g_depth is not None else None,
                                   want_latent_grad=ctx.need_latent, latent_shape=ctx.latent_shape)
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for d in shp:
                n *= d
            grads.append(gp[off:off + n].view(shp))
            off += n
        assert off == gp.numel()
        return (None, None, None, dl) + tuple(grads)


class NeRFRendererDGS(torch.nn.Module):
    def __init__(self, n_samples=40, n_depth_candidates=1000, n_gaussian=15, eval_batch_size=100000,
                 white_bkgd=True):
        super().__init__()
        self.n_samples = n_samples
        self.n_depth_candidates = n_depth_candidates
        self.n_gaussian = n_gaussian
        self.eval_batch_size = eval_batch_size      # kept for API parity; the fused kernels need no chunking
        self.white_bkgd = white_bkgd
        self.noise = None   # optional dict(u_coarse, g_noise, u_fill, seed): injected draws (tests) / seed
        self._calls = 0

    def _noise_for_call(self, ray_offset=0):
        """Injected dense noise / fixed seed when `self.noise` is set; else a fresh seed per call.  `ray_offset` = logical index
        of the call's first ray (ray-sharded renders), see diner_noise.ray_offset."""
        if self.noise is not None:
            return dict(self.noise, ray_offset=ray_offset) if ray_offset else self.noise
        self._calls += 1
        return dict(seed=(torch.initial_seed() * 1000003 + self._calls) & 0xFFFFFFFFFFFFFFFF, ray_offset=ray_offset)

    @torch.no_grad()
    def sample_depthguided(self, rays, model, n_samples, n_candidates, depth_diff_max=0.05, n_gaussian=None):
        """Depth-guided shortlist, nerf_renderer.py:65-190.  Returns (SB,NR,n_samples) with 0 = empty slot,
        sorted ascending (the reference returns the same multiset ordered by likelihood)."""
        G = self.n_gaussian if n_gaussian is None else n_gaussian
        assert n_samples >= G
        ctx = model.context()
        ctx.set_float_option("depth_diff_max", depth_diff_max)
        try:
            _, zd = ctx.sample(rays.float().contiguous(), n_samples, n_candidates, G, self._noise_for_call(), want_dgs=True)
        finally:
            ctx.set_float_option("depth_diff_max", 0.05)         # forward() always uses the default (nerf_renderer.py:415-417)
        return zd

    def composite(self, model, rays, z_samp):
        """nerf_renderer.py:286-365 -> (weights (SB,B,K), rgb (SB,B,3), depth (SB,B))."""
        model._no_grad_only(rays, z_samp)
        return model.context().composite(rays.float().contiguous(), z_samp.float().contiguous(), self.white_bkgd,
                                         model.mode_id(), want_weights=True)

    def forward(self, model, rays, want_weights=False):
        """rays (SB,B,8) [origin3, dir3, near, far] -> DotMap(fine=DotMap(rgb (SB,B,3), depth (SB,B)[, weights]))."""
        assert len(rays.shape) == 3
        if (torch.is_grad_enabled() and not want_weights and
                (any(p.requires_grad for p in model.mlp_fine.parameters()) or model.encoder.latent.requires_grad)):
            named = dict(model.mlp_fine.named_parameters())
            params = [named[k] for k in mlp_param_order(model.mlp_fine)]
            rgb, depth = _RenderWithGrad.apply(self, model, rays.float().contiguous(), model.encoder.latent, *params)
            return DotMap(fine=self._format_outputs(None, rgb, depth, False))
        model._no_grad_only(rays)
        rgb, depth, w, _ = model.context().render(
            rays.float().contiguous(), int(self.n_samples), int(self.n_depth_candidates), int(self.n_gaussian),
            self.white_bkgd, model.mode_id(), self._noise_for_call(), want_weights=want_weights)
        return DotMap(fine=self._format_outputs(w, rgb, depth, want_weights))

    @torch.no_grad()
    def render_packed(self, model, rays, out=None, ray_offset=0):
        """forward() with the outputs packed as (SB,B,4) = [r,g,b,depth] (optionally written into `out`): what the ray-sharded
        multi-GPU render all-gathers (diner_b200/multi_gpu.py).  `ray_offset`: index of rays[:, 0] in the full ray list."""
        assert len(rays.shape) == 3
        return model.context().render_rgbd(rays.float().contiguous(), int(self.n_samples), int(self.n_depth_candidates),
                                           int(self.n_gaussian), self.white_bkgd, model.mode_id(),
                                           self._noise_for_call(ray_offset), out=out)

    @torch.no_grad()
    def render_image(self, model, target_extrinsics, target_intrinsics, H, W, z_near, z_far):
        """Whole target view(s) in ONE library call: gen_rays (src/util/cam_geometry.py:5-48) + the ray_batch_size split /
        torch.cat loop of DINER.predict_imgs_from_batch (src/models/diner.py:79-92).  -> rgb (SB,H*W,3), depth (SB,H*W)."""
        noise = self._noise_for_call()
        return model.context().render_image(
            target_extrinsics.float().contiguous(), target_intrinsics.float().contiguous(), int(H), int(W), float(z_near),
            float(z_far), int(self.n_samples), int(self.n_depth_candidates), int(self.n_gaussian), self.white_bkgd,
            model.mode_id(), dict(seed=noise.get("seed", 0)))

    def _format_outputs(self, weights, rgb, depth, want_weights):
        out = DotMap(rgb=rgb, depth=depth)
        if want_weights:
            out.weights = weights
        return out

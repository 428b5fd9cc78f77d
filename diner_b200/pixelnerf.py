"""PixelNeRF with the reference's constructor, buffers, `encode` and `forward` signatures
(src/models/pixelnerf.py:13-145); `forward` is served by libdiner_b200 (no PyTorch compute path).

Extra, non-reference attributes:
  mode        'parity' (tcgen05 fp16x3, default) | 'fast' (tcgen05 fp16) | 'fp32' (CUDA cores)
              -- also settable through the environment variable DINER_B200_MODE
"""
import os

import torch
from torchvision.transforms import Normalize

from . import capi
from .import_helper import import_obj
from .positional_encoding import PositionalEncoding
from .scene_ops import depth2normal


def _kw(conf):
    k = conf["kwargs"] if isinstance(conf, dict) else conf.kwargs
    return dict(k)


def _mod(conf):
    return conf["module"] if isinstance(conf, dict) else conf.module


class PixelNeRF(torch.nn.Module):
    def __init__(self, poscode_conf, encoder_conf, mlp_fine_conf):
        super().__init__()
        self.poscode = PositionalEncoding(**_kw(poscode_conf), d_in=3)
        self.depthcode = PositionalEncoding(**_kw(poscode_conf), d_in=1)
        self.encoder = import_obj(_mod(encoder_conf))(**_kw(encoder_conf))
        self.d_in = self.poscode.d_out + self.depthcode.d_out + 3
        self.d_latent = self.encoder.latent_size
        self.d_out = 4
        self.mlp_fine = import_obj(_mod(mlp_fine_conf))(**_kw(mlp_fine_conf), d_latent=self.d_latent,
                                                        d_in=self.d_in, d_out=self.d_out)
        self.register_buffer("poses", torch.empty(1, 3, 4), persistent=False)
        self.register_buffer("image_shape", torch.empty(2), persistent=False)
        self.register_buffer("focal", torch.empty(1, 2), persistent=False)
        self.register_buffer("c", torch.empty(1, 2), persistent=False)
        self.normalize_rgb = Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        self.mode = os.environ.get("DINER_B200_MODE", "parity")
        self._ctx = None
        self._mlp_stamp = None
        self._scene_stamp = None
        self._beta = 0.0

    # ------------------------------------------------------------------------------------------
    def encode(self, images, depths, depths_std, extrinsics, intrinsics):
        """Feature maps + camera buffers for the following forward() calls (pixelnerf.py:35-53)."""
        images = self.normalize_rgb(images)
        if depths.is_cuda:      # CUDA kernel (diner_depth2normal); the torch version only serves CPU-side module tests
            normals = self._bare_context(depths.device).depth2normal(
                depths.flatten(end_dim=1).float().contiguous(), intrinsics.flatten(end_dim=1).float().contiguous()).reshape_as(images)
        else:
            normals = depth2normal(depths.flatten(end_dim=1), intrinsics.flatten(end_dim=1)).reshape_as(images)
        self.encoder(images, depths, depths_std, normals)
        self.set_cameras(extrinsics, intrinsics, images.shape[-1], images.shape[-2])

    def set_cameras(self, extrinsics, intrinsics, W, H):
        self.poses = extrinsics
        self.c = intrinsics[:, :, :2, -1]
        self.focal = torch.stack((intrinsics[:, :, 0, 0], intrinsics[:, :, 1, 1]), dim=-1)
        self.image_shape[0] = W
        self.image_shape[1] = H
        self._scene_stamp = None

    # ------------------------------------------------------------------------------------------
    def _bare_context(self, dev):
        if dev.type != "cuda":
            raise RuntimeError("diner_b200 renders on CUDA only; move the model and batch to a B200 (got %s)" % dev)
        dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        if self._ctx is None or self._ctx.device != dev:
            self._ctx = capi.Context(dev)
            self._mlp_stamp = self._scene_stamp = None
            self._beta = 0.0
        return self._ctx

    def context(self):
        """libdiner_b200 context with the current parameters and scene uploaded (lazy, versioned)."""
        dev = self.poses.device
        self._bare_context(dev)
        m = self.mlp_fine
        if m.combine_type != "average":                     # the reference's combine() raises for anything else (resnetfc.py:9-14)
            raise NotImplementedError(m.combine_type)
        beta = float(getattr(m, "beta", 0.0) or 0.0)
        if beta != self._beta:                              # Softplus(beta) activations (resnetfc.py:124-127)
            self._ctx.set_float_option("softplus_beta", beta)
            self._beta = beta
        sd, stamp = m.packed_state()
        if stamp != self._mlp_stamp:
            self._ctx.set_mlp(sd, m.d_in, m.d_latent, m.d_hidden, m.d_out, m.n_blocks, m.combine_layer)
            self._mlp_stamp = stamp
        enc = self.encoder
        if enc.nviews is None:
            raise RuntimeError("PixelNeRF.encode() must be called before rendering")
        sstamp = (enc.scene_version, enc.latent.data_ptr(), enc.latent._version, self.poses.data_ptr(), self.poses._version)
        if sstamp != self._scene_stamp:
            if enc.index_interp != "bilinear" or enc.index_padding != "border":
                raise NotImplementedError("libdiner_b200 implements bilinear/border latent indexing only")
            f32 = lambda t: t.detach().float().contiguous()
            lat = enc.latent.detach().float()
            nhwc = lat.dim() == 5 and lat.permute(0, 1, 3, 4, 2).is_contiguous() and not lat.is_contiguous()
            # channels-last latent (what SpatialEncoder.forward emits): handed over as is and borrowed by the library -- no
            # re-layout pass, no second copy in HBM; an NCHW latent (reference layout, e.g. set_scene in tests) is transposed once
            self._latent_keepalive = lat if nhwc else None
            self._ctx.set_option("latent_layout", 2 if nhwc else 0)
            self._ctx.set_scene(lat if nhwc else lat.contiguous(), f32(enc.depths), f32(enc.depths_std), f32(enc.normals),
                                f32(self.poses), f32(self.focal), f32(self.c), enc.feature_padding,
                                self.poscode.num_freqs, self.poscode.freq_factor)
            self._scene_stamp = sstamp
        return self._ctx

    def mode_id(self):
        if self.mode not in capi.MODES:
            raise ValueError("mode must be one of %s" % list(capi.MODES))
        if float(getattr(self.mlp_fine, "beta", 0.0) or 0.0) > 0:
            return capi.MODE_FP32       # Softplus networks are served by the fp32 CUDA-core kernels (the tcgen05 epilogues are ReLU)
        return capi.MODES[self.mode]

    def _no_grad_only(self, *tensors):
        if torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or
                                        any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError(
                "libdiner_b200 implements the forward render path only (backward is a 'next' row, SURVEY §8(f)); "
                "call under torch.no_grad()")

    def forward(self, xyz, viewdirs):
        """(SB,B,3) world points and view directions -> (SB,B,4) [sigmoid rgb, relu sigma]."""
        self._no_grad_only(xyz, viewdirs)
        assert xyz.shape[0] == self.encoder.nobjects
        return self.context().query(xyz.float().contiguous(), viewdirs.float().contiguous(), self.mode_id())

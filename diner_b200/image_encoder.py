"""SpatialEncoder with the reference's constructor, state_dict (`model.*` = torchvision ResNet-34,
`positional_encoding._freqs/_phases`) and stored scene tensors (latent, depths, depths_std, normals,
nviews, nobjects) -- reference src/models/image_encoder.py:19-95,225-291.

The ResNet trunk runs once per scene on cuDNN and is outside the hot path.  The
per-sample lookups the reference does through `index*` (image_encoder.py:97-223) happen inside the
fused CUDA kernels; the `index*` methods here keep the public API for external callers."""
import functools

import numpy as np
import torch
import torch.nn.functional as F
import torchvision
from torch import nn

from .positional_encoding import PositionalEncoding

STD_PAD, STD_DOUBLE_WIDTH = 100, 12     # image_encoder.py:193-194


class SpatialEncoder(nn.Module):
    def __init__(self, backbone="resnet34", pretrained=True, num_layers=4, index_interp="bilinear",
                 index_padding="border", upsample_interp="bilinear", use_first_pool=True, image_padding=0,
                 padding_pe=-1):
        super().__init__()
        norm_layer = functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
        weights = "DEFAULT" if pretrained else None
        self.model = getattr(torchvision.models, backbone)(weights=weights, norm_layer=norm_layer)
        self.model.fc = nn.Sequential()
        self.model.avgpool = nn.Sequential()
        self.use_first_pool = use_first_pool
        self.num_layers = num_layers
        self.latent_size = [0, 64, 128, 256, 512, 1024][num_layers]
        self.image_padding = image_padding
        self.feature_padding = image_padding / self.model.conv1.stride[0]
        assert self.feature_padding % 1 == 0
        self.pad_layer = nn.ReplicationPad2d([image_padding] * 4)
        self.padding_pe = padding_pe
        self.positional_encoding = None
        if padding_pe >= 0 and self.feature_padding != 0:
            # extra input channels carrying a 2-D positional code on the padded border (image_encoder.py:63-83)
            self.positional_encoding = PositionalEncoding(padding_pe, freq_factor=np.pi, d_in=2, include_input=True)
            old = self.model.conv1
            new = nn.Conv2d(old.in_channels + self.positional_encoding.d_out, old.out_channels,
                            kernel_size=old.kernel_size, stride=old.stride, padding=old.padding, bias=old.bias,
                            dilation=old.dilation, padding_mode=old.padding_mode, groups=old.groups)
            nn.init.kaiming_normal_(new.weight, mode="fan_out", nonlinearity="relu")
            with torch.no_grad():
                new.weight[:, :old.weight.shape[1]] = old.weight
            self.model.conv1 = new
        self.index_interp, self.index_padding, self.upsample_interp = index_interp, index_padding, upsample_interp
        self.register_buffer("latent", torch.empty(1, 1, 1, 1), persistent=False)
        self.nviews = self.nobjects = None
        self.scene_version = 0          # bumped by forward(); PixelNeRF re-uploads the scene when it changes

    # ------------------------------------------------------------------------------------------
    def forward(self, imgs, depths, depths_std, normals):
        """imgs (SB,NV,3,H,W) -> stores latent (SB,NV,L,Hl,Wl) and the depth / std / normal maps."""
        SB, NV, Cin, H, W = imgs.shape
        self.depths, self.depths_std, self.normals = depths, depths_std, normals
        self.nviews, self.nobjects = NV, SB
        x = self.pad_layer(imgs.reshape(SB * NV, Cin, H, W))
        p = self.image_padding
        if self.positional_encoding is not None:
            gy, gx = torch.meshgrid(torch.linspace(-1, 1, H + 2 * p, device=x.device),
                                    torch.linspace(-1, 1, W + 2 * p, device=x.device), indexing="ij")
            pe = self.positional_encoding(torch.stack((gx, gy), dim=-1))
            pe[p:-p, p:-p] = 0
            x = torch.cat((x, pe.permute(2, 0, 1).unsqueeze(0).expand(SB * NV, -1, -1, -1)), dim=1)
        m = self.model
        x = m.relu(m.bn1(m.conv1(x)))
        pyramid = [x]
        if self.num_layers > 1:
            x = m.layer1(m.maxpool(x) if self.use_first_pool else x)
            pyramid.append(x)
        for i, layer in enumerate((m.layer2, m.layer3, m.layer4)):
            if self.num_layers > i + 2:
                x = layer(x)
                pyramid.append(x)
        size = pyramid[0].shape[-2:]
        ac = None if self.index_interp == "nearest " else True     # (sic) image_encoder.py:281
        # The reference concatenates the upsampled levels into an NCHW tensor (image_encoder.py:282-291).  Same values here, but the
        # levels are written straight into channels-last storage -- the layout the render kernels gather from -- so no separate
        # NCHW->NHWC pass over the 0.8-5.4 GB latent is needed afterwards.  `self.latent` keeps the reference's
        # logical shape (SB,NV,L,Hl,Wl); only its strides differ.
        L = sum(t.shape[1] for t in pyramid)
        buf = torch.empty(SB * NV, size[0], size[1], L, device=x.device, dtype=pyramid[0].dtype)
        c0 = 0
        for t in pyramid:
            up = F.interpolate(t, size, mode=self.upsample_interp, align_corners=ac)
            buf[..., c0:c0 + up.shape[1]] = up.permute(0, 2, 3, 1)
            c0 += up.shape[1]
        self.latent = buf.permute(0, 3, 1, 2).view(SB, NV, L, *size)
        self.scene_version += 1
        return self.latent

    def set_scene(self, latent, depths, depths_std, normals):
        """Installs externally computed feature maps (tests / synthetic benchmark: no ResNet)."""
        SB, NV = latent.shape[:2]
        self.latent, self.depths, self.depths_std, self.normals = latent, depths, depths_std, normals
        self.nviews, self.nobjects = NV, SB
        self.scene_version += 1

    # ---- public lookup API of the reference (not used by the fused render path) -----------------
    def _grid(self, t, uv):
        SB, NV, N, _ = uv.shape
        return t.reshape(SB * NV, *t.shape[-3:]), uv.reshape(SB * NV, N, 1, 2), (SB, NV, N)

    def index(self, uv):
        assert uv.shape[:2] == self.latent.shape[:2]
        size = torch.tensor([self.latent.shape[-1], self.latent.shape[-2]], device=uv.device)
        uv = uv * ((size - self.feature_padding * 2) / size).view(1, 1, 1, 2)
        t, g, (SB, NV, N) = self._grid(self.latent, uv)
        s = F.grid_sample(t, g, align_corners=False, mode=self.index_interp, padding_mode=self.index_padding)
        return s[..., 0].view(SB, NV, -1, N)

    def index_depth(self, uv):
        assert uv.shape[:2] == self.depths.shape[:2]
        t, g, (SB, NV, N) = self._grid(self.depths, uv)
        return F.grid_sample(t, g, align_corners=False, mode="nearest", padding_mode="border")[..., 0].view(SB, NV, -1, N)

    def index_depth_std(self, uv):
        assert uv.shape[:2] == self.depths_std.shape[:2]
        t, g, (SB, NV, N) = self._grid(self.depths_std, uv)
        H, W = t.shape[-2:]
        ring = torch.arange(STD_PAD - 1, -1, -1, device=t.device, dtype=t.dtype)
        ry = torch.cat((ring, torch.zeros(H, device=t.device), ring.flip(0))).view(-1, 1)
        rx = torch.cat((ring, torch.zeros(W, device=t.device), ring.flip(0))).view(1, -1)
        gain = torch.exp(torch.maximum(ry, rx) / STD_DOUBLE_WIDTH * np.log(2))
        padded = F.pad(t, [STD_PAD] * 4, mode="replicate") * gain
        sz = torch.tensor([W, H], dtype=torch.float, device=t.device)
        g = g * (sz / (sz + 2 * STD_PAD)).view(1, 1, 1, 2)
        return F.grid_sample(padded, g, mode="nearest", padding_mode="zeros", align_corners=False)[..., 0].view(SB, NV, -1, N)

    def index_normal(self, uv):
        assert uv.shape[:2] == self.normals.shape[:2]
        t, g, (SB, NV, N) = self._grid(self.normals, uv)
        return F.grid_sample(t, g, align_corners=False, mode="nearest", padding_mode="zeros")[..., 0].view(SB, NV, -1, N)

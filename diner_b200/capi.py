"""ctypes binding of libdiner_b200.so (include/diner_b200.h).

PyTorch is only the owner of device memory and streams here: every call passes raw `data_ptr()`s and
the current CUDA stream handle.  There is no fallback: if the shared library is missing or a call
fails, a RuntimeError carrying diner_last_error() is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiner_b200.so")

MODE_FP32, MODE_PARITY, MODE_FAST = 0, 1, 2
MODES = {"fp32": MODE_FP32, "parity": MODE_PARITY, "fast": MODE_FAST}

_c_float_p = ctypes.c_void_p


class DinerNoise(ctypes.Structure):
    _fields_ = [("u_coarse", ctypes.c_void_p), ("g_noise", ctypes.c_void_p), ("u_fill", ctypes.c_void_p),
                ("seed", ctypes.c_uint64), ("ray_offset", ctypes.c_uint64)]


# name -> (restype, argtypes); mirrors include/diner_b200.h one to one
_I, _LL, _F, _P, _U64 = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p, ctypes.c_uint64
_PP = ctypes.POINTER(ctypes.c_void_p)
SIGNATURES = {
    "diner_last_error": (ctypes.c_char_p, []),
    "diner_version": (_I, []),
    "diner_create": (_I, [_PP, _I]),
    "diner_destroy": (None, [_P]),
    "diner_set_mlp": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _PP, _PP, _PP, _PP, _PP, _PP, _P]),
    "diner_set_scene": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _F, _I, _F, _P]),
    "diner_render": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, ctypes.POINTER(DinerNoise), _P, _P, _P, _P, _P]),
    "diner_render_rgbd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, ctypes.POINTER(DinerNoise), _P, _P]),
    "diner_render_image": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _I, _I, _I, _I, _I, ctypes.POINTER(DinerNoise), _P, _P, _P]),
    "diner_gen_rays": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _P, _P]),
    "diner_depth2normal": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "diner_render_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "diner_mlp_param_count": (_LL, [_P]),
    "diner_render_host": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _U64, _P, _P, _P]),
    "diner_sample": (_I, [_P, _P, _I, _I, _I, _I, _I, ctypes.POINTER(DinerNoise), _P, _P, _P]),
    "diner_query": (_I, [_P, _P, _P, _I, _LL, _I, _P, _P]),
    "diner_composite": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "diner_set_option": (_I, [_P, ctypes.c_char_p, _LL]),
    "diner_set_float_option": (_I, [_P, ctypes.c_char_p, ctypes.c_double]),
    "diner_debug_sync": (_I, [_P]),
    "diner_launch_count": (_LL, [_P]),
    "diner_set_timing": (_I, [_P, _I]),
    "diner_last_mlp_ms": (_F, [_P]),
    "diner_last_stage_ms": (_F, [_P, _I]),
}

_lib = None


def load_library():
    """Loads libdiner_b200.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libdiner_b200.so not found at %s -- run `python __graft_entry__.py` (build()) first; "
                           "there is no CPU or PyTorch fallback for the render path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the header and the binary disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(t, shape=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (the render path has no CPU implementation)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Context:
    """One libdiner_b200 context = one (MLP, encoded scene) pair on one device."""

    def __init__(self, device):
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("diner_b200 needs a CUDA device, got %s" % device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        h = ctypes.c_void_p()
        self._check(self.lib.diner_create(ctypes.byref(h), idx))
        self.handle = h
        self._keep = []
        for key, env in (("ray_image_width", "DINER_RAY_IMAGE_WIDTH"), ("backward_tc", "DINER_B200_BACKWARD_TC"), ("fused", "DINER_TC_FUSED"), ("post_tiles", "DINER_TC_POST_TILES"), ("tail_kb", "DINER_TC_TAIL_KB"), ("sub_batch", "DINER_TC_SUB_BATCH"),
                         ("dbg_skip", "DINER_TC_DBG_SKIP"), ("early_split", "DINER_TC_EARLY_SPLIT"), ("early_lin", "DINER_TC_EARLY_LIN"), ("warm_rounds", "DINER_TC_WARM_ROUNDS")):
            if os.environ.get(env):
                self.set_option(key, int(os.environ[env]))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("libdiner_b200 error %d: %s" % (rc, self.lib.diner_last_error().decode()))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.diner_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters / scene -----------------------------------------------------------------
    def set_mlp(self, sd, d_in, d_latent, d_hidden, d_out, n_blocks, combine_layer):
        """sd: dict name -> CUDA fp32 tensor with the reference ResnetFC state_dict keys."""
        nz = min(combine_layer, n_blocks)

        def arr(fmt, n):
            a = (ctypes.c_void_p * max(n, 1))()
            for i in range(n):
                a[i] = sd[fmt % i].data_ptr()
                _ptr(sd[fmt % i], name=fmt % i)
            return a

        with torch.cuda.device(self.device):
            self._check(self.lib.diner_set_mlp(
                self.handle, d_in, d_latent, d_hidden, d_out, n_blocks, combine_layer,
                _ptr(sd["lin_in.weight"], (d_hidden, d_in), "lin_in.weight"), _ptr(sd["lin_in.bias"], (d_hidden,)),
                _ptr(sd["lin_out.weight"], (d_out, d_hidden), "lin_out.weight"), _ptr(sd["lin_out.bias"], (d_out,)),
                arr("blocks.%d.fc_0.weight", n_blocks), arr("blocks.%d.fc_0.bias", n_blocks),
                arr("blocks.%d.fc_1.weight", n_blocks), arr("blocks.%d.fc_1.bias", n_blocks),
                arr("lin_z.%d.weight", nz), arr("lin_z.%d.bias", nz), _stream(self.device)))

    def set_scene(self, latent, depths, depths_std, normals, poses, focal, c, feature_padding, num_freqs,
                  freq_factor):
        SB, NV, L, Hl, Wl = latent.shape
        H, W = depths.shape[-2:]
        if not latent.is_contiguous():          # channels-last storage (option latent_layout = 1 | 2): check that layout instead
            if not latent.permute(0, 1, 3, 4, 2).is_contiguous():
                raise RuntimeError("latent must be contiguous NCHW or channels-last")
            lat_ptr = _ptr(latent.permute(0, 1, 3, 4, 2), name="latent")
        else:
            lat_ptr = _ptr(latent, name="latent")
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_set_scene(
                self.handle, SB, NV, L, Hl, Wl, H, W, lat_ptr,
                _ptr(depths, (SB, NV, 1, H, W), "depths"), _ptr(depths_std, (SB, NV, 1, H, W), "depths_std"),
                _ptr(normals, (SB, NV, 3, H, W), "normals"), _ptr(poses, (SB, NV, 4, 4), "poses"),
                _ptr(focal, (SB, NV, 2), "focal"), _ptr(c, (SB, NV, 2), "c"), float(feature_padding),
                int(num_freqs), float(freq_factor), _stream(self.device)))

    # ---- render path -------------------------------------------------------------------------
    @staticmethod
    def _noise(noise, SB, NR, K, C, G):
        if noise is None:
            return None, ()
        n = DinerNoise()
        keep = []
        for key, shape in (("u_coarse", (SB, NR, C)), ("g_noise", (SB, NR, G)), ("u_fill", (SB, NR, K))):
            t = noise.get(key)
            if t is not None and t.numel() > 0:
                setattr(n, key, _ptr(t, shape, key).value)
                keep.append(t)
        n.seed = int(noise.get("seed", 0))
        n.ray_offset = int(noise.get("ray_offset", 0))
        return n, keep

    def render(self, rays, K, C, G, white_bkgd, mode, noise=None, want_weights=False, want_z=False):
        SB, NR, _ = rays.shape
        dev = rays.device
        rgb = torch.empty(SB, NR, 3, device=dev)
        depth = torch.empty(SB, NR, device=dev)
        w = torch.empty(SB, NR, K, device=dev) if want_weights else None
        z = torch.empty(SB, NR, K, device=dev) if want_z else None
        n, keep = self._noise(noise, SB, NR, K, C, G)
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_render(
                self.handle, _ptr(rays, (SB, NR, 8), "rays"), SB, NR, K, C, G, int(bool(white_bkgd)), mode,
                ctypes.byref(n) if n is not None else None, _ptr(rgb), _ptr(depth), _ptr(w), _ptr(z),
                _stream(self.device)))
        return rgb, depth, w, z

    def render_rgbd(self, rays, K, C, G, white_bkgd, mode, noise=None, out=None):
        """diner_render_rgbd: rays (SB,NR,8) -> packed (SB,NR,4) [r,g,b,depth]; `out` may be a contiguous (SB,NR,4) view of a
        larger buffer (the rank's slice of the image all-gather)."""
        SB, NR, _ = rays.shape
        if out is None:
            out = torch.empty(SB, NR, 4, device=rays.device)
        n, keep = self._noise(noise, SB, NR, K, C, G)
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_render_rgbd(
                self.handle, _ptr(rays, (SB, NR, 8), "rays"), SB, NR, K, C, G, int(bool(white_bkgd)), mode,
                ctypes.byref(n) if n is not None else None, _ptr(out, (SB, NR, 4), "rgbd"), _stream(self.device)))
        return out

    def render_image(self, target_extrinsics, target_intrinsics, H, W, z_near, z_far, K, C, G, white_bkgd, mode, noise=None):
        """gen_rays + the ray-batch loop of DINER.predict_imgs_from_batch (diner.py:79-92) in one library call.
        Returns rgb (SB,H*W,3), depth (SB,H*W); noise may only carry a seed (dense noise is per ray batch)."""
        SB = target_extrinsics.shape[0]
        dev = target_extrinsics.device
        rgb = torch.empty(SB, H * W, 3, device=dev)
        depth = torch.empty(SB, H * W, device=dev)
        n, keep = self._noise(noise, SB, H * W, K, C, G)
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_render_image(
                self.handle, _ptr(target_extrinsics, (SB, 4, 4), "target_extrinsics"),
                _ptr(target_intrinsics, (SB, 3, 3), "target_intrinsics"), SB, H, W, float(z_near), float(z_far), K, C, G,
                int(bool(white_bkgd)), mode, ctypes.byref(n) if n is not None else None, _ptr(rgb), _ptr(depth),
                _stream(self.device)))
        return rgb, depth

    def gen_rays(self, target_extrinsics, target_intrinsics, H, W, z_near, z_far):
        """src/util/cam_geometry.py:5-48 on the device: (SB,H*W,8)."""
        SB = target_extrinsics.shape[0]
        rays = torch.empty(SB, H * W, 8, device=target_extrinsics.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_gen_rays(
                self.handle, _ptr(target_extrinsics, (SB, 4, 4), "target_extrinsics"),
                _ptr(target_intrinsics, (SB, 3, 3), "target_intrinsics"), SB, H, W, float(z_near), float(z_far),
                _ptr(rays), _stream(self.device)))
        return rays

    def depth2normal(self, depths, intrinsics):
        """src/util/depth2normal.py:6-87 on the device: depths (N,1,H,W), intrinsics (N,3,3) -> (N,3,H,W)."""
        N, _, H, W = depths.shape
        normals = torch.empty(N, 3, H, W, device=depths.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_depth2normal(self.handle, _ptr(depths, (N, 1, H, W), "depths"),
                                                    _ptr(intrinsics, (N, 3, 3), "intrinsics"), N, H, W, _ptr(normals),
                                                    _stream(self.device)))
        return normals

    def render_backward(self, rays, z, white_bkgd, g_rgb, g_depth=None, want_latent_grad=True, latent_shape=None):
        """diner_render_backward (see include/diner_b200.h): returns (flat parameter gradients in diner_set_mlp order, d_latent NCHW)."""
        SB, NR, K = z.shape
        n = int(self.lib.diner_mlp_param_count(self.handle))
        gp = torch.zeros(n, device=z.device)
        dl = torch.zeros(latent_shape, device=z.device) if want_latent_grad else None
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_render_backward(
                self.handle, _ptr(rays, (SB, NR, 8), "rays"), _ptr(z, name="z"), SB, NR, K, int(bool(white_bkgd)),
                _ptr(g_rgb, (SB, NR, 3), "g_rgb"), _ptr(g_depth, (SB, NR), "g_depth") if g_depth is not None else None,
                _ptr(gp), _ptr(dl) if dl is not None else None, _stream(self.device)))
        return gp, dl

    def sample(self, rays, K, C, G, noise=None, want_dgs=False):
        SB, NR, _ = rays.shape
        z = torch.empty(SB, NR, K, device=rays.device)
        zd = torch.empty(SB, NR, K, device=rays.device) if want_dgs else None
        n, keep = self._noise(noise, SB, NR, K, C, G)
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_sample(self.handle, _ptr(rays, (SB, NR, 8), "rays"), SB, NR, K, C, G,
                                              ctypes.byref(n) if n is not None else None, _ptr(z), _ptr(zd),
                                              _stream(self.device)))
        return (z, zd) if want_dgs else z

    def query(self, xyz, viewdirs, mode):
        SB, B, _ = xyz.shape
        out = torch.empty(SB, B, 4, device=xyz.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_query(self.handle, _ptr(xyz, (SB, B, 3), "xyz"),
                                             _ptr(viewdirs, (SB, B, 3), "viewdirs"), SB, B, mode, _ptr(out),
                                             _stream(self.device)))
        return out

    def composite(self, rays, z, white_bkgd, mode, want_weights=True):
        SB, NR, K = z.shape
        rgb = torch.empty(SB, NR, 3, device=z.device)
        depth = torch.empty(SB, NR, device=z.device)
        w = torch.empty(SB, NR, K, device=z.device) if want_weights else None
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_composite(self.handle, _ptr(rays, (SB, NR, 8), "rays"), _ptr(z, name="z"),
                                                 SB, NR, K, int(bool(white_bkgd)), mode, _ptr(rgb), _ptr(depth),
                                                 _ptr(w), _stream(self.device)))
        return w, rgb, depth

    def render_host(self, rays_host, K, C, G, white_bkgd, mode, seed, rgb_host, depth_host):
        """Host-buffer entry (pinned CPU tensors): H2D + render + D2H + stream sync inside the call."""
        SB, NR, _ = rays_host.shape
        for t in (rays_host, rgb_host, depth_host):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("render_host needs contiguous float32 CPU tensors")
        with torch.cuda.device(self.device):
            self._check(self.lib.diner_render_host(
                self.handle, ctypes.c_void_p(rays_host.data_ptr()), SB, NR, K, C, G, int(bool(white_bkgd)), mode,
                int(seed), ctypes.c_void_p(rgb_host.data_ptr()), ctypes.c_void_p(depth_host.data_ptr()),
                _stream(self.device)))

    def set_option(self, key, value):
        self._check(self.lib.diner_set_option(self.handle, key.encode(), int(value)))

    def set_float_option(self, key, value):
        self._check(self.lib.diner_set_float_option(self.handle, key.encode(), float(value)))

    def debug_sync(self):
        self._check(self.lib.diner_debug_sync(self.handle))

    def launch_count(self):
        return int(self.lib.diner_launch_count(self.handle))

    def set_timing(self, enabled):
        self._check(self.lib.diner_set_timing(self.handle, int(enabled)))

    def last_stage_ms(self):
        """dict of device ms of the last call's stages (needs set_timing(True))."""
        names = ("sampler", "mlp_pre", "mlp_post", "composite", "lin_z_maps")
        return {n: float(self.lib.diner_last_stage_ms(self.handle, i)) for i, n in enumerate(names)}

    def last_mlp_ms(self):
        return float(self.lib.diner_last_mlp_ms(self.handle))

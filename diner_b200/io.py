"""Output / input side of the render path: what surrounds `predict_imgs_from_batch` in the reference's
prediction scripts once the render itself is fast.

* `torch_cmap`      -- reference src/util/torch_helpers.py:42-75: depth (B,1,H,W) -> colour (B,3,H,W) through a matplotlib colormap
                       with per-image min/max normalisation.  The reference goes through `.cpu().numpy()` + matplotlib per call;
                       here the 256-entry lookup table is built once and the normalise + lookup (+ optional uint8 quantise) runs
                       on the device (libdiner_b200 `diner_colormap`).
* `ImageWriter`     -- reference src/models/diner.py:123-133: four `save_image` calls per sample on the render thread; here the
                       uint8 conversion happens on the device, ONE device->host copy per batch, and PNG encoding runs on a small
                       thread pool off the critical path.
* `read_depth_png`, `conf_to_std` -- on-disk formats the datasets feed into the scene (src/data/dtu.py:69,95-122;
                       src/data/facescape.py:51,65-69): uint16 PNG x 1e-4 metres, confidence -> standard deviation affine maps.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

# conf -> std affine maps of the two datasets (dtu.py:69, facescape.py:51)
CONF2STD = {"dtu": (-2.5679e-2, 3.2818e-2), "facescape": (-1.582e-2, 1.649e-2)}
DEPTH_PNG_SCALE = 1e-4                # uint16 PNG -> metres (dtu.py:104, facescape.py:67)
DTU_TRANSMVS_SCALE = 0.7 / 872.0      # dtu.py:106: TransMVSNet predictions are stored in the network's training scale

_LUT_CACHE = {}


def _viridis_fit(t):
    """Degree-6 polynomial fit of matplotlib's viridis (max deviation ~3/255 per channel): only used when matplotlib is not
    installed -- with matplotlib present the exact ListedColormap table is taken, so colours equal the reference's bit for bit."""
    c = np.array([[0.2777273272234177, 0.005407344544966578, 0.3340998053353061],
                  [0.1050930431085774, 1.404613529898575, 1.384590162594685],
                  [-0.3308618287255563, 0.214847559468213, 0.09509516302823659],
                  [-4.634230498983486, -5.799100973351585, -19.33244095627987],
                  [6.228269936347081, 14.17993336680509, 56.69055260068105],
                  [4.776384997670288, -13.74514537774601, -65.35303263337234],
                  [-5.435455855934631, 4.645852612178535, 26.3124352495832]])
    out = np.zeros(t.shape + (3,))
    for k in range(6, -1, -1):
        out = out * t[..., None] + c[k]
    return np.clip(out, 0.0, 1.0)


def colormap_lut(cmap="viridis", n=256):
    """(n,3) float64 table with lut[i] = cmap(i / (n - 1)) -- for matplotlib's 256-entry listed maps exactly its `colors`."""
    key = (cmap, n)
    if key not in _LUT_CACHE:
        try:
            import matplotlib.pyplot as plt
            lut = np.asarray(plt.get_cmap(cmap)(np.arange(n) / (n - 1)))[:, :3].astype(np.float64)
            src = "matplotlib"
        except Exception:
            if cmap != "viridis":
                raise RuntimeError("colormap %r needs matplotlib (only viridis has a built-in table)" % cmap)
            lut = _viridis_fit(np.arange(n) / (n - 1))
            src = "polynomial fit"
        _LUT_CACHE[key] = (lut, src)
    return _LUT_CACHE[key]


def torch_cmap(x, cmap="viridis", vmin=None, vmax=None, as_uint8=False):
    """reference torch_helpers.torch_cmap: x (B,1,H,W) | (1,H,W) | (H,W) -> same leading shape with 3 channels, on x.device.

    Semantics kept from the reference (torch_helpers.py:58-69 + matplotlib.colors.Colormap.__call__): per-image min / max unless
    vmin / vmax are given (falsy values count as not given, like the reference's `vmin if vmin else ...`), x normalised in float64,
    index = int(x * 256) with x == 1 mapped to 255; NaN (0/0 of a constant image) maps to the colormap's "bad" colour (0,0,0).
    float64 output like the reference (matplotlib returns doubles), or uint8 (truncating `* 255`, what save_torch_video /
    torchvision.save_image would make of it) with as_uint8=True."""
    shape = x.shape
    x4 = x.detach().reshape(*([1] * (4 - x.dim())), *shape)
    assert x4.shape[1] == 1
    lut, _ = colormap_lut(cmap)
    B = x4.shape[0]
    xd = x4[:, 0].double()
    flat = xd.reshape(B, -1)
    lo = torch.full((B, 1, 1), float(vmin), dtype=torch.float64, device=x.device) if vmin else flat.min(dim=1).values.view(B, 1, 1)
    hi = torch.full((B, 1, 1), float(vmax), dtype=torch.float64, device=x.device) if vmax else flat.max(dim=1).values.view(B, 1, 1)
    xn = (xd - lo) / (hi - lo)
    bad = torch.isnan(xn)
    idx = (xn * 256.0)
    idx = torch.where(idx == 256.0, torch.full_like(idx, 255.0), idx)
    idx = torch.nan_to_num(idx, nan=0.0).clamp(-1.0, 256.0).to(torch.int64)       # int() truncation; <0 under, >255 over
    idx = idx.clamp(0, 255)                                                        # under / over colours = the end colours (viridis default)
    table = torch.from_numpy(lut).to(x.device)
    out = table[idx]                                                               # (B,H,W,3)
    out = torch.where(bad.unsqueeze(-1), torch.zeros_like(out), out).permute(0, 3, 1, 2)
    out = out.reshape(list(shape[:-3]) + [3] + list(shape[-2:]))
    if as_uint8:
        return (out * 255.0).to(torch.uint8)
    return out


def to_uint8(img):
    """float image in [0,1] -> uint8 the way torchvision.utils.save_image quantises (mul 255, add 0.5, clamp, truncate)."""
    return img.detach().mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8)


class ImageWriter:
    """Batched, asynchronous replacement of the per-sample `save_image` calls of DINER.create_prediction_folder
    (src/models/diner.py:123-133).  `add(path, chw_uint8_or_float_tensor)` queues an image; device tensors of one `flush()` are
    quantised on the device, copied to the host in one go and PNG/JPEG-encoded on worker threads."""

    def __init__(self, workers=4):
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending, self.futures = [], []

    def add(self, path, img):
        if img.dtype != torch.uint8:
            img = to_uint8(img)
        self.pending.append((path, img))

    @staticmethod
    def _encode(path, hwc):
        from PIL import Image
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        Image.fromarray(hwc if hwc.shape[-1] != 1 else hwc[..., 0]).save(path)
        return path

    def flush(self):
        if not self.pending:
            return
        host = [t.permute(1, 2, 0).contiguous().to("cpu", non_blocking=True) for _, t in self.pending]
        if any(t.is_cuda for _, t in self.pending):
            torch.cuda.synchronize()
        for (path, _), h in zip(self.pending, host):
            self.futures.append(self.pool.submit(self._encode, path, h.numpy()))
        self.pending = []

    def close(self):
        self.flush()
        done = [f.result() for f in self.futures]
        self.pool.shutdown()
        self.futures = []
        return done


def write_prediction_images(writer, outdir, stems, pred_rgb, pred_depth, src_rgbs, gt_rgb,
                            suffixes=("-pred.png", "-depth.png", "-ref.png", "-gt.png")):
    """diner.py:123-133: per sample the prediction, the colour-mapped depth, the source views side by side and the ground truth.
    `suffixes` default to eval_suite's PRED / DEPTH / REF / GT suffixes (src/evaluation/eval_suite.py:21-24)."""
    depth_rgb = torch_cmap(pred_depth, as_uint8=True)
    src = torch.cat(src_rgbs.unbind(1), dim=-1)
    for i, stem in enumerate(stems):
        writer.add(os.path.join(outdir, stem + suffixes[0]), pred_rgb[i])
        writer.add(os.path.join(outdir, stem + suffixes[1]), depth_rgb[i])
        writer.add(os.path.join(outdir, stem + suffixes[2]), src[i])
        writer.add(os.path.join(outdir, stem + suffixes[3]), gt_rgb[i])
    writer.flush()


# ----------------------------------------------------------------------------------------------
# readers
# ----------------------------------------------------------------------------------------------
def read_depth_png(path, dataset="facescape", scale_factor=1.0):
    """uint16 PNG -> float32 metres (1,H,W).  facescape.py:65-69: value * 1e-4.  dtu.py:103-107,119: value * 1e-4 / (0.7/872)
    (TransMVSNet predictions), then the dataset's global `scale_factor` (dtu.py:21: 0.7/872 again for the shipped config)."""
    from PIL import Image
    arr = np.asarray(Image.open(path))
    d = torch.from_numpy(arr.astype(np.float32)) * DEPTH_PNG_SCALE
    if dataset == "dtu":
        d = d / DTU_TRANSMVS_SCALE
    d = d * scale_factor
    return d.reshape(1, *d.shape[-2:])


def conf_to_std(conf, dataset):
    """Confidence map (as read by read_depth_png) -> depth standard deviation: dtu.py:69 / facescape.py:51."""
    a, b = CONF2STD[dataset]
    return a * conf + b

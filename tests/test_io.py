"""Output / input side: colormap semantics of the reference's torch_cmap, the image writer and the
depth / confidence readers -- CPU tests (the functions are device-agnostic torch code; the GPU suite runs them on cuda)."""
import os

import numpy as np
import pytest
import torch

from diner_b200 import io as IO
from diner_b200 import synthetic as S


def _reference_cmap(x, lut, vmin=None, vmax=None):
    """torch_helpers.py:58-69 restated in numpy with matplotlib's Colormap.__call__ index rule (N = 256)."""
    x = x.astype(float)
    vmin = vmin if vmin else np.min(x.reshape(x.shape[0], -1), axis=-1).reshape((-1, 1, 1, 1))
    vmax = vmax if vmax else np.max(x.reshape(x.shape[0], -1), axis=-1).reshape((-1, 1, 1, 1))
    with np.errstate(invalid="ignore", divide="ignore"):
        xn = ((x - vmin) / (vmax - vmin))[:, 0]
        xa = xn * 256
        xa[xa == 256] = 255
        under, over, bad = xa < 0, xa >= 256, np.isnan(xa)
        xi = xa.astype(int)
    xi[under], xi[over], xi[bad] = 0, 255, 0
    out = lut[np.clip(xi, 0, 255)]
    out[bad] = 0.0
    return out.transpose(0, 3, 1, 2)


@pytest.mark.parametrize("device", ["cpu"] + (["cuda"] if torch.cuda.is_available() else []))
def test_torch_cmap_matches_reference_semantics(device):
    lut, _ = IO.colormap_lut("viridis")
    assert lut.shape == (256, 3) and lut.min() >= 0 and lut.max() <= 1
    x = S.hash_uniform((3, 1, 17, 23), 5) * 2.5
    x[1] = 0.75                                   # constant image: 0/0 -> NaN -> "bad" colour
    x[2, 0, 0, 0] = 0.0
    got = IO.torch_cmap(x.to(device))
    assert got.shape == (3, 3, 17, 23) and got.dtype == torch.float64
    assert np.array_equal(got.cpu().numpy(), _reference_cmap(x.numpy(), lut))
    got2 = IO.torch_cmap(x.to(device), vmin=0.5, vmax=2.0)              # explicit range: under / over colours
    assert np.array_equal(got2.cpu().numpy(), _reference_cmap(x.numpy(), lut, 0.5, 2.0))
    assert IO.torch_cmap(x[0, 0].to(device)).shape == (3, 17, 23)       # (H,W) input like the reference allows
    u8 = IO.torch_cmap(x.to(device), as_uint8=True)
    assert u8.dtype == torch.uint8 and torch.equal(u8.cpu(), (got.cpu() * 255).to(torch.uint8))


def test_viridis_fit_is_close_to_the_known_anchor_colours():
    """Without matplotlib the table is a polynomial fit; its ends / middle must be viridis (values from matplotlib's table)."""
    lut, src = IO.colormap_lut("viridis")
    for i, rgb in ((0, (0.267004, 0.004874, 0.329415)), (127, (0.128729, 0.563265, 0.551229)), (255, (0.993248, 0.906157, 0.143936))):
        assert np.abs(lut[i] - np.array(rgb)).max() < (1e-6 if src == "matplotlib" else 0.02), (i, lut[i], src)


def test_image_writer_and_depth_reader_roundtrip(tmp_path):
    from PIL import Image
    w = IO.ImageWriter(workers=2)
    rgb = S.hash_uniform((2, 3, 12, 20), 3)
    depth = S.hash_uniform((2, 1, 12, 20), 4) + 1.0
    src = S.hash_uniform((2, 4, 3, 12, 20), 6)
    IO.write_prediction_images(w, str(tmp_path), ["a", "b"], rgb, depth, src, rgb)
    files = w.close()
    assert len(files) == 8 and all(os.path.exists(f) for f in files)
    back = torch.from_numpy(np.asarray(Image.open(tmp_path / "a-pred.png"))).permute(2, 0, 1)
    assert torch.equal(back, IO.to_uint8(rgb[0]))                                     # == torchvision.utils.save_image quantisation
    assert np.asarray(Image.open(tmp_path / "b-ref.png")).shape == (12, 80, 3)        # 4 source views side by side
    # uint16 depth PNG (x 1e-4 m) and the confidence -> std maps
    raw = (np.arange(12 * 20, dtype=np.uint16).reshape(12, 20) * 37)
    Image.fromarray(raw).save(tmp_path / "d.png")
    d = IO.read_depth_png(tmp_path / "d.png", "facescape")
    assert d.shape == (1, 12, 20) and torch.allclose(d[0], torch.from_numpy(raw.astype(np.float32)) * 1e-4)
    d2 = IO.read_depth_png(tmp_path / "d.png", "dtu", scale_factor=0.7 / 872.0)
    assert torch.allclose(d2, d, rtol=1e-6)                                           # dtu.py:106 and :119 cancel for the shipped scale
    c = torch.tensor([0.0, 0.5, 1.0])
    assert torch.allclose(IO.conf_to_std(c, "dtu"), torch.tensor([3.2818e-2, 3.2818e-2 - 1.28395e-2, 3.2818e-2 - 2.5679e-2]))
    assert torch.allclose(IO.conf_to_std(c, "facescape"), -1.582e-2 * c + 1.649e-2)

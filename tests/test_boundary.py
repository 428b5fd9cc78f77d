"""Drop-in boundary: import paths, constructor signatures, state_dict keys, C-ABI exports, loud failures."""
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    from src.models.pixelnerf import PixelNeRF
    return PixelNeRF(
        poscode_conf=dict(kwargs=dict(num_freqs=6, freq_factor=6.28, include_input=True)),
        encoder_conf=dict(module="src.models.image_encoder.SpatialEncoder",
                          kwargs=dict(image_padding=64, padding_pe=4, pretrained=False)),
        mlp_fine_conf=dict(module="src.models.resnetfc.ResnetFC",
                           kwargs=dict(n_blocks=5, d_hidden=512, combine_layer=3, combine_type="average")))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports everything include/diner_b200.h declares (no compute calls)."""
    import ctypes
    from diner_b200 import capi
    lib = capi.load_library()
    header = open(os.path.join(ROOT, "include", "diner_b200.h")).read()
    declared = set(re.findall(r"\b(diner_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "libdiner_b200.so does not export %s" % name
    assert declared == set(capi.SIGNATURES), "ctypes table and header disagree: %s" % (declared ^ set(capi.SIGNATURES))
    assert lib.diner_version() >= 1
    assert isinstance(lib.diner_last_error(), bytes)


def test_state_dict_keys_match_reference():
    """Checkpoints of the reference must load: same keys and shapes (SURVEY §5 'Checkpoint / resume')."""
    want = [l.split() for l in open(os.path.join(ROOT, "tests", "golden", "reference_state_dict_keys.txt"))]
    sd = _model().state_dict()
    assert [k for k, _ in want] == list(sd.keys())
    for k, shp in want:
        assert ("x".join(map(str, sd[k].shape)) or "scalar") == shp, k


def test_import_paths_and_signatures():
    from src.models.nerf_renderer import NeRFRendererDGS
    from src.models.resnetfc import ResnetFC
    from src.models.image_encoder import SpatialEncoder
    from src.util.import_helper import import_obj
    assert import_obj("src.models.nerf_renderer.NeRFRendererDGS") is NeRFRendererDGS
    sig = inspect.signature(NeRFRendererDGS.__init__)
    assert list(sig.parameters)[1:] == ["n_samples", "n_depth_candidates", "n_gaussian", "eval_batch_size", "white_bkgd"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [40, 1000, 15, 100000, True]
    assert list(inspect.signature(NeRFRendererDGS.forward).parameters)[1:] == ["model", "rays", "want_weights"]
    assert list(inspect.signature(ResnetFC.__init__).parameters)[1:] == [
        "d_in", "d_out", "n_blocks", "d_latent", "d_hidden", "beta", "combine_layer", "combine_type"]
    assert list(inspect.signature(SpatialEncoder.__init__).parameters)[1:] == [
        "backbone", "pretrained", "num_layers", "index_interp", "index_padding", "upsample_interp",
        "use_first_pool", "image_padding", "padding_pe"]
    r = NeRFRendererDGS()
    r.n_samples, r.n_gaussian = 64, int(15 * 64 / 40)     # the CLI mutates these (create_prediction_folder.py:44-47)
    assert len(list(r.parameters())) == 0


def test_no_cpu_fallback():
    """The render path must fail loudly without CUDA instead of silently computing elsewhere."""
    from src.models.nerf_renderer import NeRFRendererDGS
    from diner_b200 import synthetic as S
    m = _model().eval()
    b = S.make_scene(32, 32, 2, 1)
    lat = S.make_latent(1, 2, 512, 80, 80)
    from diner_b200.scene_ops import depth2normal
    n = depth2normal(b["src_depths"].flatten(end_dim=1), b["src_intrinsics"].flatten(end_dim=1)).reshape(1, 2, 3, 32, 32)
    m.encoder.set_scene(lat, b["src_depths"], b["src_depth_stds"], n)
    m.set_cameras(b["src_extrinsics"], b["src_intrinsics"], 32, 32)
    rays = torch.zeros(1, 4, 8)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="CUDA"):
            NeRFRendererDGS()(m, rays)
        with pytest.raises(RuntimeError, match="CUDA"):
            m(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        NeRFRendererDGS()(m, rays)        # grad enabled (training step): same loud failure, no autograd-through-PyTorch fallback
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))     # PixelNeRF.forward on its own has no backward


def test_scene_ops_match_oracle():
    from diner_b200 import synthetic as S
    from diner_b200.scene_ops import depth2normal
    from oracle import diner_oracle as O
    b = S.make_scene(48, 64, 3, 2, seed=3)
    d, K = b["src_depths"].flatten(end_dim=1), b["src_intrinsics"].flatten(end_dim=1)
    assert torch.equal(depth2normal(d, K), O.depth2normal(d, K))


def test_leaf_modules_match_oracle():
    from diner_b200.positional_encoding import PositionalEncoding
    from diner_b200.resnetfc import ResnetFC
    from diner_b200 import synthetic as S
    from oracle import diner_oracle as O
    x = S.hash_normal((5, 7, 3), 1)
    assert torch.equal(PositionalEncoding(6, 3, 6.28)(x), O.positional_encoding(x, 6, 6.28))
    sd = S.make_mlp_state(d_in=55, d_latent=32, d_hidden=64, seed=2)
    net = ResnetFC(55, 4, 5, 32, 64, combine_layer=3)
    net.load_state_dict(sd)
    zx = S.hash_normal((2, 4, 9, 32 + 55), 3)
    sc = O.Scene(*([None] * 8), mlp=sd)
    with torch.no_grad():
        assert torch.allclose(net(zx, combine_dim=1), O.resnetfc(sc, zx), atol=1e-6)


def test_ctypes_signatures_match_header_arity():
    """Every ctypes prototype in diner_b200/capi.py has as many arguments as the declaration in include/diner_b200.h
    (a mismatch would corrupt the call silently: ctypes does not check arity against the binary)."""
    from diner_b200 import capi
    header = open(os.path.join(ROOT, "include", "diner_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    decls = dict(re.findall(r"\b(diner_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    assert set(decls) == set(capi.SIGNATURES)
    for name, params in decls.items():
        params = params.strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert n == len(capi.SIGNATURES[name][1]), "%s: header has %d parameters, ctypes table %d" % (
            name, n, len(capi.SIGNATURES[name][1]))


def test_encoder_emits_channels_last_latent_with_reference_values():
    """SpatialEncoder.forward writes the pyramid levels straight into channels-last storage (scene prepare):
    logical shape and values as the reference (image_encoder.py:281-291) -- checked live against the unmodified reference encoder
    when its tree is present, gradients included -- but NHWC strides, which libdiner_b200 borrows without a re-layout pass."""
    from diner_b200 import synthetic as S
    from src.models.image_encoder import SpatialEncoder
    torch.manual_seed(0)
    enc = SpatialEncoder(pretrained=False, image_padding=64, padding_pe=4).eval()
    b = S.make_scene(64, 64, 2, 1)
    nrm = torch.zeros(1, 2, 3, 64, 64)
    with torch.no_grad():
        lat = enc(b["src_rgbs"], b["src_depths"], b["src_depth_stds"], nrm)
    assert lat.shape == (1, 2, 512, 96, 96) and lat.permute(0, 1, 3, 4, 2).is_contiguous() and not lat.is_contiguous()
    if not os.path.isdir("/root/reference/src"):
        return
    from oracle import ref_import
    ns = ref_import.load()
    ref = ns.image_encoder.SpatialEncoder(pretrained=False, image_padding=64, padding_pe=4).eval()
    ref.load_state_dict(enc.state_dict())
    with torch.no_grad():
        ref(b["src_rgbs"], b["src_depths"], b["src_depth_stds"], nrm)
    assert torch.equal(ref.latent, lat)
    g = S.hash_normal(tuple(lat.shape), 3)
    (enc(b["src_rgbs"], b["src_depths"], b["src_depth_stds"], nrm) * g).sum().backward()
    ref(b["src_rgbs"], b["src_depths"], b["src_depth_stds"], nrm)
    (ref.latent * g).sum().backward()
    assert torch.equal(enc.model.conv1.weight.grad, ref.model.conv1.weight.grad)

"""Test infrastructure ONLY (never imported by the product path).

Imports the *unmodified* reference modules from /root/reference with the three import
stubs SURVEY.md §8(c) lists (dotmap, matplotlib.pyplot, imageio).  Exists only in the
build container: /root/reference is absent on the GPU box, so this module is used solely by
``oracle/make_golden.py`` (fixture generation) and by CPU tests that are skipped when the
reference tree is missing.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("DINER_REFERENCE_ROOT", "/root/reference")


class _DotMap(dict):
    """Minimal stand-in for dotmap.DotMap (attribute access + kwargs ctor)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "models"))


def load():
    """Returns a namespace with the reference classes/functions of the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "dotmap" not in sys.modules:
        m = types.ModuleType("dotmap")
        m.DotMap = _DotMap
        sys.modules["dotmap"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    # The reference package is called ``src`` -- the same name as this repo's drop-in shim
    # package.  Load it under a private alias so both can coexist in one interpreter.
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    # the reference's `src` is a namespace package (no __init__.py); a regular `src` package anywhere on
    # sys.path (this repo's shim) would win over it, so hide those entries while importing
    old_path = list(sys.path)
    sys.path[:] = [REF_ROOT] + [p for p in old_path
                               if not os.path.exists(os.path.join(p or os.getcwd(), "src", "__init__.py"))]
    try:
        import importlib
        ns = types.SimpleNamespace()
        ns.nerf_renderer = importlib.import_module("src.models.nerf_renderer")
        ns.pixelnerf = importlib.import_module("src.models.pixelnerf")
        ns.resnetfc = importlib.import_module("src.models.resnetfc")
        ns.image_encoder = importlib.import_module("src.models.image_encoder")
        ns.positional_encoding = importlib.import_module("src.models.positional_encoding")
        ns.torch_helpers = importlib.import_module("src.util.torch_helpers")
        ns.cam_geometry = importlib.import_module("src.util.cam_geometry")
        ns.depth2normal = importlib.import_module("src.util.depth2normal")
    finally:
        sys.path[:] = old_path
        ref_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    return ns

"""`module: dotted.path` plugin loader -- the reference's plugin boundary (src/util/import_helper.py:16-24)."""
import importlib


def import_obj(path: str):
    """'pkg.mod.Name' -> the object `Name` of module `pkg.mod`."""
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)

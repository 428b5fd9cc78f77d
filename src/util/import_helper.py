"""Drop-in import path of the reference plugin loader (src/util/import_helper.py:16-24)."""
from diner_b200.import_helper import import_obj  # noqa: F401

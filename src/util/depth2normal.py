"""Drop-in import path for depth2normal (reference src/util/depth2normal.py:6-87)."""
from diner_b200.scene_ops import depth2normal  # noqa: F401

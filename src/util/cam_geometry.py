"""Drop-in import path for gen_rays (reference src/util/cam_geometry.py:5-48)."""
from diner_b200.synthetic import gen_rays  # noqa: F401

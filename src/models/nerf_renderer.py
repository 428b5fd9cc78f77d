"""Drop-in import path of the reference (configs/train_dtu.yaml:32-53 name `src.models.nerf_renderer`); implementation in diner_b200/nerf_renderer.py."""
from diner_b200.nerf_renderer import *  # noqa: F401,F403
from diner_b200.nerf_renderer import NeRFRendererDGS  # noqa: F401

"""Drop-in import path of the reference (configs/train_dtu.yaml:32-53 name `src.models.image_encoder`); implementation in diner_b200/image_encoder.py."""
from diner_b200.image_encoder import *  # noqa: F401,F403
from diner_b200.image_encoder import SpatialEncoder  # noqa: F401

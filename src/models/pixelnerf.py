"""Drop-in import path of the reference (configs/train_dtu.yaml:32-53 name `src.models.pixelnerf`); implementation in diner_b200/pixelnerf.py."""
from diner_b200.pixelnerf import *  # noqa: F401,F403
from diner_b200.pixelnerf import PixelNeRF  # noqa: F401

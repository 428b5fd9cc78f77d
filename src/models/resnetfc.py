"""Drop-in import path of the reference (configs/train_dtu.yaml:32-53 name `src.models.resnetfc`); implementation in diner_b200/resnetfc.py."""
from diner_b200.resnetfc import *  # noqa: F401,F403
from diner_b200.resnetfc import ResnetFC, ResnetBlockFC  # noqa: F401

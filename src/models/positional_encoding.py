"""Drop-in import path of the reference (configs/train_dtu.yaml:32-53 name `src.models.positional_encoding`); implementation in diner_b200/positional_encoding.py."""
from diner_b200.positional_encoding import *  # noqa: F401,F403
from diner_b200.positional_encoding import PositionalEncoding  # noqa: F401

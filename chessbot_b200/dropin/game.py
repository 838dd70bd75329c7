"""Bare-name shim: put chessbot_b200/dropin on sys.path ahead of the reference's main/ and its drivers
(train.py, main.ipynb, puzzles.py) import the B200 implementation of `game` unchanged."""
from chessbot_b200.game import *  # noqa: F401,F403
from chessbot_b200 import game as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})

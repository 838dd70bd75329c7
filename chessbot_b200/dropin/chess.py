"""`import chess` shim for hosts without python-chess: the subset the reference's path uses, served by
the C ABI host rules (chessbot_b200/board.py).  Remove this file where python-chess is installed."""
from chessbot_b200.board import *  # noqa: F401,F403
from chessbot_b200.board import Board, Move, Outcome, Termination  # noqa: F401

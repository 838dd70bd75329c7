"""ORACLE (test infrastructure). Restatement of game.py:8-157 over oracle.pychess_shim."""
import numpy as np

from . import pychess_shim as chess
from . import ref_gameimage


class OracleConfig:
    """The constants of config.py:3-42 that parameterise the hot path."""

    def __init__(self, T=8, conv_filters=48, num_residual_blocks=4, max_game_length=512):
        self.T = T
        self.num_actions = 4672
        self.conv_filters = conv_filters
        self.num_residual_blocks = num_residual_blocks
        self.input_dims = (8, 8, 14 * T + 7)
        self.max_game_length = max_game_length
        self.num_mcts_sims = (50, 300)
        self.num_mcts_sims_p = 0.25
        self.num_mcts_sampling_moves = 30


class Game:
    def __init__(self, config, board=None):
        self.board = chess.Board() if board is None else board
        self.config = config
        self.outcome = None
        self.search_statistics = []

    @property
    def history(self):
        return [m.uci() for m in self.board.move_stack]

    @property
    def history_len(self):
        return len(self.board.move_stack)

    @property
    def outcome_str(self):  # game.py:27-36
        if self.outcome is None and self.history_len == self.config.max_game_length:
            return "Draw: outcome=Max moves reached"
        name = self.outcome.termination.name
        if self.outcome.winner is True:
            return f"White wins: outcome={name}"
        if self.outcome.winner is False:
            return f"Black wins: outcome={name}"
        return f"Draw: outcome={name}"

    def terminal(self):  # game.py:49-59 (falls through to None when not over)
        if self.history_len == self.config.max_game_length:
            return True
        o = self.board.outcome(claim_draw=True)
        if o is not None:
            self.outcome = o
            return True
        return None

    def terminal_value(self, player):  # game.py:61-77
        if not self.outcome or self.outcome.winner is None:
            return 0
        return 1 if self.outcome.winner == player else -1

    def legal_moves(self):
        return [m.uci() for m in self.board.legal_moves]

    def make_move(self, move: str):
        self.board.push_uci(move)

    def make_image(self, state_index: int = -1):  # game.py:117-129
        b = self.board
        if state_index != -1:
            b = b.copy()
            while len(b.move_stack) > state_index:
                b.pop()
        return ref_gameimage.board_to_image(b, self.config.T).astype(np.int16)

    def clone(self):
        return Game(self.config, board=self.board.copy())

    def to_play(self):
        return self.board.turn

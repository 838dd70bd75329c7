"""The reference's `Game` surface (game.py:8-157) over this package's rules engine and CUDA encoder.

Same attributes and methods, same argument meaning and return conventions (callers: mcts.py, train.py, puzzles.py, the reference's
tests/game.py), written against the C-ABI board (`chessbot_b200.board.Board`, include/chessbot_b200.h `cb_board_*`): one
`cb_board_outcome` call decides "is it over" and "who won", move lists are decoded from the packed move codes the library returns,
images come from the device encoder.  A board that is not ours (a python-chess object) is served through the same python-chess calls
the reference makes.
"""
import numpy as np

from . import actionspace as asp
from . import board as chess
from . import gameimage
from .config import Config

_RESULT_NAMES = {True: "White wins", False: "Black wins", None: "Draw"}


class Game:
    def __init__(self, config: Config, board=None):
        self.config = config
        self.board = board if board is not None else chess.Board()
        self.outcome = None               # chess.Outcome once terminal() has seen the game end by the rules
        self.search_statistics = []       # train.py:43-47 appends one visit distribution per recorded move

    # ---- move history (game.py:10-26)
    @property
    def history(self) -> list:
        return [m.uci() for m in self.board.move_stack]

    @property
    def history_len(self) -> int:
        return len(self.board.move_stack)

    # ---- end of game (game.py:28-77)
    def _length_cap_reached(self) -> bool:
        return self.history_len == self.config.max_game_length

    def terminal(self) -> bool:
        """True at config.max_game_length plies or when the rules (draw claims included) end the game; remembers the outcome."""
        if self._length_cap_reached():
            return True
        result = self.board.outcome(claim_draw=True)
        if result is None:
            return False
        self.outcome = result
        return True

    def terminal_value(self, player: bool) -> int:
        """+1 / -1 from `player`'s side once terminal() recorded a decisive outcome, else 0."""
        winner = self.outcome.winner if self.outcome else None
        if winner is None:
            return 0
        return 1 if winner == player else -1

    @property
    def outcome_str(self) -> str:
        if self.outcome is None and self._length_cap_reached():
            return "Draw: outcome=Max moves reached"
        return f"{_RESULT_NAMES[self.outcome.winner]}: outcome={self.outcome.termination.name}"

    # ---- moves (game.py:79-101)
    def legal_moves(self) -> list:
        return [m.uci() for m in self.board.legal_moves]

    def legal_actions(self) -> list:
        side = self.to_play()
        return [asp.uci_to_action(uci, side) for uci in self.legal_moves()]

    def make_move(self, move: str) -> None:
        self.board.push_uci(move)

    def to_play(self) -> bool:
        return self.board.turn

    def clone(self) -> "Game":
        return Game(self.config, self.board.copy())

    # ---- training views (game.py:103-157)
    def make_target(self, state_index: int) -> tuple:
        ply = self.history_len - 1 if state_index == -1 else state_index
        mover = chess.WHITE if ply % 2 == 0 else chess.BLACK
        return np.array(self.search_statistics[state_index]), self.terminal_value(mover)

    def make_image(self, state_index: int = -1) -> np.ndarray:
        """int16 [8, 8, 14 T + 7] planes of the current position, or of the position after `state_index` plies."""
        position = self.board
        if state_index != -1:
            position = self.board.copy()
            for _ in range(len(position.move_stack) - state_index):
                position.pop()
        return gameimage.board_to_image(position, self.config.T).astype(np.int16)

    def make_image_sample(config: Config):
        """game.py:152-157 (called on the class): the image after 1. e4 e5 2. Nf3 Nc6."""
        position = chess.Board()
        for san in ("e4", "e5", "Nf3", "Nc6"):
            position.push_san(san)
        return gameimage.board_to_image(position, config.T).astype(np.int16)

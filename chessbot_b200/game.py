"""Drop-in for the reference's game.py (game.py:8-157): Game over a python-chess-compatible board.
Rules come from the C ABI host engine (chessbot_b200.board), images from the CUDA encoder."""
import numpy as np

from . import actionspace as asp
from . import board as chess
from . import gameimage
from .config import Config


class Game:
    def __init__(self, config: Config, board=None):
        self.board = chess.Board() if board is None else board
        self.config = config
        self.outcome = None
        self.search_statistics = []

    @property
    def history(self) -> list:
        return [move.uci() for move in self.board.move_stack]

    @property
    def history_len(self) -> int:
        return len(self.board.move_stack)

    @property
    def outcome_str(self) -> str:
        if self.outcome is None and self.history_len == self.config.max_game_length:
            return "Draw: outcome=Max moves reached"
        if self.outcome.winner == chess.WHITE and self.outcome.winner is not None:
            return f"White wins: outcome={self.outcome.termination.name}"
        elif self.outcome.winner == chess.BLACK and self.outcome.winner is not None:
            return f"Black wins: outcome={self.outcome.termination.name}"
        return f"Draw: outcome={self.outcome.termination.name}"

    def terminal(self) -> bool:
        if self.history_len == self.config.max_game_length:
            return True
        if self.board.is_game_over(claim_draw=True):
            self.outcome = self.board.outcome(claim_draw=True)
            return True

    def terminal_value(self, player: bool) -> int:
        if not self.outcome:
            return 0
        if self.outcome.winner is None:
            return 0
        return 1 if self.outcome.winner == player else -1

    def legal_moves(self) -> list:
        return [move.uci() for move in self.board.legal_moves]

    def legal_actions(self) -> list:
        return [asp.uci_to_action(move, self.to_play()) for move in self.legal_moves()]

    def make_move(self, move: str) -> None:
        self.board.push_uci(move)

    def make_target(self, state_index: int) -> tuple:
        index = self.history_len - 1 if state_index == -1 else state_index
        to_play = chess.WHITE if index % 2 == 0 else chess.BLACK
        return np.array(self.search_statistics[state_index]), self.terminal_value(to_play)

    def make_image(self, state_index: int = -1) -> np.ndarray:
        if state_index == -1:
            return gameimage.board_to_image(self.board, self.config.T).astype(np.int16)
        tmp_board = self.board.copy()
        while len(tmp_board.move_stack) > state_index:
            tmp_board.pop()
        return gameimage.board_to_image(tmp_board, self.config.T).astype(np.int16)

    def clone(self) -> "Game":
        return Game(config=self.config, board=self.board.copy())

    def to_play(self) -> bool:
        return self.board.turn

    def make_image_sample(config: Config):
        board = chess.Board()
        for move in ["e4", "e5", "Nf3", "Nc6"]:
            board.push_san(move)
        return gameimage.board_to_image(board, config.T).astype(np.int16)

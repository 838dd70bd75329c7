"""Host logic of the batched puzzle driver (selfplay.solve_puzzles; puzzles.py:83-86 / eval_trt.py:41-52), CPU only.
The stand-in below keeps the reference Puzzle's interface contract (board, status, move_uci plays the scripted reply)."""
from chessbot_b200 import selfplay
from chessbot_b200.board import Board


class _Puzzle:
    def __init__(self, fen, moves):
        self.moves, self.played, self.status = moves.split(), [], "IN_PROGRESS"
        self.board = Board(fen)
        self._scripted()

    def _scripted(self):
        m = self.moves[len(self.played)]
        self.board.push_uci(m)
        self.played.append(m)

    def move_uci(self, move):
        self.board.push_uci(move)
        self.played.append(move)
        if self.played != self.moves[:len(self.played)]:
            self.status = "FAILED"
        elif len(self.played) == len(self.moves):
            self.status = "COMPLETED"
        else:
            self._scripted()


class _ScriptBot:
    """answers from a table keyed by position: stands in for the device search"""
    def __init__(self, answers):
        self.answers, self.batches = answers, []

    def get_moves(self, boards):
        self.batches.append(len(boards))
        return [self.answers[b.fen()] for b in boards]


def _line(fen, moves):
    b, out = Board(fen), {}
    for i, m in enumerate(moves.split()):
        if i % 2 == 1:
            out[b.fen()] = m
        b.push_uci(m)
    return out


def test_all_live_puzzles_advance_together_and_finish_like_sequential_solve():
    start = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
    lines = [(start, "e2e4 e7e5 g1f3 b8c6"), (start, "d2d4 d7d5"), (start, "c2c4 e7e5 b1c3 g8f6 g2g3 d7d5")]
    answers = {}
    for fen, mv in lines:
        answers.update(_line(fen, mv))
    puzzles = [_Puzzle(f, m) for f, m in lines]
    bot = _ScriptBot(answers)
    rounds = selfplay.solve_puzzles(puzzles, bot)
    assert [p.status for p in puzzles] == ["COMPLETED"] * 3
    assert [p.played for p in puzzles] == [m.split() for _, m in lines]
    assert rounds == 3 and bot.batches == [3, 2, 1]          # one device batch per round, finished puzzles drop out


def test_wrong_answer_fails_only_that_puzzle():
    start = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
    lines = [(start, "e2e4 e7e5 g1f3 b8c6"), (start, "d2d4 d7d5 c2c4 e7e6")]
    answers = {}
    for fen, mv in lines:
        answers.update(_line(fen, mv))
    b = Board(start)
    b.push_uci("e2e4")
    answers[b.fen()] = "c7c5"                                  # not the scripted answer
    puzzles = [_Puzzle(f, m) for f, m in lines]
    selfplay.solve_puzzles(puzzles, _ScriptBot(answers))
    assert [p.status for p in puzzles] == ["FAILED", "COMPLETED"]

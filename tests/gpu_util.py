"""Helpers shared by the -m gpu parity tests."""
import numpy as np

from oracle import pychess_shim as ochess


def oracle_board(fen, moves):
    b = ochess.Board(str(fen))
    for m in str(moves).split():
        b.push_uci(m)
    return b


def device_board(fen, moves):
    from chessbot_b200.board import Board
    b = Board(str(fen))
    for m in str(moves).split():
        b.push_uci(m)
    return b


def random_games(n, seed, min_plies=8, max_plies=160, back_bias=0.15):
    """SURVEY.md §8(d) synthetic positions: seeded random playouts, full move stacks kept.
    Returns list of (fen, 'uci uci ...') with non-terminal final positions."""
    import random
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        b = ochess.Board()
        moves = []
        target = rnd.randint(min_plies, max_plies)
        ok = True
        for _ in range(target):
            ms = list(b.legal_moves)
            if not ms or b.is_game_over(claim_draw=True):
                ok = False
                break
            back = [m for m in ms if len(b.move_stack) >= 2 and m.from_square == b.move_stack[-2].to_square
                    and m.to_square == b.move_stack[-2].from_square]
            m = back[0] if back and rnd.random() < back_bias else rnd.choice(ms)
            b.push(m)
            moves.append(m.uci())
        if ok and not b.is_game_over(claim_draw=True):
            out.append((ochess.STARTING_FEN, " ".join(moves)))
    return out


def bf16_round(x):
    import torch
    return torch.as_tensor(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()

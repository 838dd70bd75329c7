"""The reference's own unit tests (main/tests/{game,gameimage,actionspace}.py, 55 tests) as recorded known answers.

tests/golden/reference_tests.npz is the log of every call those tests make into the reference's surface, with the results the
reference itself produced while all 55 tests passed (scripts/run_reference_tests.py --over shim --record, run in the build container
over the verbatim reference files).  Here the calls are replayed through this package's drop-in modules:
* host calls (Game.terminal / terminal_value / legal_moves / history / history_len / to_play, actionspace.uci_to_action /
  uci_to_actionstr) on the CPU: C-ABI rules engine and the closed-form action index;
* Game.make_image on the GPU (-m gpu): the CUDA encoder, bit for bit (all 31 tests of main/tests/gameimage.py go through it).
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _load():
    z = np.load(os.path.join(GOLDEN, "reference_tests.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


def _game(entry):
    from chessbot_b200 import board as chess
    from chessbot_b200.config import Config
    from chessbot_b200.game import Game
    cfg = Config()
    cfg.T = entry["T"]
    cfg.max_game_length = entry["max_len"]
    b = chess.Board(entry["fen"])
    for u in entry["moves"].split():
        b.push_uci(u)
    return Game(cfg, board=b)


def test_golden_holds_the_whole_suite():
    meta, z = _load()
    fns = {e["fn"] for e in meta}
    assert {"Game.make_image", "Game.terminal", "Game.terminal_value", "Game.legal_moves", "Game.history", "Game.history_len", "Game.to_play",
            "actionspace.uci_to_actionstr"} <= fns
    assert sum(e["fn"] == "Game.make_image" for e in meta) == len([k for k in z.files if k.startswith("image_")])
    assert len(meta) > 400


def test_host_calls_of_the_reference_suite_replay_exactly():
    import chessbot_b200.actionspace as asp
    meta, _ = _load()
    n = 0
    for e in meta:
        if e["fn"].startswith("actionspace."):
            got = getattr(asp, e["fn"].split(".")[1])(*e["args"])
            assert got == e["out"], e
            n += 1
        elif e["fn"] != "Game.make_image":
            g = _game(e)
            name = e["fn"].split(".")[1]
            attr = getattr(g, name)
            if name == "terminal_value":
                g.terminal()                               # the reference's tests call terminal() first: it records the outcome
            got = attr(*e["args"]) if callable(attr) else attr
            if name == "terminal":
                assert bool(got) == bool(e["out"]), e      # the reference returns None where the game goes on
            else:
                assert got == e["out"], (e, got)
            n += 1
    assert n > 400


@pytest.mark.gpu
def test_gpu_make_image_calls_of_the_reference_suite_replay_bit_exactly():
    meta, z = _load()
    n = 0
    for e in meta:
        if e["fn"] != "Game.make_image":
            continue
        want = z[f"image_{e['out']['image']}"]
        got = _game(e).make_image(*e["args"])
        assert got.dtype == np.int16 and list(got.shape) == e["out"]["shape"]
        assert np.array_equal(got, want), e
        n += 1
    assert n >= 40

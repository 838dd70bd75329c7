#!/usr/bin/env python
"""Runs the reference's OWN unit tests (main/tests/{game,gameimage,actionspace}.py, 55 tests) verbatim, from the read-only reference tree.
Container-only: /root/reference does not exist on the GPU box and none of its sources are copied into this repository.

    python scripts/run_reference_tests.py --over shim      # verbatim reference modules over oracle/pychess_shim (CPU)
    python scripts/run_reference_tests.py --over dropin    # the same test files over chessbot_b200/dropin (gameimage needs the GPU)
    python scripts/run_reference_tests.py --over shim --record tests/golden/reference_tests.npz

--record (with --over shim): while the tests run, every call they make into the reference's surface
(Game.make_image / terminal / terminal_value / legal_moves / history / history_len / to_play, actionspace.uci_to_action /
uci_to_actionstr) is logged with its inputs (start FEN, move list, T, arguments) and its result.  Only if ALL tests pass the log is
written as the golden file: those results are then known answers the reference's own test-suite accepted.  The -m gpu test
tests/test_gpu_reference_suite.py replays the calls through chessbot_b200/dropin (CUDA encoder, C-ABI rules) and compares bit for bit.
"""
import argparse
import importlib.util
import json
import os
import sys
import unittest

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_TESTS = "/root/reference/main/tests"
FILES = ("game", "actionspace", "gameimage")


def _load_test_module(name):
    spec = importlib.util.spec_from_file_location(f"reference_tests_{name}", os.path.join(REF_TESTS, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _position_key(board):
    """(start FEN, uci moves) of a python-chess-compatible board"""
    tmp = board.copy()
    moves = []
    while len(tmp.move_stack) > 0:
        moves.append(tmp.pop().uci())
    return tmp.fen(), " ".join(reversed(moves))


def install_recorder(ns, log):
    """wraps the reference's Game methods and actionspace functions (modules loaded by oracle.verbatim_reference)"""
    Game, asp = ns.game.Game, ns.actionspace

    def wrap_method(name, is_property=False):
        orig = getattr(Game, name)
        getter = orig.fget if is_property else orig

        def rec(self, *args):
            fen, moves = _position_key(self.board)
            out = getter(self, *args)
            log.append({"fn": "Game." + name, "fen": fen, "moves": moves, "T": int(self.config.T), "max_len": int(self.config.max_game_length),
                        "args": [a if not isinstance(a, (bool, np.bool_)) else bool(a) for a in args], "out": out})
            return out
        setattr(Game, name, property(rec) if is_property else rec)

    for m in ("make_image", "terminal", "terminal_value", "legal_moves", "to_play"):
        wrap_method(m)
    for m in ("history", "history_len"):
        wrap_method(m, is_property=True)
    for fn in ("uci_to_action", "uci_to_actionstr"):
        orig = getattr(asp, fn)

        def rec(uci, player, _orig=orig, _fn=fn):
            out = _orig(uci, player)
            log.append({"fn": "actionspace." + _fn, "args": [uci, bool(player)], "out": out})
            return out
        setattr(asp, fn, rec)


def save_golden(path, log):
    images = [e["out"] for e in log if e["fn"] == "Game.make_image"]
    meta = []
    for e in log:
        e = dict(e)
        if e["fn"] == "Game.make_image":
            e["out"] = {"image": len([1 for m in meta if m["fn"] == "Game.make_image"]), "shape": list(e["out"].shape)}
        elif isinstance(e["out"], (np.integer,)):
            e["out"] = int(e["out"])
        elif isinstance(e["out"], (np.bool_,)):
            e["out"] = bool(e["out"])
        meta.append(e)
    arrays = {f"image_{i}": np.asarray(a, dtype=np.int16) for i, a in enumerate(images)}
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    return len(meta), len(images)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--over", choices=["shim", "dropin"], default="shim")
    ap.add_argument("--record", default=None)
    ap.add_argument("--files", default=",".join(FILES))
    a = ap.parse_args()
    if not os.path.isdir(REF_TESTS):
        print("reference tree not present: this script only runs in the build container")
        return 2
    log = []
    if a.over == "shim":
        from oracle import verbatim_reference
        ns = verbatim_reference.load()                     # registers chess / config / model stand-ins + the verbatim modules in sys.modules
        for k in ("actionspace", "gameimage", "game", "mcts"):
            sys.modules[k] = getattr(ns, k)
        if a.record:
            install_recorder(ns, log)
    else:
        if a.record:
            raise SystemExit("--record needs --over shim (known answers come from the reference, not from this repository)")
        sys.path.insert(0, os.path.join(ROOT, "chessbot_b200", "dropin"))
    suite = unittest.TestSuite()
    for name in a.files.split(","):
        suite.addTests(unittest.defaultTestLoader.loadTestsFromModule(_load_test_module(name)))
    res = unittest.TextTestRunner(verbosity=1).run(suite)
    print(f"reference tests over {a.over}: ran {res.testsRun}, failures {len(res.failures)}, errors {len(res.errors)}")
    if a.record:
        if not res.wasSuccessful():
            raise SystemExit("not recording: the reference's tests did not all pass")
        n, k = save_golden(a.record, log)
        print(f"recorded {n} calls ({k} images) -> {a.record}")
    return 0 if res.wasSuccessful() else 1


if __name__ == "__main__":
    sys.exit(main())

"""ORACLE — test infrastructure only, never product code.

CPU restatement of the reference hot path (tomaz-piko/ChessBot `main/`):
  pychess_shim.py   python-chess subset (third-party, un-vendored, un-pinned)
  ref_actionspace.py actionspace.py:116-180 in closed form
  ref_gameimage.py  gameimage.py:149-192 in closed form over bitboards
  ref_game.py       game.py:8-157
  ref_mcts.py       mcts.py:22-161 as a struct-of-arrays tree, NumPy-2 (NEP 50) float semantics
  ref_model.py      model.py:19-56 forward in torch fp32
  verbatim_reference.py  container-only loader of the reference's own files (validation + golden vectors)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
package.  The product (chessbot_b200/) never does; it fails loudly without its CUDA library.
"""

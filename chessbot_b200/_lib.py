"""ctypes binding of libchessbot_b200.so (include/chessbot_b200.h).  There is no CPU fallback: if the
library is missing the import fails and says how to build it."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libchessbot_b200.so")
if not os.path.exists(_PATH):
    raise ImportError(f"{_PATH} not found: build it with `python -m chessbot_b200.build` "
                      "(chessbot_b200 has no CPU fallback)")
lib = C.CDLL(_PATH)

vp, i32, u64, u32, u16p = C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint16)
_SIGS = {
    "cb_last_error": (C.c_char_p, []),
    "cb_version": (C.c_char_p, []),
    "cb_board_new": (vp, [C.c_char_p]),
    "cb_board_free": (None, [vp]),
    "cb_board_copy": (vp, [vp]),
    "cb_board_push_uci": (i32, [vp, C.c_char_p]),
    "cb_board_pop": (i32, [vp]),
    "cb_board_ply": (i32, [vp]),
    "cb_board_turn": (i32, [vp]),
    "cb_board_halfmove_clock": (i32, [vp]),
    "cb_board_castling": (i32, [vp]),
    "cb_board_ep_square": (i32, [vp]),
    "cb_board_pieces_mask": (u64, [vp, i32, i32]),
    "cb_board_legal_moves": (i32, [vp, u16p]),
    "cb_board_move_stack": (i32, [vp, u16p, i32]),
    "cb_board_outcome": (i32, [vp, i32]),
    "cb_board_is_repetition": (i32, [vp, i32]),
    "cb_board_is_check": (i32, [vp]),
    "cb_board_fen": (i32, [vp, C.c_char_p, i32]),
    "cb_board_mirror": (i32, [vp]),
    "cb_board_perft": (u64, [vp, i32]),
    "cb_board_export_record": (i32, [vp, vp, i32]),
    "cb_boards_export_records": (i32, [vp, i32, vp, i32, vp, vp]),
    "cb_uci_to_action": (i32, [C.c_char_p, i32]),
    "cb_move_to_action": (i32, [C.c_uint16, i32]),
    "cb_move_to_uci": (i32, [C.c_uint16, C.c_char_p]),
    "cb_set_device_allocator": (i32, [vp, vp, vp]),
    "cb_search_workspace_bytes": (C.c_size_t, [i32, i32, i32, i32, i32]),
    "cb_ctx_create": (vp, [i32]),
    "cb_ctx_destroy": (None, [vp]),
    "cb_ctx_sm_count": (i32, [vp]),
    "cb_ctx_device": (i32, [vp]),
    "cb_set_exp_table": (i32, [vp, vp]),
    "cb_planes_act_bytes": (C.c_size_t, [i32]),
    "cb_encode": (i32, [vp, vp, vp, i32, vp, vp, vp]),
    "cb_net_create": (i32, [vp, i32, i32, i32, C.c_float, C.c_float]),
    "cb_net_set_tensor": (i32, [vp, C.c_char_p, vp, C.c_size_t]),
    "cb_net_finalize": (i32, [vp]),
    "cb_net_set_stream_precision": (i32, [vp, i32]),
    "cb_net_forward": (i32, [vp, vp, i32, vp, vp, vp, vp]),
    "cb_synth_forward": (i32, [vp, vp, i32, u32, vp, vp, vp]),
    "cb_policy_softmax": (i32, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]),
    "cb_leaf_eval": (i32, [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]),
    "cb_leaf_eval_packed_bytes": (C.c_size_t, [i32, i32]),
    "cb_leaf_eval_host_sync": (i32, [vp, vp, vp, i32, vp, C.c_size_t, vp]),
    "cb_search_create": (i32, [vp, i32, i32, i32, i32]),
    "cb_search_set_leaves": (i32, [vp, i32]),
    "cb_search_begin": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, u32, vp]),
    "cb_search_begin_csr": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, u32, vp]),
    "cb_search_roots_compact_bytes": (C.c_size_t, [i32, i32]),
    "cb_search_read_roots_compact_sync": (i32, [vp, vp, vp, vp]),
    "cb_search_read_roots_compact_async": (i32, [vp, vp, vp]),
    "cb_search_read_roots_compact_wait": (i32, [vp, vp, vp]),
    "cb_search_step": (i32, [vp, vp]),
    "cb_search_run": (i32, [vp, i32, vp]),
    "cb_search_pending_planes_sync": (i32, [vp, vp, vp, vp]),
    "cb_search_step_external": (i32, [vp, vp, vp, vp]),
    "cb_search_status_sync": (i32, [vp, vp, vp]),
    "cb_search_read_roots_sync": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "cb_profile": (i32, [vp, i32]),
    "cb_profile_read_sync": (i32, [vp, vp, vp]),
    "cb_debug_conv_layer": (i32, [vp, i32, vp, vp, vp, i32, i32, vp]),
    "cb_train_param_count": (C.c_size_t, [i32, i32, C.POINTER(C.c_size_t)]),
    "cb_train_create": (vp, [vp, i32, i32, i32, C.c_float, C.c_float, C.c_float, vp, vp, vp]),
    "cb_train_destroy": (None, [vp]),
    "cb_train_set_tensor": (i32, [vp, C.c_char_p, i32, vp, C.c_size_t, vp]),
    "cb_train_get_tensor": (i32, [vp, C.c_char_p, i32, vp, C.c_size_t, vp]),
    "cb_train_forward_backward": (i32, [vp, vp, vp, vp, i32, vp]),
    "cb_train_apply": (i32, [vp, C.c_float, C.c_float, i32, C.c_float, C.c_float, vp]),
    "cb_train_losses_sync": (i32, [vp, vp, vp]),
    "cb_train_outputs_sync": (i32, [vp, vp, vp]),
    "cb_train_profile": (i32, [vp, i32]),
    "cb_train_profile_read_sync": (i32, [vp, vp, vp]),
    "cb_debug_wgrad": (i32, [vp, vp, vp, i32, i32, i32, vp, i32, vp]),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

EXPORTED = sorted(_SIGS)
CB_MAX_MOVES = 224
CB_POS_BYTES = 96
CB_NUM_ACTIONS = 4672


def last_error() -> str:
    return lib.cb_last_error().decode()


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        raise RuntimeError(f"chessbot_b200 {what}: {last_error()}")
    return rc

// Host side of the C ABI: the python-chess surface of game.py / gameimage.py / puzzles.py over
// csrc/rules.cuh, and the export of game records to the device.  See include/chessbot_b200.h.
#include <string.h>

#include <string>
#include <vector>

#include "../../include/chessbot_b200.h"
#include "rules.cuh"

using namespace cb;

static thread_local std::string g_err;
void cb_set_error(const std::string& s) { g_err = s; }

struct cb_board {
    std::vector<Pos> rec;  // rec[0] = root (FEN) position, rec.back() = current
};

namespace {
struct VecPrev {  // predecessor inside a contiguous record
    const Pos* begin;
    const Pos* operator()(const Pos* q) const { return q > begin ? q - 1 : nullptr; }
};

bool parse_fen(const char* fen, Pos& p, std::string& err) {
    memset(&p, 0, sizeof(p));
    std::string s(fen);
    std::vector<std::string> parts;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && s[i] == ' ') ++i;
        size_t j = i;
        while (j < s.size() && s[j] != ' ') ++j;
        if (j > i) parts.push_back(s.substr(i, j - i));
        i = j;
    }
    if (parts.empty()) { err = "empty fen"; return false; }
    int r = 7, f = 0;
    for (char c : parts[0]) {
        if (c == '/') {
            if (f != 8) { err = "bad row in fen"; return false; }
            --r; f = 0;
        } else if (c >= '1' && c <= '8') {
            f += c - '0';
        } else {
            const char* sym = "pnbrqk";
            char lc = (char)(c | 0x20);
            const char* q = strchr(sym, lc);
            if (!q || f > 7 || r < 0) { err = "bad piece in fen"; return false; }
            set_piece(p, r * 8 + f, (int)(q - sym), c == lc ? BLACK : WHITE);
            ++f;
        }
    }
    if (r != 0 || f != 8) { err = "expected 8 rows in position part of fen"; return false; }
    p.turn = WHITE;
    if (parts.size() > 1) {
        if (parts[1] == "w") p.turn = WHITE;
        else if (parts[1] == "b") p.turn = BLACK;
        else { err = "bad turn in fen"; return false; }
    }
    uint8_t cr = 0;
    if (parts.size() > 2 && parts[2] != "-")
        for (char c : parts[2]) {
            if (c == 'K') cr |= CR_WK;
            else if (c == 'Q') cr |= CR_WQ;
            else if (c == 'k') cr |= CR_BK;
            else if (c == 'q') cr |= CR_BQ;
            else { err = "bad castling in fen"; return false; }
        }
    p.castling = clean_castling(p, cr);
    p.ep = -1;
    if (parts.size() > 3 && parts[3] != "-") {
        if (parts[3].size() != 2 || parts[3][0] < 'a' || parts[3][0] > 'h' || parts[3][1] < '1' || parts[3][1] > '8') {
            err = "bad ep square in fen"; return false;
        }
        p.ep = (int8_t)((parts[3][1] - '1') * 8 + (parts[3][0] - 'a'));
    }
    p.halfmove = parts.size() > 4 ? (uint16_t)atoi(parts[4].c_str()) : 0;
    p.ply = 0;
    p.irrev_in = 1;  // start of the record: nothing before it
    p.occur = 1;
    p.last_move = 0;
    p.ep_key = has_legal_en_passant(p) ? p.ep : (int8_t)-1;
    p.hash = key_hash(p);
    return true;
}

int parse_uci(const char* uci, Move& m) {
    size_t n = strlen(uci);
    if (n != 4 && n != 5) return -1;
    for (int k = 0; k < 4; k += 2)
        if (uci[k] < 'a' || uci[k] > 'h' || uci[k + 1] < '1' || uci[k + 1] > '8') return -1;
    int f = (uci[1] - '1') * 8 + (uci[0] - 'a'), t = (uci[3] - '1') * 8 + (uci[2] - 'a');
    int promo = 0;
    if (n == 5) {
        switch (uci[4]) {
            case 'n': promo = 2; break;
            case 'b': promo = 3; break;
            case 'r': promo = 4; break;
            case 'q': promo = 5; break;
            default: return -1;
        }
    }
    m = make_move(f, t, promo);
    return 0;
}
}  // namespace

extern "C" {

const char* cb_last_error(void) { return g_err.c_str(); }
const char* cb_version(void) { return "chessbot_b200 0.1 (sm_100a)"; }

cb_board* cb_board_new(const char* fen) {
    Pos p;
    std::string err;
    if (!parse_fen(fen ? fen : "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", p, err)) {
        cb_set_error(err);
        return nullptr;
    }
    cb_board* b = new cb_board;
    b->rec.push_back(p);
    return b;
}
void cb_board_free(cb_board* b) { delete b; }
cb_board* cb_board_copy(const cb_board* b) { return new cb_board(*b); }

int cb_board_push_uci(cb_board* b, const char* uci) {
    Move m;
    if (parse_uci(uci, m) != 0) { cb_set_error(std::string("invalid uci: ") + uci); return -1; }
    MoveList ml;
    generate_legal(b->rec.back(), ml);
    bool ok = false;
    for (int i = 0; i < ml.n; ++i) ok |= ml.m[i] == m;
    if (!ok) { cb_set_error(std::string("illegal uci: ") + uci); return -1; }
    Pos nxt = push_move(b->rec.back(), m);
    nxt.occur = (uint8_t)count_occurrences(nxt, &b->rec.back(), VecPrev{b->rec.data()});
    b->rec.push_back(nxt);
    return 0;
}
int cb_board_pop(cb_board* b) {
    if (b->rec.size() <= 1) { cb_set_error("pop from empty move stack"); return -1; }
    int m = b->rec.back().last_move;
    b->rec.pop_back();
    return m;
}
int cb_board_ply(const cb_board* b) { return (int)b->rec.size() - 1; }
int cb_board_turn(const cb_board* b) { return b->rec.back().turn; }
int cb_board_halfmove_clock(const cb_board* b) { return b->rec.back().halfmove; }
int cb_board_castling(const cb_board* b) { return b->rec.back().castling; }
int cb_board_ep_square(const cb_board* b) { return b->rec.back().ep; }
uint64_t cb_board_pieces_mask(const cb_board* b, int piece_type, int color) {
    if (piece_type < 1 || piece_type > 6) return 0;
    return b->rec.back().bb[piece_type - 1] & b->rec.back().occ[color ? WHITE : BLACK];
}
int cb_board_legal_moves(const cb_board* b, uint16_t* moves) {
    MoveList ml;
    generate_legal(b->rec.back(), ml);
    memcpy(moves, ml.m, sizeof(Move) * ml.n);
    return ml.n;
}
int cb_board_move_stack(const cb_board* b, uint16_t* moves, int cap) {
    int n = (int)b->rec.size() - 1;
    for (int i = 0; i < n && i < cap; ++i) moves[i] = b->rec[i + 1].last_move;
    return n;
}
int cb_board_outcome(const cb_board* b, int claim_draw) {
    const Pos& cur = b->rec.back();
    MoveList ml;
    generate_legal(cur, ml);
    const Pos* before = b->rec.size() > 1 ? &b->rec[b->rec.size() - 2] : nullptr;
    int o = outcome_claim_draw(cur, ml, before, VecPrev{b->rec.data()}, 0);
    if (!claim_draw && (o == OUT_FIFTY_MOVES || o == OUT_THREEFOLD_REPETITION)) return OUT_NONE;
    return o;
}
int cb_board_is_repetition(const cb_board* b, int count) { return b->rec.back().occur >= count; }
int cb_board_is_check(const cb_board* b) { return checkers_mask(b->rec.back()) != 0; }
int cb_board_mirror(cb_board* b) {
    Pos p = b->rec.back();
    auto flip = [](u64 x) { return __builtin_bswap64(x); };
    for (int i = 0; i < 6; ++i) p.bb[i] = flip(p.bb[i]);
    u64 w = flip(p.occ[WHITE]), k = flip(p.occ[BLACK]);
    p.occ[WHITE] = k;
    p.occ[BLACK] = w;
    p.castling = (uint8_t)(((p.castling & 3) << 2) | ((p.castling >> 2) & 3));
    if (p.ep >= 0) p.ep ^= 56;
    p.turn ^= 1;
    p.ply = 0;
    p.irrev_in = 1;
    p.occur = 1;
    p.last_move = 0;
    p.ep_key = has_legal_en_passant(p) ? p.ep : (int8_t)-1;
    p.hash = key_hash(p);
    b->rec.clear();
    b->rec.push_back(p);
    return 0;
}
int cb_board_fen(const cb_board* b, char* out, int cap) {
    const Pos& p = b->rec.back();
    std::string s;
    for (int r = 7; r >= 0; --r) {
        int empty = 0;
        for (int f = 0; f < 8; ++f) {
            int t = piece_type_at(p, r * 8 + f);
            if (t < 0) { ++empty; continue; }
            if (empty) { s += (char)('0' + empty); empty = 0; }
            char c = "pnbrqk"[t];
            s += (p.occ[WHITE] & bit(r * 8 + f)) ? (char)(c - 32) : c;
        }
        if (empty) s += (char)('0' + empty);
        if (r) s += '/';
    }
    s += p.turn == WHITE ? " w " : " b ";
    std::string cr;
    if (p.castling & CR_WK) cr += 'K';
    if (p.castling & CR_WQ) cr += 'Q';
    if (p.castling & CR_BK) cr += 'k';
    if (p.castling & CR_BQ) cr += 'q';
    s += cr.empty() ? "-" : cr;
    s += ' ';
    if (p.ep_key >= 0) { s += (char)('a' + (p.ep_key & 7)); s += (char)('1' + (p.ep_key >> 3)); }
    else s += '-';
    // python-chess fullmove number: root FENs are rare on this path; derive from ply assuming move 1 start
    s += ' ' + std::to_string(p.halfmove) + ' ' + std::to_string(1 + (p.ply + (b->rec[0].turn == BLACK ? 1 : 0)) / 2);
    if ((int)s.size() + 1 > cap) return -(int)s.size() - 1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
static uint64_t perft_rec(const Pos& p, int depth) {
    MoveList ml;
    generate_legal(p, ml);
    if (depth == 1) return ml.n;
    uint64_t n = 0;
    for (int i = 0; i < ml.n; ++i) n += perft_rec(push_move(p, ml.m[i]), depth - 1);
    return n;
}
uint64_t cb_board_perft(const cb_board* b, int depth) { return depth <= 0 ? 1 : perft_rec(b->rec.back(), depth); }

int cb_board_export_record(const cb_board* b, void* pos_out, int cap) {
    // Needed: 8 frames for the encoder (gameimage.py:169-181), each frame's own repetition flags are
    // already in Pos::occur; for draw claims at and below this position, everything back to the last
    // irreversible move.  The oldest exported position is marked as the start of the record.
    int cur = (int)b->rec.size() - 1;
    int first = cur;
    while (first > 0 && !b->rec[first].irrev_in) --first;
    int first_frames = cur - 7 < 0 ? 0 : cur - 7;
    if (first_frames < first) first = first_frames;
    int n = cur - first + 1;
    if (n > cap) return n;
    Pos* out = (Pos*)pos_out;
    for (int i = 0; i < n; ++i) out[i] = b->rec[first + i];
    // barrier only if we cut at an irreversible move or at the root; otherwise frames older than the
    // barrier are kept for the encoder and the scan stops at the real barrier anyway
    return n;
}

// The same for `n` boards in one call (roots of a batched search): records concatenated in pos_out, rec_len[i] positions
// each, n_legal[i] = number of legal moves of board i (size of its root noise vector, mcts.py:158).  Returns the total
// number of positions, or the required total if cap is too small (nothing is written past cap).
int cb_boards_export_records(const cb_board* const* boards, int n, void* pos_out, int cap, int32_t* rec_len, int32_t* n_legal) {
    int total = 0;
    for (int i = 0; i < n; ++i) {
        const int k = cb_board_export_record(boards[i], total < cap ? (Pos*)pos_out + total : nullptr, cap > total ? cap - total : 0);
        rec_len[i] = k;
        total += k;
        if (n_legal) {
            MoveList ml;
            generate_legal(boards[i]->rec.back(), ml);
            n_legal[i] = ml.n;
        }
    }
    return total;
}

int cb_move_to_action(uint16_t move, int player_white) { return move_to_action(move, player_white ? WHITE : BLACK); }
int cb_uci_to_action(const char* uci, int player_white) {
    Move m;
    if (parse_uci(uci, m) != 0) { cb_set_error(std::string("invalid uci: ") + uci); return -2; }
    return move_to_action(m, player_white ? WHITE : BLACK);
}
int cb_move_to_uci(uint16_t move, char out[6]) {
    int f = move_from(move), t = move_to(move), p = move_promo(move);
    out[0] = (char)('a' + (f & 7)); out[1] = (char)('1' + (f >> 3));
    out[2] = (char)('a' + (t & 7)); out[3] = (char)('1' + (t >> 3));
    int n = 4;
    if (p) out[n++] = " \0nbrq"[p] ? "  nbrq"[p] : 'q';
    out[n] = 0;
    return n;
}
}

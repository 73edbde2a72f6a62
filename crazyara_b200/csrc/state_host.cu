// Host-side game state (control plane): the BoardState of the reference (engine/src/environments/chess_related/
// boardstate.{h,cpp}) -- a position plus the (key, repetition) history Stockfish keeps in its StateInfo chain.
// Used by the UCI front-end / self-play driver to follow "position ... moves ..." and to hand roots to the GPU search;
// nothing here is on the hot path (expansion, move generation and plane encoding of the search run on the device).
#include <memory>
#include <string>
#include <vector>

#include "abi_common.h"
#include "ara_b200.h"
#include "chess_host.h"

namespace ara {
struct HostState {
    Board b;
    std::vector<uint64_t> keys;
    std::vector<int16_t> reps;
};
}  // namespace ara
using namespace ara;

extern "C" ara_state_t ara_state_create(const char* fen, int variant, int is960) {
    if (variant < 0 || variant > V_THREECHECK) {
        set_error("ara_state_create: unsupported variant %d", variant);
        return nullptr;
    }
    std::unique_ptr<HostState> s(new HostState());
    const char* f = (fen == nullptr || fen[0] == 0) ? start_fen(variant) : fen;
    if (!board_from_fen(&s->b, f, variant, is960)) {
        set_error("ara_state_create: cannot parse FEN '%s'", f);
        return nullptr;
    }
    return reinterpret_cast<ara_state_t>(s.release());
}
extern "C" ara_state_t ara_state_clone(ara_state_t h) {
    if (h == nullptr) return nullptr;
    return reinterpret_cast<ara_state_t>(new HostState(*reinterpret_cast<HostState*>(h)));
}
extern "C" void ara_state_destroy(ara_state_t h) { delete reinterpret_cast<HostState*>(h); }

extern "C" int ara_state_do_move(ara_state_t h, unsigned short move) {
    if (h == nullptr) return set_error("ara_state_do_move: null state");
    HostState* s = reinterpret_cast<HostState*>(h);
    bool legal = false;
    for (Move m : legal_moves_host(s->b)) legal = legal || m == move;
    if (!legal) return set_error("ara_state_do_move: move 0x%04x is not legal in %s", move, board_to_fen(s->b).c_str());
    s->keys.push_back(s->b.key);
    s->reps.push_back(s->b.repetition);
    do_move(s->b, move);
    s->b.repetition = static_cast<int16_t>(
        repetition_from_history(s->b, s->keys.data(), s->reps.data(), static_cast<int>(s->keys.size())));
    return 0;
}
extern "C" int ara_state_do_uci(ara_state_t h, const char* uci) {
    if (h == nullptr || uci == nullptr) return set_error("ara_state_do_uci: null argument");
    HostState* s = reinterpret_cast<HostState*>(h);
    const Move m = uci_to_move(s->b, uci);
    if (m == 0) return set_error("ara_state_do_uci: '%s' is not legal in %s", uci, board_to_fen(s->b).c_str());
    return ara_state_do_move(h, m);
}
extern "C" int ara_state_board(ara_state_t h, ara_board_t* out) {
    if (h == nullptr || out == nullptr) return set_error("ara_state_board: null argument");
    memcpy(out, &reinterpret_cast<HostState*>(h)->b, sizeof(Board));
    return 0;
}
extern "C" int ara_state_history(ara_state_t h, const unsigned long long** keys, const short** reps, int* len) {
    if (h == nullptr) return set_error("ara_state_history: null state");
    HostState* s = reinterpret_cast<HostState*>(h);
    if (keys) *keys = reinterpret_cast<const unsigned long long*>(s->keys.data());
    if (reps) *reps = s->reps.data();
    if (len) *len = static_cast<int>(s->keys.size());
    return 0;
}
extern "C" int ara_state_fen(ara_state_t h, char* buf, int buf_len) {
    if (h == nullptr) return set_error("ara_state_fen: null state");
    const std::string f = board_to_fen(reinterpret_cast<HostState*>(h)->b);
    if (static_cast<int>(f.size()) + 1 > buf_len) return set_error("ara_state_fen: buffer too small");
    memcpy(buf, f.c_str(), f.size() + 1);
    return 0;
}
extern "C" int ara_state_legal_moves(ara_state_t h, unsigned short* moves_out) {
    if (h == nullptr) return set_error("ara_state_legal_moves: null state");
    const std::vector<Move> mv = legal_moves_host(reinterpret_cast<HostState*>(h)->b);
    for (size_t i = 0; i < mv.size(); ++i) moves_out[i] = mv[i];
    return static_cast<int>(mv.size());
}
extern "C" int ara_state_side_to_move(ara_state_t h) { return h ? reinterpret_cast<HostState*>(h)->b.stm : -1; }
extern "C" int ara_state_in_check(ara_state_t h) {
    if (h == nullptr) return set_error("ara_state_in_check: null state");
    return in_check(reinterpret_cast<HostState*>(h)->b) ? 1 : 0;
}
extern "C" int ara_state_move_to_san(ara_state_t h, unsigned short move, int leads_to_win, char* buf16) {
    if (h == nullptr || buf16 == nullptr) return set_error("ara_state_move_to_san: null argument");
    const Board& b = reinterpret_cast<HostState*>(h)->b;
    const std::vector<Move> legal = legal_moves_host(b);
    bool found = false;
    for (Move m : legal) found = found || m == move;
    if (!found) return set_error("ara_state_move_to_san: move 0x%04x is not legal here", move);
    const std::string s = move_to_san(b, move, legal, leads_to_win != 0);
    memcpy(buf16, s.c_str(), s.size() + 1);
    return 0;
}
extern "C" int ara_state_is_terminal(ara_state_t h) {
    if (h == nullptr) return set_error("ara_state_is_terminal: null state");
    const Board& b = reinterpret_cast<HostState*>(h)->b;
    const int n = static_cast<int>(legal_moves_host(b).size());
    return terminal_type(b, n, in_check(b));
}

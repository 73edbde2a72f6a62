// Host emulation of the DEVICE rules / plane / search code (1-lane warp) for CPU-only unit tests.
// Test scaffolding: compiled by tests/ with g++, never part of the product library.
#include <cstring>
#include <string>
#include <vector>

#include "chess_host.h"
#include "planes_dev.cuh"

using namespace ara;

struct HeState {
    Board b;
    std::vector<uint64_t> keys;
    std::vector<int16_t> reps;
};

extern "C" {

HeState* he_new(const char* fen, int variant, int is960) {
    HeState* s = new HeState();
    if (!board_from_fen(&s->b, fen, variant, is960)) {
        delete s;
        return nullptr;
    }
    return s;
}
void he_free(HeState* s) { delete s; }
HeState* he_clone(const HeState* s) { return new HeState(*s); }
int he_legal_moves(const HeState* s, uint16_t* out) {
    std::vector<Move> mv = legal_moves_host(s->b);
    for (size_t i = 0; i < mv.size(); ++i) out[i] = mv[i];
    return static_cast<int>(mv.size());
}
void he_move_uci(const HeState* s, uint16_t m, char* buf) { strcpy(buf, move_to_uci(m, s->b.chess960 != 0).c_str()); }
uint16_t he_uci_move(const HeState* s, const char* uci) { return uci_to_move(s->b, uci); }
void he_do_move(HeState* s, uint16_t m) {
    s->keys.push_back(s->b.key);
    s->reps.push_back(s->b.repetition);
    do_move(s->b, m);
    s->b.repetition = static_cast<int16_t>(
        repetition_from_history(s->b, s->keys.data(), s->reps.data(), static_cast<int>(s->keys.size())));
}
void he_fen(const HeState* s, char* buf) { strcpy(buf, board_to_fen(s->b).c_str()); }
unsigned long long he_key(const HeState* s) { return s->b.key; }
unsigned long long he_key_scratch(const HeState* s) { return compute_key(s->b); }
int he_in_check(const HeState* s) { return in_check(s->b) ? 1 : 0; }
int he_repetition(const HeState* s) { return s->b.repetition; }
int he_terminal(const HeState* s) {
    Move scratch[kMaxMoves];
    int n = 0;
    const bool any = has_legal_move(s->b, scratch, &n);
    return terminal_type(s->b, any ? 1 : 0, in_check(s->b));
}
int he_policy_index(const HeState* s, uint16_t m) { return policy_map_index(m, s->b.stm, s->b.chess960); }
int he_planes(const HeState* s, int mode, int version, int normalize, float* out) {
    const int c = planes_channels(mode, version);
    if (c < 0) return -1;
    NchwF32Writer w{out};
    encode_planes(s->b, mode, version, normalize != 0, w);
    return c;
}
int he_sizeof_board() { return static_cast<int>(sizeof(Board)); }
const void* he_board(const HeState* s) { return &s->b; }
}

// Test infrastructure: C entry points over two more pieces of the UNMODIFIED reference that compile from their own
// sources -- the PGN writer (engine/src/rl/gamepgn.{h,cpp}) and the chess960 start-position generator
// (engine/src/environments/chess_related/chess960position.h).  Built by `make -C oracle ref` into oracle/_ref/; used only
// by tests/ to pin crazyara_b200/pgn.py::GamePGN and crazyara_b200/selfplay.py::chess960_fen.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>

#include "environments/chess_related/chess960position.h"
#include "rl/gamepgn.h"

extern "C" int ref_pgn_render(const char* const* header10, const char* const* moves, int n_moves, char* out, int cap) {
    GamePGN g;
    g.variant = header10[0];
    g.event = header10[1];
    g.date = header10[2];
    g.site = header10[3];
    g.round = header10[4];
    g.fen = header10[5];
    g.white = header10[6];
    g.black = header10[7];
    g.result = header10[8];
    g.timeControl = header10[9];
    for (int i = 0; i < n_moves; ++i) g.gameMoves.emplace_back(moves[i]);
    std::ostringstream os;
    os << g;
    const std::string s = os.str();
    if (static_cast<int>(s.size()) + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return static_cast<int>(s.size());
}

// the start position the reference draws after srand(seed) (its generator reads the C library's rand())
extern "C" void ref_chess960_fen(unsigned seed, char* out128) {
    srand(seed);
    const std::string fen = chess960fen();
    strncpy(out128, fen.c_str(), 127);
    out128[127] = 0;
}

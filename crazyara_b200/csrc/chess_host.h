// Host-only helpers around ara::Board: FEN parsing / printing and UCI move strings (State::set, State::fen,
// StateConstants::action_to_uci, State::uci_to_action of the reference: engine/src/state.h:287-509,
// environments/chess_related/boardstate.cpp:61-141).  Header-only; used by the C-ABI and the C++ host classes.
#pragma once
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "chess_dev.cuh"

namespace ara {

inline const char* start_fen(int variant) {  // boardstate.h:312-375
    switch (variant) {
        case V_CRAZYHOUSE: return "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR[] w KQkq - 0 1";
        case V_THREECHECK: return "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 3+3 0 1";
        default: return "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1";
    }
}

inline bool board_from_fen(Board* b, const char* fen, int variant, int is960) {
    memset(b, 0, sizeof(*b));
    b->variant = static_cast<uint8_t>(variant);
    b->chess960 = static_cast<uint8_t>(is960 ? 1 : 0);
    b->ep = 0xFF;
    for (int i = 0; i < 4; ++i) b->castle_rook[i] = 0xFF;
    static const char* kPieces = "PNBRQK";
    const char* c = fen;
    while (*c == ' ') ++c;
    int r = 7, f = 0, last_sq = -1;
    bool pocket = false;
    for (; *c && *c != ' '; ++c) {
        const char ch = *c;
        if (ch == '[') { pocket = true; continue; }
        if (ch == ']') { pocket = false; continue; }
        if (ch == '/') {
            --r;
            f = 0;
            if (r < 0) pocket = true;
            continue;
        }
        if (ch == '~') {
            if (last_sq >= 0) b->promoted |= bit(last_sq);
            continue;
        }
        if (isdigit(static_cast<unsigned char>(ch)) && !pocket) {
            f += ch - '0';
            continue;
        }
        if (ch == '-' && pocket) continue;
        const char* q = strchr(kPieces, toupper(static_cast<unsigned char>(ch)));
        if (q == nullptr) return false;
        const int pt = static_cast<int>(q - kPieces);
        const int col = islower(static_cast<unsigned char>(ch)) ? 1 : 0;
        if (pocket) {
            if (pt < 5) b->hand[col][pt]++;
            continue;
        }
        if (r < 0 || f > 7) return false;
        last_sq = r * 8 + f;
        b->by_type[pt] |= bit(last_sq);
        b->by_color[col] |= bit(last_sq);
        ++f;
    }
    while (*c == ' ') ++c;
    b->stm = (*c == 'b') ? 1 : 0;
    if (*c) ++c;
    while (*c == ' ') ++c;
    for (; *c && *c != ' '; ++c) {
        const char ch = *c;
        if (ch == '-') continue;
        const int col = islower(static_cast<unsigned char>(ch)) ? 1 : 0;
        const int rank = col ? 7 : 0;
        const int ks = king_square(*b, col);
        if (ks < 0 || (ks >> 3) != rank) continue;
        const uint64_t rooks = pieces(*b, col, PT_ROOK);
        int rs = -1;
        const char up = static_cast<char>(toupper(static_cast<unsigned char>(ch)));
        if (up == 'K') {
            for (int ff = 7; ff > (ks & 7); --ff)
                if (rooks & bit(rank * 8 + ff)) { rs = rank * 8 + ff; break; }
        } else if (up == 'Q') {
            for (int ff = 0; ff < (ks & 7); ++ff)
                if (rooks & bit(rank * 8 + ff)) { rs = rank * 8 + ff; break; }
        } else if (up >= 'A' && up <= 'H') {
            if (rooks & bit(rank * 8 + (up - 'A'))) rs = rank * 8 + (up - 'A');
        }
        if (rs < 0) continue;
        b->castle_rook[col * 2 + (rs > ks ? 0 : 1)] = static_cast<uint8_t>(rs);
    }
    while (*c == ' ') ++c;
    if (*c && *c != '-') {
        if (c[0] >= 'a' && c[0] <= 'h' && (c[1] == '3' || c[1] == '6')) {
            const int eps = (c[1] - '1') * 8 + (c[0] - 'a');
            const int them = b->stm ^ 1;
            const uint64_t attackers = pawn_attacks_bb(bit(eps), them) & pieces(*b, b->stm, PT_PAWN);
            const int pushed = b->stm ? eps + 8 : eps - 8;
            if (attackers && (pieces(*b, them, PT_PAWN) & bit(pushed)) && !(occupied(*b) & bit(eps)))
                b->ep = static_cast<uint8_t>(eps);
        }
        while (*c && *c != ' ') ++c;
    } else if (*c) {
        ++c;
    }
    while (*c == ' ') ++c;
    {
        const char* e = c;
        while (*e && *e != ' ') ++e;
        const char* plus = static_cast<const char*>(memchr(c, '+', static_cast<size_t>(e - c)));
        if (plus != nullptr && plus > c) {
            int gw = 3 - atoi(c), gb = 3 - atoi(plus + 1);
            if (gw < 0) gw = 0;
            if (gb < 0) gb = 0;
            b->checks = static_cast<uint8_t>((gw & 3) | ((gb & 3) << 2));
            c = e;
            while (*c == ' ') ++c;
        }
    }
    int half = 0, full = 1;
    if (*c) {
        half = atoi(c);
        while (*c && *c != ' ') ++c;
        while (*c == ' ') ++c;
        if (*c) full = atoi(c);
    }
    b->rule50 = static_cast<uint8_t>(half > 255 ? 255 : half);
    int ply = 2 * (full - 1);
    if (ply < 0) ply = 0;
    b->game_ply = static_cast<uint16_t>(ply + b->stm);
    b->key = compute_key(*b);
    return true;
}

inline std::string board_to_fen(const Board& b) {
    static const char* kW = "PNBRQK";
    static const char* kB = "pnbrqk";
    std::string o;
    for (int r = 7; r >= 0; --r) {
        int empty = 0;
        for (int f = 0; f < 8; ++f) {
            const int sq = r * 8 + f;
            const int pt = piece_type_on(b, sq);
            if (pt < 0) {
                ++empty;
                continue;
            }
            if (empty) o += static_cast<char>('0' + empty), empty = 0;
            o += (b.by_color[1] & bit(sq)) ? kB[pt] : kW[pt];
            if (b.variant == V_CRAZYHOUSE && (b.promoted & bit(sq))) o += '~';
        }
        if (empty) o += static_cast<char>('0' + empty);
        if (r) o += '/';
    }
    if (b.variant == V_CRAZYHOUSE) {
        o += '[';
        for (int c = 0; c < 2; ++c)
            for (int pt = 4; pt >= 0; --pt)
                for (int k = 0; k < b.hand[c][pt]; ++k) o += c ? kB[pt] : kW[pt];
        o += ']';
    }
    o += b.stm ? " b " : " w ";
    bool any = false;
    for (int i = 0; i < 4; ++i) {
        if (b.castle_rook[i] == 0xFF) continue;
        any = true;
        char ch = b.chess960 ? static_cast<char>('A' + (b.castle_rook[i] & 7)) : ((i & 1) ? 'Q' : 'K');
        if (i >= 2) ch = static_cast<char>(tolower(ch));
        o += ch;
    }
    if (!any) o += '-';
    o += ' ';
    if (b.ep != 0xFF) {
        o += static_cast<char>('a' + (b.ep & 7));
        o += static_cast<char>('1' + (b.ep >> 3));
    } else {
        o += '-';
    }
    char buf[64];
    if (b.variant == V_THREECHECK) {
        snprintf(buf, sizeof(buf), " %d+%d", 3 - checks_given(b, 0), 3 - checks_given(b, 1));
        o += buf;
    }
    snprintf(buf, sizeof(buf), " %d %d", b.rule50, 1 + (b.game_ply - b.stm) / 2);
    o += buf;
    return o;
}

inline std::string move_to_uci(Move m, bool chess960) {  // UCI::move of the engine's Stockfish fork
    static const char* kW = "PNBRQK";
    const int from = mv_from(m), flag = mv_flag(m);
    int to = mv_to(m);
    std::string s;
    if (flag >= MF_DROP) {
        s += kW[flag - MF_DROP];
        s += '@';
        s += static_cast<char>('a' + (to & 7));
        s += static_cast<char>('1' + (to >> 3));
        return s;
    }
    if (flag == MF_CASTLE && !chess960) to = (from & 56) | ((to & 7) > (from & 7) ? 6 : 2);
    s += static_cast<char>('a' + (from & 7));
    s += static_cast<char>('1' + (from >> 3));
    s += static_cast<char>('a' + (to & 7));
    s += static_cast<char>('1' + (to >> 3));
    if (flag >= MF_PROMO_N && flag <= MF_PROMO_Q) s += "nbrq"[flag - 1];
    return s;
}

inline std::vector<Move> legal_moves_host(const Board& b) {
    Move scratch[kMaxMoves], out[kMaxMoves];
    MoveGenScratch mg;
    const int n = gen_legal(b, mg, scratch, out);
    return std::vector<Move>(out, out + n);
}

inline Move uci_to_move(const Board& b, const std::string& uci) {
    for (Move m : legal_moves_host(b))
        if (move_to_uci(m, b.chess960 != 0) == uci) return m;
    return 0;
}

// pgn_move / is_pgn_move_ambiguous (environments/chess_related/board.cpp:277-385): the SAN spelling the reference
// writes into its PGN files -- promotions without '=', the origin file on pawn captures, disambiguation by file unless
// another candidate shares the file (then the rank; both: the whole square), '+' / '#' from gives_check.
inline std::string move_to_san(const Board& b, Move m, const std::vector<Move>& legal, bool leads_to_win) {
    static const char* kW = "PNBRQK";
    if (m == 0) return "(none)";
    const int from = mv_from(m), to = mv_to(m), flag = mv_flag(m);
    auto square = [](int s) { return std::string{static_cast<char>('a' + (s & 7)), static_cast<char>('1' + (s >> 3))}; };
    std::string out;
    if (flag >= MF_DROP) {
        out = std::string{kW[flag - MF_DROP], '@'} + square(to);
    } else if (flag == MF_CASTLE) {
        out = (from & 7) < (to & 7) ? "O-O" : "O-O-O";
    } else {
        const int pt = piece_type_on(b, from);
        bool ambiguous = false, same_file = false, same_rank = false;
        for (Move o : legal) {
            if (mv_flag(o) >= MF_DROP) continue;
            const int of = mv_from(o);
            if (mv_to(o) == to && of != from && piece_type_on(b, of) == pt) {
                ambiguous = true;
                if ((of & 7) == (from & 7)) same_file = true;
                if ((of >> 3) == (from >> 3)) same_rank = true;
            }
        }
        std::string origin;
        if (ambiguous)
            origin = same_file && same_rank ? square(from)
                                            : (same_file ? std::string{static_cast<char>('1' + (from >> 3))}
                                                         : std::string{static_cast<char>('a' + (from & 7))});
        const bool capture = flag == MF_EP || ((b.by_color[0] | b.by_color[1]) & bit(to)) != 0;
        if (pt == PT_PAWN)
            out = capture ? std::string{static_cast<char>('a' + (from & 7)), 'x'} + square(to) : square(to);
        else
            out = std::string{kW[pt]} + origin + (capture ? "x" : "") + square(to);
        if (flag >= MF_PROMO_N && flag <= MF_PROMO_Q) out += kW[flag];
    }
    Board after = b;
    do_move(after, m);
    if (in_check(after)) out += leads_to_win ? "#" : "+";
    return out;
}

}  // namespace ara

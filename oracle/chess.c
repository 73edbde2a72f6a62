/* oracle/chess.c -- CPU ORACLE (test infrastructure only).  See oracle/chess.h for what it stands in for.
 *
 * Deliberately simple: mailbox board, ray walking, pseudo-legal generation filtered by "apply the move on a copy of
 * the board and look whether the own king is attacked".  The product's device-side generator
 * (crazyara_b200/csrc/chess_dev.cuh) is written independently with bitboards; the two are compared move-set by
 * move-set in tests/.
 */
#include "chess.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define COLOR_OF(pc) ((pc) >> 3)
#define TYPE_OF(pc) ((pc) & 7)
#define MAKE_PC(c, t) (((c) << 3) | (t))
#define RANK(s) ((s) >> 3)
#define FILE_(s) ((s) & 7)
#define SQ(r, f) ((r) * 8 + (f))

static const int KNIGHT_D[8][2] = {{2, 1}, {1, 2}, {-1, 2}, {-2, 1}, {-2, -1}, {-1, -2}, {1, -2}, {2, -1}};
static const int KING_D[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
static const int ROOK_D[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
static const int BISHOP_D[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};

/* ------------------------------------------------------------------ Zobrist (shared definition with the product) */
uint64_t opos_zobrist(int idx) {
    uint64_t z = (uint64_t)idx * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
void opos_init_tables(void) {}
size_t opos_sizeof(void) { return sizeof(OPos); }

uint64_t opos_compute_key(const OPos* p) {
    uint64_t k = 0;
    for (int s = 0; s < 64; ++s)
        if (p->board[s]) k ^= opos_zobrist(p->board[s] * 64 + s);
    if (p->stm) k ^= opos_zobrist(1024);
    for (int r = 0; r < 4; ++r)
        if (p->castle_rook[r] >= 0) k ^= opos_zobrist(1025 + r);
    if (p->ep >= 0) k ^= opos_zobrist(1029 + FILE_(p->ep));
    if (p->variant == OV_CRAZYHOUSE)
        for (int c = 0; c < 2; ++c)
            for (int t = 1; t <= 5; ++t)
                if (p->hand[c][t] > 0) k ^= opos_zobrist(1040 + (c * 8 + t) * 32 + (p->hand[c][t] & 31));
    if (p->variant == OV_THREECHECK)
        for (int c = 0; c < 2; ++c)
            if (p->checks_given[c] > 0) k ^= opos_zobrist(1600 + c * 4 + (p->checks_given[c] & 3));
    return k;
}

/* ------------------------------------------------------------------ attacks on a raw board */
static int on_board(int r, int f) { return r >= 0 && r < 8 && f >= 0 && f < 8; }

static int sq_attacked(const int8_t* b, int sq, int by) {
    const int r = RANK(sq), f = FILE_(sq);
    /* pawns: a white pawn on (r-1, f+-1) attacks sq */
    {
        const int pr = by == 0 ? r - 1 : r + 1;
        for (int df = -1; df <= 1; df += 2)
            if (on_board(pr, f + df) && b[SQ(pr, f + df)] == MAKE_PC(by, OP_PAWN)) return 1;
    }
    for (int i = 0; i < 8; ++i) {
        int rr = r + KNIGHT_D[i][0], ff = f + KNIGHT_D[i][1];
        if (on_board(rr, ff) && b[SQ(rr, ff)] == MAKE_PC(by, OP_KNIGHT)) return 1;
        rr = r + KING_D[i][0];
        ff = f + KING_D[i][1];
        if (on_board(rr, ff) && b[SQ(rr, ff)] == MAKE_PC(by, OP_KING)) return 1;
    }
    for (int i = 0; i < 4; ++i) {
        int rr = r + ROOK_D[i][0], ff = f + ROOK_D[i][1];
        while (on_board(rr, ff)) {
            const int pc = b[SQ(rr, ff)];
            if (pc) {
                if (pc == MAKE_PC(by, OP_ROOK) || pc == MAKE_PC(by, OP_QUEEN)) return 1;
                break;
            }
            rr += ROOK_D[i][0];
            ff += ROOK_D[i][1];
        }
        rr = r + BISHOP_D[i][0];
        ff = f + BISHOP_D[i][1];
        while (on_board(rr, ff)) {
            const int pc = b[SQ(rr, ff)];
            if (pc) {
                if (pc == MAKE_PC(by, OP_BISHOP) || pc == MAKE_PC(by, OP_QUEEN)) return 1;
                break;
            }
            rr += BISHOP_D[i][0];
            ff += BISHOP_D[i][1];
        }
    }
    return 0;
}

/* bitboard of pieces of colour `by` attacking sq (Position::attackers_to restricted to one colour) */
static uint64_t attackers_bb(const int8_t* b, int sq, int by) {
    uint64_t bb = 0;
    const int r = RANK(sq), f = FILE_(sq);
    const int pr = by == 0 ? r - 1 : r + 1;
    for (int df = -1; df <= 1; df += 2)
        if (on_board(pr, f + df) && b[SQ(pr, f + df)] == MAKE_PC(by, OP_PAWN)) bb |= 1ULL << SQ(pr, f + df);
    for (int i = 0; i < 8; ++i) {
        int rr = r + KNIGHT_D[i][0], ff = f + KNIGHT_D[i][1];
        if (on_board(rr, ff) && b[SQ(rr, ff)] == MAKE_PC(by, OP_KNIGHT)) bb |= 1ULL << SQ(rr, ff);
        rr = r + KING_D[i][0];
        ff = f + KING_D[i][1];
        if (on_board(rr, ff) && b[SQ(rr, ff)] == MAKE_PC(by, OP_KING)) bb |= 1ULL << SQ(rr, ff);
    }
    for (int i = 0; i < 4; ++i) {
        for (int diag = 0; diag < 2; ++diag) {
            const int dr = diag ? BISHOP_D[i][0] : ROOK_D[i][0], df = diag ? BISHOP_D[i][1] : ROOK_D[i][1];
            int rr = r + dr, ff = f + df;
            while (on_board(rr, ff)) {
                const int pc = b[SQ(rr, ff)];
                if (pc) {
                    if (pc == MAKE_PC(by, diag ? OP_BISHOP : OP_ROOK) || pc == MAKE_PC(by, OP_QUEEN)) bb |= 1ULL << SQ(rr, ff);
                    break;
                }
                rr += dr;
                ff += df;
            }
        }
    }
    return bb;
}

static int king_sq(const int8_t* b, int c) {
    for (int s = 0; s < 64; ++s)
        if (b[s] == MAKE_PC(c, OP_KING)) return s;
    return -1;
}

int opos_in_check(const OPos* p) {
    const int k = king_sq(p->board, p->stm);
    return k >= 0 && sq_attacked(p->board, k, p->stm ^ 1);
}
uint64_t opos_checkers_bb(const OPos* p) {
    const int k = king_sq(p->board, p->stm);
    return k < 0 ? 0 : attackers_bb(p->board, k, p->stm ^ 1);
}

uint64_t opos_pieces_bb(const OPos* p, int color, int pt) {
    uint64_t bb = 0;
    for (int s = 0; s < 64; ++s) {
        const int pc = p->board[s];
        if (pc && COLOR_OF(pc) == color && (pt == 0 || TYPE_OF(pc) == pt)) bb |= 1ULL << s;
    }
    return bb;
}
int opos_count(const OPos* p, int color, int pt) { return __builtin_popcountll(opos_pieces_bb(p, color, pt)); }
int opos_can_castle(const OPos* p, int right) { return p->castle_rook[right] >= 0; }

/* ------------------------------------------------------------------ applying a move to a raw board */
static void castle_squares(int us, int king_side, int* kto, int* rto) {
    const int r = us ? 7 : 0;
    *kto = SQ(r, king_side ? 6 : 2);
    *rto = SQ(r, king_side ? 5 : 3);
}

static void apply_board(int8_t* b, uint32_t m, int us) {
    const int from = OM_FROM(m), to = OM_TO(m), type = OM_TYPE(m);
    if (type == OM_DROP) {
        b[to] = (int8_t)MAKE_PC(us, OM_PT(m));
        return;
    }
    if (type == OM_CASTLING) {
        int kto, rto;
        castle_squares(us, FILE_(to) > FILE_(from), &kto, &rto);
        b[from] = 0;
        b[to] = 0;
        b[kto] = (int8_t)MAKE_PC(us, OP_KING);
        b[rto] = (int8_t)MAKE_PC(us, OP_ROOK);
        return;
    }
    int pc = b[from];
    if (type == OM_ENPASSANT) b[SQ(RANK(from), FILE_(to))] = 0;
    if (type == OM_PROMOTION) pc = MAKE_PC(us, OM_PT(m));
    b[from] = 0;
    b[to] = (int8_t)pc;
}

/* ------------------------------------------------------------------ pseudo-legal generation */
static int variant_end(const OPos* p);

static int gen_pseudo(const OPos* p, uint32_t* out) {
    int n = 0;
    const int us = p->stm, them = us ^ 1;
    const int8_t* b = p->board;
    for (int s = 0; s < 64; ++s) {
        const int pc = b[s];
        if (!pc || COLOR_OF(pc) != us) continue;
        const int r = RANK(s), f = FILE_(s);
        switch (TYPE_OF(pc)) {
            case OP_PAWN: {
                const int dr = us ? -1 : 1, start = us ? 6 : 1, promo = us ? 1 : 6;
                const int r1 = r + dr;
                if (!on_board(r1, f)) break;
                if (!b[SQ(r1, f)]) {
                    if (r == promo) {
                        for (int t = OP_QUEEN; t >= OP_KNIGHT; --t) out[n++] = OMOVE(s, SQ(r1, f), OM_PROMOTION, t);
                    } else {
                        out[n++] = OMOVE(s, SQ(r1, f), OM_NORMAL, 0);
                        if (r == start && !b[SQ(r + 2 * dr, f)]) out[n++] = OMOVE(s, SQ(r + 2 * dr, f), OM_NORMAL, 0);
                    }
                }
                for (int df = -1; df <= 1; df += 2) {
                    if (!on_board(r1, f + df)) continue;
                    const int t = SQ(r1, f + df);
                    if (b[t] && COLOR_OF(b[t]) == them) {
                        if (r == promo) {
                            for (int q = OP_QUEEN; q >= OP_KNIGHT; --q) out[n++] = OMOVE(s, t, OM_PROMOTION, q);
                        } else {
                            out[n++] = OMOVE(s, t, OM_NORMAL, 0);
                        }
                    } else if (t == p->ep && p->ep >= 0 && !b[t]) {
                        out[n++] = OMOVE(s, t, OM_ENPASSANT, 0);
                    }
                }
                break;
            }
            case OP_KNIGHT:
            case OP_KING: {
                const int(*d)[2] = TYPE_OF(pc) == OP_KNIGHT ? KNIGHT_D : KING_D;
                for (int i = 0; i < 8; ++i) {
                    const int rr = r + d[i][0], ff = f + d[i][1];
                    if (!on_board(rr, ff)) continue;
                    const int t = SQ(rr, ff);
                    if (!b[t] || COLOR_OF(b[t]) == them) out[n++] = OMOVE(s, t, OM_NORMAL, 0);
                }
                break;
            }
            default: {
                for (int i = 0; i < 4; ++i) {
                    for (int diag = 0; diag < 2; ++diag) {
                        if (TYPE_OF(pc) == OP_ROOK && diag) continue;
                        if (TYPE_OF(pc) == OP_BISHOP && !diag) continue;
                        const int dr = diag ? BISHOP_D[i][0] : ROOK_D[i][0], df = diag ? BISHOP_D[i][1] : ROOK_D[i][1];
                        int rr = r + dr, ff = f + df;
                        while (on_board(rr, ff)) {
                            const int t = SQ(rr, ff);
                            if (!b[t]) {
                                out[n++] = OMOVE(s, t, OM_NORMAL, 0);
                            } else {
                                if (COLOR_OF(b[t]) == them) out[n++] = OMOVE(s, t, OM_NORMAL, 0);
                                break;
                            }
                            rr += dr;
                            ff += df;
                        }
                    }
                }
            }
        }
    }
    /* castling (encoded king-from -> rook-from, as Stockfish does) */
    if (!opos_in_check(p)) {
        const int ks = king_sq(b, us);
        for (int side = 0; side < 2; ++side) {
            const int rs = p->castle_rook[us * 2 + side];
            if (rs < 0 || ks < 0) continue;
            int kto, rto;
            castle_squares(us, side == 0, &kto, &rto);
            int ok = 1;
            /* castling_impeded: squares between king/rook origins and destinations must be empty (bar king and rook) */
            {
                int lo = ks < kto ? ks : kto, hi = ks < kto ? kto : ks;
                for (int s = lo; s <= hi && ok; ++s)
                    if (s != ks && s != rs && b[s]) ok = 0;
                lo = rs < rto ? rs : rto;
                hi = rs < rto ? rto : rs;
                for (int s = lo; s <= hi && ok; ++s)
                    if (s != ks && s != rs && b[s]) ok = 0;
            }
            /* the king may not pass over an attacked square */
            if (ok) {
                const int step = kto > ks ? 1 : -1;
                for (int s = ks; s != kto; s += step)
                    if (s != ks && sq_attacked(b, s, them)) ok = 0;
                if (ok && kto != ks && sq_attacked(b, kto, them)) ok = 0;
            }
            if (ok) out[n++] = OMOVE(ks, rs, OM_CASTLING, 0);
        }
    }
    /* drops (crazyhouse) */
    if (p->variant == OV_CRAZYHOUSE) {
        for (int t = OP_PAWN; t <= OP_QUEEN; ++t) {
            if (p->hand[us][t] <= 0) continue;
            for (int s = 0; s < 64; ++s) {
                if (b[s]) continue;
                if (t == OP_PAWN && (RANK(s) == 0 || RANK(s) == 7)) continue;
                out[n++] = OMOVE(s, s, OM_DROP, t);
            }
        }
    }
    return n;
}

int opos_legal_moves(const OPos* p, uint32_t* out) {
    if (variant_end(p)) return 0;
    uint32_t tmp[OPOS_MAX_MOVES * 2];
    const int n = gen_pseudo(p, tmp);
    int k = 0;
    for (int i = 0; i < n; ++i) {
        int8_t b[64];
        memcpy(b, p->board, 64);
        apply_board(b, tmp[i], p->stm);
        const int ks = king_sq(b, p->stm);
        if (ks >= 0 && sq_attacked(b, ks, p->stm ^ 1)) continue;
        out[k++] = tmp[i];
    }
    return k;
}

int opos_gives_check(const OPos* p, uint32_t m) {
    int8_t b[64];
    memcpy(b, p->board, 64);
    apply_board(b, m, p->stm);
    const int ks = king_sq(b, p->stm ^ 1);
    return ks >= 0 && sq_attacked(b, ks, p->stm);
}

/* ------------------------------------------------------------------ do_move (Position::do_move + Board::do_move) */
static void push_hist(OPos* p) {
    if (p->hist_len == OPOS_MAX_HIST) {
        memmove(p->hist_key, p->hist_key + 256, (OPOS_MAX_HIST - 256) * sizeof(uint64_t));
        memmove(p->hist_rep, p->hist_rep + 256, (OPOS_MAX_HIST - 256) * sizeof(int16_t));
        p->hist_len -= 256;
    }
    p->hist_key[p->hist_len] = p->key;
    p->hist_rep[p->hist_len] = (int16_t)p->repetition;
    p->hist_len++;
}

void opos_do_move(OPos* p, uint32_t m) {
    const int us = p->stm, them = us ^ 1;
    const int from = OM_FROM(m), to = OM_TO(m), type = OM_TYPE(m);
    /* Board::add_move_to_list (board.cpp:216-225): most recent first, at most 8 */
    for (int i = (p->n_last < 8 ? p->n_last : 7); i > 0; --i) p->last_moves[i] = p->last_moves[i - 1];
    p->last_moves[0] = m;
    if (p->n_last < 8) p->n_last++;

    push_hist(p);
    p->rule50++;
    p->plies_from_null++;
    p->game_ply++;
    int new_ep = -1;

    if (type == OM_DROP) {
        const int pt = OM_PT(m);
        p->board[to] = (int8_t)MAKE_PC(us, pt);
        p->promoted[to] = 0;
        p->hand[us][pt]--;
        if (pt == OP_PAWN) p->rule50 = 0;
    } else if (type == OM_CASTLING) {
        int kto, rto;
        castle_squares(us, FILE_(to) > FILE_(from), &kto, &rto);
        p->board[from] = 0;
        p->board[to] = 0;
        p->promoted[from] = p->promoted[to] = 0;
        p->board[kto] = (int8_t)MAKE_PC(us, OP_KING);
        p->board[rto] = (int8_t)MAKE_PC(us, OP_ROOK);
        p->promoted[kto] = p->promoted[rto] = 0;
        p->castle_rook[us * 2] = p->castle_rook[us * 2 + 1] = -1;
    } else {
        const int pc = p->board[from];
        const int pt = TYPE_OF(pc);
        int capsq = to;
        if (type == OM_ENPASSANT) capsq = SQ(RANK(from), FILE_(to));
        const int captured = p->board[capsq];
        if (captured) {
            if (p->variant == OV_CRAZYHOUSE) p->hand[us][p->promoted[capsq] ? OP_PAWN : TYPE_OF(captured)]++;
            p->board[capsq] = 0;
            p->promoted[capsq] = 0;
            p->rule50 = 0;
            for (int r = 0; r < 2; ++r)
                if (p->castle_rook[them * 2 + r] == capsq) p->castle_rook[them * 2 + r] = -1;
        }
        const int was_promoted = p->promoted[from];
        p->board[from] = 0;
        p->promoted[from] = 0;
        p->board[to] = (int8_t)pc;
        p->promoted[to] = (uint8_t)was_promoted;
        if (pt == OP_PAWN) {
            p->rule50 = 0;
            if (type == OM_PROMOTION) {
                p->board[to] = (int8_t)MAKE_PC(us, OM_PT(m));
                p->promoted[to] = (p->variant == OV_CRAZYHOUSE) ? 1 : 0;
            } else if ((to ^ from) == 16) {
                /* en-passant square only if an enemy pawn attacks it (Position::do_move) */
                const int eps = (from + to) / 2;
                const int er = RANK(to), ef = FILE_(to);
                for (int df = -1; df <= 1; df += 2)
                    if (on_board(er, ef + df) && p->board[SQ(er, ef + df)] == MAKE_PC(them, OP_PAWN)) new_ep = eps;
            }
        }
        if (pt == OP_KING) p->castle_rook[us * 2] = p->castle_rook[us * 2 + 1] = -1;
        for (int r = 0; r < 2; ++r)
            if (p->castle_rook[us * 2 + r] == from) p->castle_rook[us * 2 + r] = -1;
    }
    p->ep = new_ep;
    p->stm = them;
    if (p->variant == OV_THREECHECK && opos_in_check(p)) p->checks_given[us]++;
    p->key = opos_compute_key(p);

    /* repetition info (Position::do_move): ply distance to the previous occurrence, negative for a 3-fold */
    p->repetition = 0;
    int end = p->variant == OV_CRAZYHOUSE ? p->plies_from_null : (p->rule50 < p->plies_from_null ? p->rule50 : p->plies_from_null);
    if (end > p->hist_len) end = p->hist_len;
    for (int i = 4; i <= end; i += 2) {
        const int idx = p->hist_len - i;
        if (p->hist_key[idx] == p->key) {
            p->repetition = p->hist_rep[idx] ? -i : i;
            break;
        }
    }
}

/* ------------------------------------------------------------------ FEN */
static const char PIECE_CHARS[] = " PNBRQK  pnbrqk";

const char* opos_start_fen(int variant) {
    switch (variant) {
        case OV_CRAZYHOUSE: return "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR[] w KQkq - 0 1";
        case OV_THREECHECK: return "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 3+3 0 1";
        case OV_ANTI: return "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w - - 0 1";
        case OV_HORDE: return "rnbqkbnr/pppppppp/8/1PP2PP1/PPPPPPPP/PPPPPPPP/PPPPPPPP/PPPPPPPP w kq - 0 1";
        case OV_RACE: return "8/8/8/8/8/8/krbnNBRK/qrbnNBRQ w - - 0 1";
        default: return "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1";
    }
}

int opos_set(OPos* p, const char* fen, int variant, int is960) {
    memset(p, 0, sizeof(*p));
    p->variant = variant;
    p->chess960 = is960;
    p->ep = -1;
    for (int i = 0; i < 4; ++i) p->castle_rook[i] = -1;
    const char* c = fen;
    while (*c == ' ') ++c;
    /* 1. piece placement (+ optional crazyhouse pocket as [..] or as a 9th rank) */
    int r = 7, f = 0, last_sq = -1, in_pocket = 0;
    for (; *c && *c != ' '; ++c) {
        const char ch = *c;
        if (ch == '[') { in_pocket = 1; continue; }
        if (ch == ']') { in_pocket = 0; continue; }
        if (ch == '/') {
            --r;
            f = 0;
            if (r < 0) in_pocket = 1;
            continue;
        }
        if (ch == '~') {
            if (last_sq >= 0) p->promoted[last_sq] = 1;
            continue;
        }
        if (isdigit((unsigned char)ch) && !in_pocket) {
            f += ch - '0';
            continue;
        }
        const char* q = strchr(PIECE_CHARS, ch);
        if (!q || ch == ' ') return -1;
        const int pc = (int)(q - PIECE_CHARS);
        if (in_pocket) {
            if (ch != '-') p->hand[COLOR_OF(pc)][TYPE_OF(pc)]++;
            continue;
        }
        if (r < 0 || f > 7) return -1;
        last_sq = SQ(r, f);
        p->board[last_sq] = (int8_t)pc;
        ++f;
    }
    while (*c == ' ') ++c;
    /* 2. side to move */
    p->stm = (*c == 'b');
    if (*c) ++c;
    while (*c == ' ') ++c;
    /* 3. castling */
    for (; *c && *c != ' '; ++c) {
        const char ch = *c;
        if (ch == '-') continue;
        const int col = islower((unsigned char)ch) ? 1 : 0;
        const int rank = col ? 7 : 0;
        const int rook = MAKE_PC(col, OP_ROOK);
        const int ks = king_sq(p->board, col);
        if (ks < 0 || RANK(ks) != rank) continue;
        int rs = -1;
        const char up = (char)toupper((unsigned char)ch);
        if (up == 'K') {
            for (int ff = 7; ff > FILE_(ks); --ff)
                if (p->board[SQ(rank, ff)] == rook) { rs = SQ(rank, ff); break; }
        } else if (up == 'Q') {
            for (int ff = 0; ff < FILE_(ks); ++ff)
                if (p->board[SQ(rank, ff)] == rook) { rs = SQ(rank, ff); break; }
        } else if (up >= 'A' && up <= 'H') {
            if (p->board[SQ(rank, up - 'A')] == rook) rs = SQ(rank, up - 'A');
        }
        if (rs < 0) continue;
        p->castle_rook[col * 2 + (rs > ks ? 0 : 1)] = rs;
    }
    while (*c == ' ') ++c;
    /* 4. en passant: kept only if a pawn of the side to move attacks it and the pushed pawn is there */
    if (*c && *c != '-') {
        if (c[0] >= 'a' && c[0] <= 'h' && (c[1] == '3' || c[1] == '6')) {
            const int eps = SQ(c[1] - '1', c[0] - 'a');
            const int them = p->stm ^ 1;
            const int pr = p->stm ? RANK(eps) + 1 : RANK(eps) - 1; /* rank of capturing pawns */
            int attacked = 0;
            for (int df = -1; df <= 1; df += 2)
                if (on_board(pr, FILE_(eps) + df) && p->board[SQ(pr, FILE_(eps) + df)] == MAKE_PC(p->stm, OP_PAWN)) attacked = 1;
            const int pushed = p->stm ? eps + 8 : eps - 8;
            if (attacked && p->board[pushed] == MAKE_PC(them, OP_PAWN) && !p->board[eps]) p->ep = eps;
        }
        while (*c && *c != ' ') ++c;
    } else if (*c) {
        ++c;
    }
    while (*c == ' ') ++c;
    /* 5. remaining checks "w+b" (three-check) */
    {
        const char* e = c;
        while (*e && *e != ' ') ++e;
        const char* plus = memchr(c, '+', (size_t)(e - c));
        if (plus && plus > c) {
            const int rw = atoi(c), rb = atoi(plus + 1);
            p->checks_given[0] = 3 - rw < 0 ? 0 : 3 - rw;
            p->checks_given[1] = 3 - rb < 0 ? 0 : 3 - rb;
            c = e;
            while (*c == ' ') ++c;
        }
    }
    /* 6. halfmove clock and fullmove number */
    int half = 0, full = 1;
    if (*c) {
        half = atoi(c);
        while (*c && *c != ' ') ++c;
        while (*c == ' ') ++c;
        if (*c) full = atoi(c);
    }
    p->rule50 = half;
    p->game_ply = 2 * (full - 1);
    if (p->game_ply < 0) p->game_ply = 0;
    p->game_ply += p->stm;
    p->key = opos_compute_key(p);
    return 0;
}

void opos_copy(OPos* dst, const OPos* src) { memcpy(dst, src, sizeof(OPos)); }

void opos_fen(const OPos* p, char* buf) {
    char* o = buf;
    for (int r = 7; r >= 0; --r) {
        int empty = 0;
        for (int f = 0; f < 8; ++f) {
            const int pc = p->board[SQ(r, f)];
            if (!pc) {
                ++empty;
                continue;
            }
            if (empty) *o++ = (char)('0' + empty), empty = 0;
            *o++ = PIECE_CHARS[pc];
            if (p->variant == OV_CRAZYHOUSE && p->promoted[SQ(r, f)]) *o++ = '~';
        }
        if (empty) *o++ = (char)('0' + empty);
        if (r) *o++ = '/';
    }
    if (p->variant == OV_CRAZYHOUSE) {
        *o++ = '[';
        for (int c = 0; c < 2; ++c)
            for (int t = OP_QUEEN; t >= OP_PAWN; --t)
                for (int k = 0; k < p->hand[c][t]; ++k) *o++ = PIECE_CHARS[MAKE_PC(c, t)];
        *o++ = ']';
    }
    *o++ = ' ';
    *o++ = p->stm ? 'b' : 'w';
    *o++ = ' ';
    int any = 0;
    for (int i = 0; i < 4; ++i) {
        if (p->castle_rook[i] < 0) continue;
        any = 1;
        char ch;
        if (p->chess960)
            ch = (char)('A' + FILE_(p->castle_rook[i]));
        else
            ch = (i & 1) ? 'Q' : 'K';
        if (i >= 2) ch = (char)tolower((unsigned char)ch);
        *o++ = ch;
    }
    if (!any) *o++ = '-';
    *o++ = ' ';
    if (p->ep >= 0) {
        *o++ = (char)('a' + FILE_(p->ep));
        *o++ = (char)('1' + RANK(p->ep));
    } else {
        *o++ = '-';
    }
    if (p->variant == OV_THREECHECK) o += sprintf(o, " %d+%d", 3 - p->checks_given[0], 3 - p->checks_given[1]);
    sprintf(o, " %d %d", p->rule50, 1 + (p->game_ply - p->stm) / 2);
}

/* ------------------------------------------------------------------ UCI strings (UCI::move / UCI::to_move) */
void opos_move_to_uci(const OPos* p, uint32_t m, char* buf) {
    const int from = OM_FROM(m), type = OM_TYPE(m);
    int to = OM_TO(m);
    if (type == OM_DROP) {
        sprintf(buf, "%c@%c%c", PIECE_CHARS[OM_PT(m)], 'a' + FILE_(to), '1' + RANK(to));
        return;
    }
    if (type == OM_CASTLING && !p->chess960) to = SQ(RANK(from), to > from ? 6 : 2);
    char* o = buf;
    *o++ = (char)('a' + FILE_(from));
    *o++ = (char)('1' + RANK(from));
    *o++ = (char)('a' + FILE_(to));
    *o++ = (char)('1' + RANK(to));
    if (type == OM_PROMOTION) *o++ = (char)tolower((unsigned char)PIECE_CHARS[OM_PT(m)]);
    *o = 0;
}

uint32_t opos_uci_to_move(const OPos* p, const char* uci) {
    uint32_t mv[OPOS_MAX_MOVES];
    const int n = opos_legal_moves(p, mv);
    char buf[8];
    for (int i = 0; i < n; ++i) {
        opos_move_to_uci(p, mv[i], buf);
        if (strcmp(buf, uci) == 0) return mv[i];
    }
    return 0;
}

/* ------------------------------------------------------------------ terminal rules */
static int koth_center(int s) { return s == 27 || s == 28 || s == 35 || s == 36; }

static int variant_end(const OPos* p) {
    if (p->variant == OV_KOTH) {
        const int k0 = king_sq(p->board, 0), k1 = king_sq(p->board, 1);
        return (k0 >= 0 && koth_center(k0)) || (k1 >= 0 && koth_center(k1));
    }
    if (p->variant == OV_THREECHECK) return p->checks_given[0] >= 3 || p->checks_given[1] >= 3;
    return 0;
}

int opos_number_repetitions(const OPos* p) { return p->repetition == 0 ? 0 : 1; } /* board.cpp:132-141 */

static int insufficient_material(const OPos* p) { /* board.cpp:170-213 */
    if (p->variant != OV_CHESS && p->variant != OV_ATOMIC) return 0;
    const int all = opos_count(p, 0, 0) + opos_count(p, 1, 0);
    if (all > 4) return 0;
    const int bishops = opos_count(p, 0, OP_BISHOP) + opos_count(p, 1, OP_BISHOP);
    const int knights = opos_count(p, 0, OP_KNIGHT) + opos_count(p, 1, OP_KNIGHT);
    return all == 2 || (all == 3 && bishops == 1) || (all == 3 && knights == 1) ||
           (all == 4 && (opos_count(p, 0, OP_KNIGHT) == 2 || opos_count(p, 1, OP_KNIGHT) == 2));
}

/* BoardState::is_terminal (boardstate.cpp:143-226) */
int opos_is_terminal(const OPos* p, int n_legal) {
    if (p->variant == OV_KOTH) {
        const int ku = king_sq(p->board, p->stm), kt = king_sq(p->board, p->stm ^ 1);
        if (ku >= 0 && koth_center(ku)) return OT_WIN;
        if (kt >= 0 && koth_center(kt)) return OT_LOSS;
    }
    if (p->variant == OV_THREECHECK) {
        if (p->checks_given[p->stm] >= 3) return OT_WIN;
        if (p->checks_given[p->stm ^ 1] >= 3) return OT_LOSS;
    }
    if (n_legal == 0) return opos_in_check(p) ? OT_LOSS : OT_DRAW;
    if (p->repetition < 0) return OT_DRAW; /* can_claim_3fold_repetition */
    if (p->variant != OV_CRAZYHOUSE && p->rule50 > 99) return OT_DRAW; /* n_legal > 0 here (board.cpp:150-159) */
    if (insufficient_material(p)) return OT_DRAW;
    return OT_NONE;
}

uint64_t opos_perft(const OPos* p, int depth) {
    uint32_t mv[OPOS_MAX_MOVES];
    const int n = opos_legal_moves(p, mv);
    if (depth <= 1) return depth == 1 ? (uint64_t)n : 1;
    uint64_t total = 0;
    OPos* q = (OPos*)malloc(sizeof(OPos));
    for (int i = 0; i < n; ++i) {
        memcpy(q, p, sizeof(OPos));
        opos_do_move(q, mv[i]);
        total += opos_perft(q, depth - 1);
    }
    free(q);
    return total;
}

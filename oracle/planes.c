/* oracle/planes.c -- CPU ORACLE (test infrastructure only).
 *
 * Restates engine/src/environments/chess_related/inputrepresentation.cpp of QueensGambit/CrazyAra:
 *   PlaneData helpers :33-109, set_plane_* :112-424, default_board_to_planes :426-501,
 *   board_to_planes_chess_v3 :536-566, board_to_planes_crazyhouse_v3/_v2 :569-595,
 *   board_to_planes_lichess_v3 :599-624, dispatcher board_to_planes :628-680;
 *   constants from boardstate.h:42-267.  The layout depends on the reference's compile-time MODE_* flag, which is
 *   a run-time `mode` argument here.
 * Pinned by the PlaneStatistics literals of engine/tests/tests.cpp (tests/test_oracle_planes.py).
 */
#include "planes.h"

#include <string.h>

typedef struct {
    const OPos* pos;
    float* base;
    float* cur;
    int flip;
    int normalize;
    int mode;
} PD;

static uint64_t flip_vertical(uint64_t x) { return __builtin_bswap64(x); } /* sfutil.cpp:178 */

static float max_prisoners(const PD* p) { return p->mode == OMODE_CRAZYHOUSE ? 32.0f : 16.0f; }
static float max_no_progress(const PD* p) { return p->mode == OMODE_CRAZYHOUSE ? 40.0f : 50.0f; }

static void inc(PD* p, int n) { p->cur += 64 * n; }
static void plane_value(PD* p, float v, int increment) {
    for (int i = 0; i < 64; ++i) p->cur[i] = v;
    if (increment) inc(p, 1);
}
static void plane_bitboard(PD* p, uint64_t bb) { /* set_bits_from_bitmap :33-46 */
    if (p->flip) bb = flip_vertical(bb);
    for (int i = 0; i < 64; ++i)
        if ((bb >> i) & 1) p->cur[i] = 1.0f;
    inc(p, 1);
}
static void single_square(PD* p, int sq, float v, int increment) {
    p->cur[p->flip ? (sq ^ 56) : sq] = v;
    if (increment) inc(p, 1);
}
static int me(const PD* p) { return p->pos->stm; }
static int you(const PD* p) { return p->pos->stm ^ 1; }

static void set_plane_pieces(PD* p) {
    const int colors[2] = {me(p), you(p)};
    for (int c = 0; c < 2; ++c)
        for (int pt = OP_PAWN; pt <= OP_KING; ++pt) plane_bitboard(p, opos_pieces_bb(p->pos, colors[c], pt));
}
static void set_plane_repetition(PD* p, int rep) {
    if (rep >= 1) {
        plane_value(p, 1.0f, 1);
        if (rep >= 2) {
            plane_value(p, 1.0f, 1);
            return;
        }
        inc(p, 1);
        return;
    }
    inc(p, 2);
}
static void set_plane_pockets(PD* p) {
    const int colors[2] = {me(p), you(p)};
    for (int c = 0; c < 2; ++c)
        for (int pt = OP_PAWN; pt <= OP_QUEEN; ++pt) {
            const int cnt = p->pos->variant == OV_CRAZYHOUSE ? p->pos->hand[colors[c]][pt] : 0;
            if (cnt > 0) plane_value(p, p->normalize ? cnt / max_prisoners(p) : (float)cnt, 0);
            inc(p, 1);
        }
}
static uint64_t promoted_bb(const OPos* pos) {
    uint64_t bb = 0;
    for (int s = 0; s < 64; ++s)
        if (pos->promoted[s] && pos->board[s]) bb |= 1ULL << s;
    return bb;
}
static void set_plane_promoted(PD* p) {
    const uint64_t pr = promoted_bb(p->pos);
    plane_bitboard(p, pr & opos_pieces_bb(p->pos, me(p), 0));
    plane_bitboard(p, pr & opos_pieces_bb(p->pos, you(p), 0));
}
static void set_plane_ep(PD* p) {
    if (p->pos->ep >= 0) single_square(p, p->pos->ep, 1.0f, 0);
    inc(p, 1);
}
static void set_plane_color(PD* p) {
    if (me(p) == 0) {
        plane_value(p, 1.0f, 1);
        return;
    }
    inc(p, 1);
}
static void set_plane_move_count(PD* p) {
    const float v = (float)((p->pos->game_ply / 2) + 1);
    plane_value(p, p->normalize ? v / 500.0f : v, 1);
}
static void set_plane_castling(PD* p) { /* :174-213: me-OO, me-OOO, you-OO, you-OOO */
    const int order[2] = {me(p), you(p)};
    for (int c = 0; c < 2; ++c)
        for (int side = 0; side < 2; ++side) {
            if (opos_can_castle(p->pos, order[c] * 2 + side)) plane_value(p, 1.0f, 0);
            inc(p, 1);
        }
}
static void set_no_progress(PD* p) {
    const float v = (float)p->pos->rule50;
    plane_value(p, p->normalize ? v / max_no_progress(p) : v, 1);
}
static void set_remaining_checks(PD* p) { /* :221-242 */
    if (p->pos->variant == OV_THREECHECK) {
        const int colors[2] = {me(p), you(p)};
        for (int c = 0; c < 2; ++c) {
            const int g = p->pos->checks_given[colors[c]];
            if (g != 0) {
                plane_value(p, 1.0f, 1);
                if (g >= 2) plane_value(p, 1.0f, 0);
                inc(p, 1);
            } else {
                inc(p, 2);
            }
        }
        return;
    }
    inc(p, 4);
}
static int variant_channel(int variant) { /* boardstate.h:269-279 CHANNEL_MAPPING_VARIANTS */
    switch (variant) {
        case OV_CHESS: return 1;
        case OV_CRAZYHOUSE: return 2;
        case OV_KOTH: return 3;
        case OV_THREECHECK: return 4;
        case OV_ANTI: return 5;
        case OV_ATOMIC: return 6;
        case OV_HORDE: return 7;
        case OV_RACE: return 8;
    }
    return 1;
}
static void set_variant_and_960(PD* p) { /* :246-260 */
    if (p->pos->chess960) plane_value(p, 1.0f, 0);
    float* pre = p->cur;
    inc(p, variant_channel(p->pos->variant));
    plane_value(p, 1.0f, 0);
    p->cur = pre;
    inc(p, 9);
}
static void set_last_moves(PD* p, int n_last_cfg) { /* :262-282 */
    float* pre = p->cur;
    const int n = p->pos->n_last < n_last_cfg ? p->pos->n_last : n_last_cfg;
    for (int i = 0; i < n; ++i) {
        const uint32_t m = p->pos->last_moves[i];
        if (OM_TYPE(m) == OM_DROP)
            inc(p, 1);
        else
            single_square(p, OM_FROM(m), 1.0f, 1);
        single_square(p, OM_TO(m), 1.0f, 1);
    }
    p->cur = pre;
    inc(p, 2 * n_last_cfg);
}
static void set_960(PD* p) {
    if (p->pos->chess960) plane_value(p, 1.0f, 0);
    inc(p, 1);
}
static void set_piece_masks(PD* p) {
    plane_bitboard(p, opos_pieces_bb(p->pos, me(p), 0));
    plane_bitboard(p, opos_pieces_bb(p->pos, you(p), 0));
}
static void set_checkerboard(PD* p) { /* :301-313 */
    int target = 1;
    for (int row = 0; row < 8; ++row) {
        for (int col = 0; col < 8; ++col) {
            if (col % 2 == target) *p->cur = 1.0f;
            ++p->cur;
        }
        target = !target;
    }
}
static void set_relative_count(PD* p, float rel) {
    if (rel != 0) plane_value(p, p->normalize ? rel / 8.0f : rel, 0);
    inc(p, 1);
}
static void set_material_diff(PD* p, int with_king) {
    for (int pt = OP_PAWN; pt <= (with_king ? OP_KING : OP_QUEEN); ++pt)
        set_relative_count(p, (float)(opos_count(p->pos, me(p), pt) - opos_count(p->pos, you(p), pt)));
}
static void set_material_count(PD* p, int with_king) {
    for (int pt = OP_PAWN; pt <= (with_king ? OP_KING : OP_QUEEN); ++pt)
        set_relative_count(p, (float)opos_count(p->pos, me(p), pt));
}
static void set_opposite_bishops(PD* p) { /* Position::opposite_bishops */
    const uint64_t wb = opos_pieces_bb(p->pos, 0, OP_BISHOP), bb = opos_pieces_bb(p->pos, 1, OP_BISHOP);
    if (__builtin_popcountll(wb) == 1 && __builtin_popcountll(bb) == 1) {
        const int ws = __builtin_ctzll(wb), bs = __builtin_ctzll(bb);
        const int wc = ((ws >> 3) + (ws & 7)) & 1, bc = ((bs >> 3) + (bs & 7)) & 1;
        if (wc != bc) plane_value(p, 1.0f, 0);
    }
    inc(p, 1);
}
static void set_checkers(PD* p) { plane_bitboard(p, opos_checkers_bb(p->pos)); }

static int channel(const PD* p) { return (int)((p->cur - p->base) / 64); }

static int default_planes(PD* p, int rep) { /* :426-501 */
    const int total = p->mode == OMODE_CRAZYHOUSE ? 34 : (p->mode == OMODE_CHESS ? 39 : 63);
    memset(p->base, 0, sizeof(float) * 64 * total);
    set_plane_pieces(p);
    set_plane_repetition(p, rep);
    if (p->mode != OMODE_CHESS) {
        set_plane_pockets(p);
        set_plane_promoted(p);
    }
    set_plane_ep(p);
    set_plane_color(p);
    set_plane_move_count(p);
    set_plane_castling(p);
    set_no_progress(p);
    if (p->mode == OMODE_LICHESS) {
        set_remaining_checks(p);
        set_variant_and_960(p);
    }
    if (p->mode == OMODE_CHESS) set_960(p);
    if (p->mode == OMODE_CHESS || p->mode == OMODE_LICHESS) set_last_moves(p, 8);
    return channel(p) == total ? total : -1;
}

static int chess_v3(PD* p, int rep) { /* :536-566 */
    set_plane_pieces(p);
    set_plane_repetition(p, rep);
    set_plane_ep(p);
    set_plane_castling(p);
    set_no_progress(p);
    set_last_moves(p, 8);
    set_960(p);
    set_piece_masks(p);
    set_checkerboard(p);
    set_material_diff(p, 0);
    set_opposite_bishops(p);
    set_checkers(p);
    set_material_count(p, 0);
    return channel(p);
}

int oplanes_channels(int mode, int version) {
    if (version <= 1) return mode == OMODE_CRAZYHOUSE ? 34 : (mode == OMODE_CHESS ? 39 : 63);
    if (mode == OMODE_CRAZYHOUSE) return version == 2 ? 51 : 64;
    if (mode == OMODE_CHESS) return version == 3 ? 52 : -1;
    return version == 3 ? 80 : 63; /* lichess: v2 falls through to the default layout (:669-676) */
}

int oplanes_encode(const OPos* pos, int mode, int version, int normalize, float* out) {
    PD pd;
    pd.pos = pos;
    pd.base = pd.cur = out;
    pd.flip = (pos->variant == OV_RACE) ? 0 : (pos->stm != 0); /* inputrepresentation.h:58-66 */
    pd.normalize = normalize;
    pd.mode = mode;
    const int rep = opos_number_repetitions(pos); /* BoardState::get_state_planes, boardstate.cpp:76-79 */
    const int total = oplanes_channels(mode, version);
    if (total < 0) return -1;
    if (version <= 1 || (mode == OMODE_LICHESS && version == 2)) return default_planes(&pd, rep);
    memset(out, 0, sizeof(float) * 64 * total);
    if (mode == OMODE_CHESS) return chess_v3(&pd, rep) == 52 ? 52 : -1;
    if (mode == OMODE_CRAZYHOUSE) {
        if (version == 3) { /* :569-577 */
            chess_v3(&pd, rep);
            set_plane_pockets(&pd);
            set_plane_promoted(&pd);
            return channel(&pd) == 64 ? 64 : -1;
        }
        /* v2 :579-595 */
        set_plane_pieces(&pd);
        set_plane_repetition(&pd, rep);
        set_plane_pockets(&pd);
        set_plane_promoted(&pd);
        set_plane_ep(&pd);
        set_plane_color(&pd);
        set_plane_move_count(&pd);
        set_plane_castling(&pd);
        set_no_progress(&pd);
        set_960(&pd);
        set_last_moves(&pd, 8);
        return channel(&pd) == 51 ? 51 : -1;
    }
    /* lichess v3 :599-624 */
    set_plane_pieces(&pd);
    set_plane_repetition(&pd, rep);
    set_plane_pockets(&pd);
    set_plane_promoted(&pd);
    set_plane_ep(&pd);
    inc(&pd, 2); /* colour info and total move count are skipped */
    set_plane_castling(&pd);
    set_no_progress(&pd);
    set_remaining_checks(&pd);
    set_variant_and_960(&pd);
    set_last_moves(&pd, 8);
    set_piece_masks(&pd);
    set_checkerboard(&pd);
    set_material_diff(&pd, 1);
    set_opposite_bishops(&pd);
    set_checkers(&pd);
    set_material_count(&pd, 1);
    return channel(&pd) == 80 ? 80 : -1;
}

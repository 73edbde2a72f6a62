// Board -> network input planes on the device (and, for tests, on the host through the same source).
//
// Semantics: engine/src/environments/chess_related/inputrepresentation.cpp (board_to_planes :628-680 and the
// per-version bodies :426-624) with the constants of boardstate.h:207-245.  The reference's compile-time MODE_*
// is the run-time `mode`.  Unlike the reference (one thread filling plane after plane) the encoder is organised
// per SQUARE: every lane owns squares {lane, lane+32} and emits their whole channel vector, which is contiguous in
// the NHWC layout the tcgen05 stem convolution reads.
#pragma once
#include "chess_dev.cuh"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace ara {

enum : int { MODE_CRAZYHOUSE = 0, MODE_CHESS = 1, MODE_LICHESS = 2 };

ARA_HD int planes_channels(int mode, int version) {
    if (version <= 1) return mode == MODE_CRAZYHOUSE ? 34 : (mode == MODE_CHESS ? 39 : 63);
    if (mode == MODE_CRAZYHOUSE) return version == 2 ? 51 : 64;
    if (mode == MODE_CHESS) return version == 3 ? 52 : -1;
    return version == 3 ? 80 : 63;
}

struct NchwF32Writer {  // the reference's [C,8,8] float layout (State::get_state_planes)
    float* out;
    ARA_HD void put(int c, int sq, float v) const { out[c * 64 + sq] = v; }
};
#if defined(__CUDACC__)
struct NhwcF16Writer {  // [64, cpad] fp16: the stem convolution's A operand
    __half* out;
    int cpad;
    ARA_HD void put(int c, int sq, float v) const { out[sq * cpad + c] = __float2half_rn(v); }
};
#endif

struct PlaneCtx {
    const Board* b;
    int mode, flip, me, you;
    bool normalize;
    uint64_t own, opp, checkers, promoted;
    int cnt[2][6];
    bool opp_bishops;
};

ARA_HD int variant_channel(int variant) { return variant + 1; }  // chess 1, crazyhouse 2, koth 3, 3check 4 (boardstate.h:269-279)

template <class W>
ARA_HD void encode_square(const PlaneCtx& p, int version, int sq, const W& w) {
    const Board& b = *p.b;
    const int src = p.flip ? (sq ^ 56) : sq;  // board square shown at output square sq
    const uint64_t sbit = bit(src);
    const float max_prisoners = p.mode == MODE_CRAZYHOUSE ? 32.0f : 16.0f;
    const float max_no_progress = p.mode == MODE_CRAZYHOUSE ? 40.0f : 50.0f;
    int c = 0;
#define EMIT(v) w.put(c++, sq, (v))
    auto pieces_planes = [&]() {
        for (int k = 0; k < 2; ++k) {
            const int col = k == 0 ? p.me : p.you;
            for (int pt = 0; pt < 6; ++pt) EMIT((b.by_type[pt] & b.by_color[col] & sbit) ? 1.0f : 0.0f);
        }
    };
    auto repetition_planes = [&]() {
        const int rep = b.repetition == 0 ? 0 : 1;  // Board::number_repetitions never returns 2 (board.cpp:132-141)
        EMIT(rep >= 1 ? 1.0f : 0.0f);
        EMIT(0.0f);
    };
    auto pockets_planes = [&]() {
        for (int k = 0; k < 2; ++k) {
            const int col = k == 0 ? p.me : p.you;
            for (int pt = 0; pt < 5; ++pt) {
                const int cnt = b.variant == V_CRAZYHOUSE ? b.hand[col][pt] : 0;
                EMIT(cnt > 0 ? (p.normalize ? cnt / max_prisoners : static_cast<float>(cnt)) : 0.0f);
            }
        }
    };
    auto promoted_planes = [&]() {
        EMIT((p.promoted & p.own & sbit) ? 1.0f : 0.0f);
        EMIT((p.promoted & p.opp & sbit) ? 1.0f : 0.0f);
    };
    auto ep_plane = [&]() { EMIT((b.ep != 0xFF && b.ep == src) ? 1.0f : 0.0f); };
    auto color_plane = [&]() { EMIT(p.me == 0 ? 1.0f : 0.0f); };
    auto move_count_plane = [&]() {
        const float v = static_cast<float>((b.game_ply / 2) + 1);
        EMIT(p.normalize ? v / 500.0f : v);
    };
    auto castling_planes = [&]() {
        for (int k = 0; k < 2; ++k) {
            const int col = k == 0 ? p.me : p.you;
            for (int side = 0; side < 2; ++side) EMIT(b.castle_rook[col * 2 + side] != 0xFF ? 1.0f : 0.0f);
        }
    };
    auto no_progress_plane = [&]() {
        const float v = static_cast<float>(b.rule50);
        EMIT(p.normalize ? v / max_no_progress : v);
    };
    auto remaining_checks_planes = [&]() {
        for (int k = 0; k < 2; ++k) {
            const int col = k == 0 ? p.me : p.you;
            const int g = b.variant == V_THREECHECK ? checks_given(b, col) : 0;
            EMIT(g != 0 ? 1.0f : 0.0f);
            EMIT(g >= 2 ? 1.0f : 0.0f);
        }
    };
    auto variant_960_planes = [&]() {
        const int vc = variant_channel(b.variant);
        for (int k = 0; k < 9; ++k) EMIT((k == 0 ? b.chess960 != 0 : k == vc) ? 1.0f : 0.0f);
    };
    auto last_moves_planes = [&]() {
        for (int i = 0; i < 8; ++i) {
            if (i < b.n_last) {
                const Move m = b.last_moves[i];
                EMIT((!mv_is_drop(m) && mv_from(m) == src) ? 1.0f : 0.0f);
                EMIT(mv_to(m) == src ? 1.0f : 0.0f);
            } else {
                EMIT(0.0f);
                EMIT(0.0f);
            }
        }
    };
    auto is960_plane = [&]() { EMIT(b.chess960 ? 1.0f : 0.0f); };
    auto masks_planes = [&]() {
        EMIT((p.own & sbit) ? 1.0f : 0.0f);
        EMIT((p.opp & sbit) ? 1.0f : 0.0f);
    };
    auto checkerboard_plane = [&]() { EMIT((((sq >> 3) + (sq & 7)) & 1) ? 1.0f : 0.0f); };  // not flipped (:301-313)
    auto rel_count = [&](float rel) { EMIT(rel != 0 ? (p.normalize ? rel / 8.0f : rel) : 0.0f); };
    auto material_diff_planes = [&](int npt) {
        for (int pt = 0; pt < npt; ++pt) rel_count(static_cast<float>(p.cnt[0][pt] - p.cnt[1][pt]));
    };
    auto material_count_planes = [&](int npt) {
        for (int pt = 0; pt < npt; ++pt) rel_count(static_cast<float>(p.cnt[0][pt]));
    };
    auto opp_bishops_plane = [&]() { EMIT(p.opp_bishops ? 1.0f : 0.0f); };
    auto checkers_plane = [&]() { EMIT((p.checkers & sbit) ? 1.0f : 0.0f); };
    auto chess_v3 = [&]() {  // :536-566
        pieces_planes(); repetition_planes(); ep_plane(); castling_planes(); no_progress_plane(); last_moves_planes();
        is960_plane(); masks_planes(); checkerboard_plane(); material_diff_planes(5); opp_bishops_plane();
        checkers_plane(); material_count_planes(5);
    };

    if (version <= 1 || (p.mode == MODE_LICHESS && version == 2)) {  // default_board_to_planes :426-501
        pieces_planes();
        repetition_planes();
        if (p.mode != MODE_CHESS) { pockets_planes(); promoted_planes(); }
        ep_plane(); color_plane(); move_count_plane(); castling_planes(); no_progress_plane();
        if (p.mode == MODE_LICHESS) { remaining_checks_planes(); variant_960_planes(); }
        if (p.mode == MODE_CHESS) is960_plane();
        if (p.mode != MODE_CRAZYHOUSE) last_moves_planes();
    } else if (p.mode == MODE_CHESS) {
        chess_v3();
    } else if (p.mode == MODE_CRAZYHOUSE) {
        if (version == 3) {  // :569-577
            chess_v3(); pockets_planes(); promoted_planes();
        } else {  // v2 :579-595
            pieces_planes(); repetition_planes(); pockets_planes(); promoted_planes(); ep_plane(); color_plane();
            move_count_plane(); castling_planes(); no_progress_plane(); is960_plane(); last_moves_planes();
        }
    } else {  // lichess v3 :599-624
        pieces_planes(); repetition_planes(); pockets_planes(); promoted_planes(); ep_plane();
        EMIT(0.0f); EMIT(0.0f);  // colour info and move count are skipped
        castling_planes(); no_progress_plane(); remaining_checks_planes(); variant_960_planes(); last_moves_planes();
        masks_planes(); checkerboard_plane(); material_diff_planes(6); opp_bishops_plane(); checkers_plane();
        material_count_planes(6);
    }
#undef EMIT
}

ARA_HD PlaneCtx make_plane_ctx(const Board& b, int mode, bool normalize) {
    PlaneCtx p;
    p.b = &b;
    p.mode = mode;
    p.flip = b.stm != 0;  // racing kings (no flip) is not a supported variant
    p.me = b.stm;
    p.you = b.stm ^ 1;
    p.normalize = normalize;
    p.own = b.by_color[p.me];
    p.opp = b.by_color[p.you];
    p.checkers = checkers_bb(b);
    p.promoted = b.promoted;
    for (int pt = 0; pt < 6; ++pt) {
        p.cnt[0][pt] = popc64(b.by_type[pt] & p.own);
        p.cnt[1][pt] = popc64(b.by_type[pt] & p.opp);
    }
    const uint64_t wb = pieces(b, 0, PT_BISHOP), bb = pieces(b, 1, PT_BISHOP);
    p.opp_bishops = false;
    if (popc64(wb) == 1 && popc64(bb) == 1) {
        const int ws = lsb64(wb), bs = lsb64(bb);
        p.opp_bishops = (((ws >> 3) + (ws & 7)) & 1) != (((bs >> 3) + (bs & 7)) & 1);
    }
    return p;
}

// all lanes of the warp call this; every lane encodes squares lane, lane+32 (all 64 on the host)
template <class W>
ARA_HD void encode_planes(const Board& b, int mode, int version, bool normalize, const W& w) {
    const PlaneCtx p = make_plane_ctx(b, mode, normalize);
    for (int sq = ARA_LANE; sq < 64; sq += ARA_WARP_N) encode_square(p, version, sq, w);
}

}  // namespace ara

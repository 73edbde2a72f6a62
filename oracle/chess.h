/* oracle/chess.h -- CPU ORACLE (test infrastructure only; never linked into the product).
 *
 * Rules engine standing in for the reference's (un-vendored) move generator:
 *   QueensGambit/Stockfish @ ae90e5ff332e4b21dfa499ef1667b1d8ab676c37 (multi-variant Stockfish fork), reached
 *   through engine/src/environments/chess_related/board.{h,cpp} and boardstate.{h,cpp}.
 * The dependency is absent from /root/reference, so its published algorithm is restated here in the simplest
 * possible form (mailbox board, pseudo-legal generation + make/unmake legality test) and pinned by
 *   - perft known answers (public chess-programming values) and
 *   - every rule/FEN assertion of engine/tests/tests.cpp that concerns chess, chess960, crazyhouse,
 *     King-of-the-Hill and Three-check (tests/test_oracle_chess.py).
 */
#ifndef ORACLE_CHESS_H
#define ORACLE_CHESS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { OV_CHESS = 0, OV_CRAZYHOUSE = 1, OV_KOTH = 2, OV_THREECHECK = 3, OV_ANTI = 4, OV_ATOMIC = 5, OV_HORDE = 6, OV_RACE = 7 };
enum { OP_NONE = 0, OP_PAWN = 1, OP_KNIGHT = 2, OP_BISHOP = 3, OP_ROOK = 4, OP_QUEEN = 5, OP_KING = 6 };
enum { OM_NORMAL = 0, OM_PROMOTION = 1, OM_ENPASSANT = 2, OM_CASTLING = 3, OM_DROP = 4 };
/* TerminalType of the reference (engine/src/state.h) */
enum { OT_LOSS = 0, OT_DRAW = 1, OT_WIN = 2, OT_CUSTOM = 3, OT_NONE = 4 };

#define OMOVE(from, to, type, pt) ((uint32_t)(from) | ((uint32_t)(to) << 6) | ((uint32_t)(type) << 12) | ((uint32_t)(pt) << 16))
#define OM_FROM(m) ((int)((m) & 63))
#define OM_TO(m) ((int)(((m) >> 6) & 63))
#define OM_TYPE(m) ((int)(((m) >> 12) & 7))
#define OM_PT(m) ((int)(((m) >> 16) & 7))

#define OPOS_MAX_HIST 1024
#define OPOS_MAX_MOVES 512

typedef struct OPos {
    int8_t board[64];      /* 0 empty; white 1..6; black 9..14 */
    uint8_t promoted[64];  /* crazyhouse: piece on this square is a promoted pawn */
    int hand[2][7];        /* pocket counts [color][piece type] */
    int stm;               /* 0 white, 1 black */
    int castle_rook[4];    /* rook origin square for W-OO, W-OOO, B-OO, B-OOO, or -1 */
    int ep;                /* en-passant square or -1 */
    int rule50;
    int game_ply;
    int checks_given[2];
    int variant;
    int chess960;
    int plies_from_null;
    int repetition;        /* Stockfish StateInfo::repetition semantics */
    uint64_t key;
    /* history since set(): keys/repetition of all earlier positions, most recent last */
    int hist_len;
    uint64_t hist_key[OPOS_MAX_HIST];
    int16_t hist_rep[OPOS_MAX_HIST];
    /* last moves, most recent first (Board::lastMoves, board.cpp:216-225) */
    int n_last;
    uint32_t last_moves[8];
} OPos;

void opos_init_tables(void);
int opos_set(OPos* p, const char* fen, int variant, int is960);
void opos_copy(OPos* dst, const OPos* src);
void opos_fen(const OPos* p, char* buf);
int opos_legal_moves(const OPos* p, uint32_t* out);
void opos_do_move(OPos* p, uint32_t m);
void opos_move_to_uci(const OPos* p, uint32_t m, char* buf);
uint32_t opos_uci_to_move(const OPos* p, const char* uci);
int opos_in_check(const OPos* p);
uint64_t opos_checkers_bb(const OPos* p);
int opos_gives_check(const OPos* p, uint32_t m);
int opos_is_terminal(const OPos* p, int n_legal);
int opos_number_repetitions(const OPos* p);
uint64_t opos_compute_key(const OPos* p);
uint64_t opos_perft(const OPos* p, int depth);
uint64_t opos_pieces_bb(const OPos* p, int color, int pt); /* pt 0 = all */
int opos_count(const OPos* p, int color, int pt);
int opos_can_castle(const OPos* p, int right);
const char* opos_start_fen(int variant);
uint64_t opos_zobrist(int idx);
size_t opos_sizeof(void);

#ifdef __cplusplus
}
#endif
#endif

/* oracle/policy.c -- CPU ORACLE (test infrastructure only).
 *
 * Restates the policy indexing of QueensGambit/CrazyAra:
 *   label generator  engine/src/environments/chess_related/outputrepresentation.cpp:66-181
 *   mirror_move      engine/src/environments/chess_related/sfutil.cpp:183-197
 *   move -> label    outputrepresentation.cpp:39-56 + sfutil.cpp:199-285 (defined on UCI strings, so it is
 *                    independent of the move encoding; 960 castling = king-from + rook-square)
 *   policy-map index DeepCrazyhouse/src/domain/variants/plane_policy_representation.py:22-213, whose output is the
 *                    frozen table FLAT_PLANE_IDX (policymaprepresentation.h) -- regenerated here, compared with the
 *                    table in tests/test_oracle_policy.py.
 */
#include "policy.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_LABELS 2400
typedef struct {
    int n;
    char s[MAX_LABELS][8];
    int flat[MAX_LABELS];
    int built;
} LabelSet;
static LabelSet g_sets[3];

static void add(LabelSet* L, const char* s) { strcpy(L->s[L->n++], s); }

static int flat_index_of(int mode, const char* l) { /* plane_policy_representation.py:167-213 */
    const int n = (int)strlen(l);
    if (l[1] == '@') {
        const char* order = "PNBRQ";
        const int piece = (int)(strchr(order, l[0]) - order);
        /* the frozen lichess table keeps the drop planes at 76..80 (overlapping the king-promotion planes):
           policymaprepresentation.h:2314 is authoritative, not the python docstring */
        const int base = 76;
        (void)mode;
        const int to = (l[3] - '1') * 8 + (l[2] - 'a');
        return (base + piece) * 64 + to;
    }
    const int ff = l[0] - 'a', fr = l[1] - '1', tf = l[2] - 'a', tr = l[3] - '1';
    const int from = fr * 8 + ff;
    const int dy = tr - fr, dx = tf - ff;
    if (n == 5) {
        int piece_id = 0; /* 1 knight, 2 bishop, 3 rook, 4 queen, 5 king */
        switch (l[4]) {
            case 'n': piece_id = 1; break;
            case 'b': piece_id = 2; break;
            case 'r': piece_id = 3; break;
            case 'q': piece_id = 4; break;
            case 'k': piece_id = 5; break;
        }
        return (64 + (piece_id - 1) * 3 + (dx + 1)) * 64 + from;
    }
    const int ax = abs(dx), ay = abs(dy);
    if ((ax < ay ? ax : ay) == 1 && (ax > ay ? ax : ay) == 2) {
        static const int cases[8][2] = {{2, 1}, {1, 2}, {-1, 2}, {-2, 1}, {-2, -1}, {-1, -2}, {1, -2}, {2, -1}};
        for (int i = 0; i < 8; ++i)
            if (cases[i][0] == dy && cases[i][1] == dx) return (56 + i) * 64 + from;
    }
    const int len = (ax > ay ? ax : ay) - 1;
    int dir = -1;
    if (dx == 0 && dy > 0) dir = 0;
    else if (dx > 0 && dy > 0) dir = 1;
    else if (dx > 0 && dy == 0) dir = 2;
    else if (dy < 0 && dx > 0) dir = 3;
    else if (dx == 0 && dy < 0) dir = 4;
    else if (dx < 0 && dy < 0) dir = 5;
    else if (dx < 0 && dy == 0) dir = 6;
    else if (dx < 0 && dy > 0) dir = 7;
    return (dir * 7 + len) * 64 + from;
}

static void build(int mode) {
    LabelSet* L = &g_sets[mode];
    if (L->built) return;
    L->n = 0;
    char buf[8];
    /* generate_uci_labels (outputrepresentation.cpp:108-145) */
    for (int f = 0; f < 8; ++f)
        for (int r = 0; r < 8; ++r) {
            int dest[64][2], nd = 0;
            for (int i = 0; i < 8; ++i) dest[nd][0] = i, dest[nd++][1] = r;
            for (int i = 0; i < 8; ++i) dest[nd][0] = f, dest[nd++][1] = i;
            for (int i = -7; i < 8; ++i) dest[nd][0] = f + i, dest[nd++][1] = r + i;
            for (int i = -7; i < 8; ++i) dest[nd][0] = f + i, dest[nd++][1] = r - i;
            static const int kf[8] = {-2, -1, -2, 1, 2, -1, 2, 1}, kr[8] = {-1, -2, 1, -2, -1, 2, 1, 2};
            for (int i = 0; i < 8; ++i) dest[nd][0] = f + kf[i], dest[nd++][1] = r + kr[i];
            for (int i = 0; i < nd; ++i) {
                const int f2 = dest[i][0], r2 = dest[i][1];
                if ((f != f2 || r != r2) && f2 >= 0 && f2 < 8 && r2 >= 0 && r2 < 8) {
                    sprintf(buf, "%c%c%c%c", 'a' + f, '1' + r, 'a' + f2, '1' + r2);
                    add(L, buf);
                }
            }
        }
    const char* promo = mode == OMODE_LICHESS ? "qrbnk" : "qrbn";
    for (int f = 0; f < 8; ++f)
        for (const char* p = promo; *p; ++p) {
            sprintf(buf, "%c2%c1%c", 'a' + f, 'a' + f, *p); add(L, buf);
            sprintf(buf, "%c7%c8%c", 'a' + f, 'a' + f, *p); add(L, buf);
            if (f > 0) {
                sprintf(buf, "%c2%c1%c", 'a' + f, 'a' + f - 1, *p); add(L, buf);
                sprintf(buf, "%c7%c8%c", 'a' + f, 'a' + f - 1, *p); add(L, buf);
            }
            if (f < 7) {
                sprintf(buf, "%c2%c1%c", 'a' + f, 'a' + f + 1, *p); add(L, buf);
                sprintf(buf, "%c7%c8%c", 'a' + f, 'a' + f + 1, *p); add(L, buf);
            }
        }
    if (mode != OMODE_CHESS) { /* generate_dropping_moves :147-163 */
        for (int f = 0; f < 8; ++f)
            for (int r = 0; r < 8; ++r)
                for (const char* p = "PNBRQ"; *p; ++p) {
                    if (*p == 'P' && (r == 0 || r == 7)) continue;
                    sprintf(buf, "%c@%c%c", *p, 'a' + f, '1' + r);
                    add(L, buf);
                }
    }
    for (int i = 0; i < L->n; ++i) L->flat[i] = flat_index_of(mode, L->s[i]);
    L->built = 1;
}

int opolicy_nb_labels(int mode) { build(mode); return g_sets[mode].n; }
int opolicy_nb_policy_channels(int mode) { return mode == OMODE_CRAZYHOUSE ? 81 : (mode == OMODE_CHESS ? 76 : 84); }
const char* opolicy_label(int mode, int idx) { build(mode); return g_sets[mode].s[idx]; }
int opolicy_flat_plane_idx(int mode, int idx) { build(mode); return g_sets[mode].flat[idx]; }
/* string -> small integer key: drops 0..511 (piece*64+sq), other moves 512 + from*64*8 + to*8 + promo */
static int label_key(const char* l) {
    if (l[1] == '@') {
        const char* order = "PNBRQ";
        const char* q = strchr(order, l[0]);
        if (!q) return -1;
        return (int)(q - order) * 64 + (l[3] - '1') * 8 + (l[2] - 'a');
    }
    int promo = 0;
    if (l[4]) {
        const char* order = "nbrqk";
        const char* q = strchr(order, l[4]);
        if (!q) return -1;
        promo = (int)(q - order) + 1;
    }
    const int from = (l[1] - '1') * 8 + (l[0] - 'a'), to = (l[3] - '1') * 8 + (l[2] - 'a');
    if (from < 0 || from > 63 || to < 0 || to > 63) return -1;
    return 512 + (from * 64 + to) * 8 + promo;
}
static int16_t g_lookup[3][512 + 64 * 64 * 8];
static int g_lookup_built[3];
int opolicy_label_index(int mode, const char* uci) {
    build(mode);
    if (!g_lookup_built[mode]) {
        const LabelSet* L = &g_sets[mode];
        for (size_t i = 0; i < sizeof(g_lookup[mode]) / sizeof(int16_t); ++i) g_lookup[mode][i] = -1;
        for (int i = 0; i < L->n; ++i) g_lookup[mode][label_key(L->s[i])] = (int16_t)i;
        g_lookup_built[mode] = 1;
    }
    const int k = label_key(uci);
    return k < 0 ? -1 : g_lookup[mode][k];
}
void opolicy_mirror(const char* uci, char* out) {
    strcpy(out, uci);
    for (int i = 0; uci[i]; ++i)
        if (uci[i] >= '1' && uci[i] <= '8') out[i] = (char)('0' + (8 - (uci[i] - '0') + 1));
}
int opolicy_move_index(const OPos* pos, uint32_t move, int mode, int is_policy_map) {
    char uci[8], mir[8];
    opos_move_to_uci(pos, move, uci);
    const int mirror = (pos->variant == OV_RACE) ? 0 : (pos->stm != 0); /* BoardState::mirror_policy */
    const char* key = uci;
    if (mirror) {
        opolicy_mirror(uci, mir);
        key = mir;
    }
    const int li = opolicy_label_index(mode, key);
    if (li < 0) return -1;
    return is_policy_map ? g_sets[mode].flat[li] : li;
}

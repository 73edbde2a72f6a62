/* oracle/mcts.c -- CPU ORACLE (test infrastructure only).
 *
 * Single-threaded, line-by-line restatement of the reference's search (QueensGambit/CrazyAra engine/src):
 *   node.h:97-246 (Node, revert_virtual_loss_and_update), node.h:819-843 (backup_value),
 *   node.cpp:82-106 (Node ctor), :365-453 (solve_for_terminal), :464-470 (sort), :507-529 (virtual loss),
 *   :571-593 (increment_no_visit_idx / fully_expand_node), :634-644 (prepare_node_for_visits), :655-679
 *   (collisions), :716-720 (set_value), :880-904 (check_for_terminal), :950-979 (dirichlet, priors), :1006-1010,
 *   :1056-1063 (u values), :1070-1109 (get_mcts_policy), :1123-1167 (best action, select_child_node), :1243-1246;
 *   nodedata.cpp:30-75; searchthread.cpp:164-271 (get_new_child_to_evaluate), :290-331, :347-380
 *   (create_mini_batch), :403-449; agents/mctsagent.cpp:166-196, :292-337; evalinfo.cpp:112-121, :195-249;
 *   util/blazeutil.h:78-88 (temperature), :113-124 (dirichlet), :155-180 (first_and_second_max).
 * It keeps the reference's cost structure on purpose (heap node per position, per-node heap vectors, per-leaf
 * "clone root state + replay the action path + do_action"), because it doubles as the CPU baseline of bench.py.
 * blaze vectors become plain arrays; blaze's argmax = first maximum.  Parity unpinned for visit counts / Q / policy
 * posterior: the reference's own tests pin none of them -- this file's dumps are the golden values.
 */
#include "mcts.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q_INIT (-1.0f)
#define NO_CHECKMATE 65535
#define LOSS_VALUE (-1)
#define DRAW_VALUE 0
#define WIN_VALUE 1

typedef struct ONode {
    /* Node */
    float* policy;     /* policyProbSmall */
    uint32_t* actions; /* legalActions */
    int* pidx;         /* policy-vector index of each action (MV_LOOKUP[_MIRRORED] applied at expansion) */
    int n_actions;
    uint64_t key;
    double value_sum;
    uint32_t real_visits;
    int plies_from_null;
    int number_parents;
    int is_terminal, has_nn, sorted;
    /* NodeData d (has_d == (d != nullptr)) */
    int has_d;
    uint32_t* n;
    float* q;
    struct ONode** child;
    uint8_t* vl;
    uint8_t* types;
    uint32_t free_visits, visit_sum;
    int checkmate_idx, end_in_ply, no_visit_idx, n_unsolved;
    int node_type;
    int inspected; /* NodeData::inspected: select_enhanced_move has looked at this node's unopened checks */
} ONode;

typedef struct {
    ONode* node;
    int child_idx;
} Step;
typedef struct {
    Step* steps;
    int len, cap;
} Traj;

enum { NB_NEW = 0, NB_COLLISION, NB_TERMINAL, NB_TRANSPOSITION };

typedef struct {
    uint32_t x; /* minstd_rand0 state = std::default_random_engine */
} MinStd;

/* glibc's rand() (stdlib/random_r.c, TYPE_3: x^31 + x^3 + 1 additive feedback over 31 words, seeded by the 16807
 * Lehmer generator, first 310 outputs discarded) -- what the reference's epsilon-greedy exploration draws from
 * (`rand() % counter`, searchthread.cpp:171-185, :124-162, :497-501).  Pinned to the live libc by tests. */
typedef struct {
    int32_t r[34];
    int f, b; /* front / rear index */
} GlibcRand;
static void glibc_srand(GlibcRand* g, unsigned seed) {
    if (seed == 0) seed = 1;
    g->r[0] = (int32_t)seed;
    int32_t word = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
        const long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        g->r[i] = word;
    }
    g->f = 3;
    g->b = 0;
    for (int i = 0; i < 310; ++i) {
        g->r[g->f] = (int32_t)((uint32_t)g->r[g->f] + (uint32_t)g->r[g->b]);
        g->f = (g->f + 1) % 31;
        g->b = (g->b + 1) % 31;
    }
}
static int glibc_rand(GlibcRand* g) {
    const uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
    g->r[g->f] = (int32_t)v;
    g->f = (g->f + 1) % 31;
    g->b = (g->b + 1) % 31;
    return (int)(v >> 1);
}
void oglibc_rand_sequence(unsigned seed, int n, int* out) { /* test hook */
    GlibcRand g;
    glibc_srand(&g, seed);
    for (int i = 0; i < n; ++i) out[i] = glibc_rand(&g);
}

struct OSearch {
    OSettings st;
    int channels, n_labels;
    OPos root_state;
    ONode* root;
    ONode* next_root; /* ownNextRoot / opponentsNextRoot of MCTSAgent (mctsagent.h): candidate root after applied moves */
    int next_root_valid;
    /* batch state (SearchThread members) */
    ONode** new_nodes;
    int* new_stm;
    int n_new;
    Traj* new_traj;
    Traj* coll_traj;
    int n_coll;
    Traj cur;
    uint32_t* actions_buf;
    int n_actions_buf, actions_cap;
    float* planes;
    /* Threads = 2: the batch state above belongs to the logical search thread `active`; the other thread's is parked
       here and swapped in by use_thread (every SearchThread of the reference owns these members, searchthread.h:48-75) */
    struct OParked {
        ONode** new_nodes;
        int* new_stm;
        int n_new;
        Traj* new_traj;
        Traj* coll_traj;
        int n_coll;
        Traj cur;
        uint32_t* actions_buf;
        int n_actions_buf, actions_cap;
        float* planes;
    } parked;
    int active;
    MinStd rng; /* the Dirichlet generator */
    GlibcRand crand; /* rand() of the epsilon-greedy exploration: seeded once (srand(seed)), advances across searches */
    unsigned long long num_nodes, sum_select_k, sum_depth;
    ONode** all_nodes;
    size_t n_all, cap_all;
};

/* ------------------------------------------------------------------ settings */
void osettings_default(OSettings* s, int mode) { /* uci/optionsuci.cpp:66-220 (non-RL build) */
    memset(s, 0, sizeof(*s));
    s->batch_size = mode == OMODE_CHESS ? 64 : 16;
    s->dirichlet_epsilon = 0.0f;
    s->dirichlet_alpha = 0.2f;
    s->node_policy_temperature = 1.7f;
    s->q_value_weight = 1.0f;
    s->q_veto_delta = 0.4f;
    s->cpuct_init = 2.5f;
    s->cpuct_base = 19652.0f;
    s->mcts_solver = 1;
    s->virtual_style = OVS_VIRTUAL_MIX;
    s->virtual_mix_threshold = 1000;
    s->simulations = 0;
    s->nodes = 0;
    s->seed = 42;
    s->threads = 1; /* the deterministic parity setting; the reference's UCI default is 2 (optionsuci.cpp:182) */
    s->mode = mode;
    s->input_version = mode == OMODE_CHESS ? 3 : 1;
}

/* ------------------------------------------------------------------ Dirichlet noise: libstdc++ restatement */
static uint32_t minstd_next(MinStd* g) {
    g->x = (uint32_t)(((uint64_t)g->x * 16807ULL) % 2147483647ULL);
    return g->x;
}
static float canonical_f(MinStd* g) { /* std::generate_canonical<float, 24>(minstd_rand0): one draw */
    const float sum = (float)(minstd_next(g) - 1u);
    float ret = sum / 2147483648.0f; /* float(2147483646.0L) */
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret;
}
static float gamma_f(MinStd* g, float alpha) { /* std::gamma_distribution<float>(alpha, 1.0f), fresh object */
    const float malpha = alpha < 1.0f ? alpha + 1.0f : alpha;
    const float a1 = malpha - 1.0f / 3.0f;
    const float a2 = 1.0f / sqrtf(9.0f * a1);
    int saved_ok = 0;
    float saved = 0.0f, u, v, n;
    do {
        do {
            if (saved_ok) {
                saved_ok = 0;
                n = saved;
            } else {
                float x, y, r2;
                do {
                    x = (float)(2.0f * canonical_f(g) - 1.0);
                    y = (float)(2.0f * canonical_f(g) - 1.0);
                    r2 = x * x + y * y;
                } while (r2 > 1.0 || r2 == 0.0);
                const float mult = sqrtf(-2 * logf(r2) / r2);
                saved = x * mult;
                saved_ok = 1;
                n = y * mult;
            }
            v = 1.0f + a2 * n;
        } while (v <= 0.0);
        v = v * v * v;
        u = canonical_f(g);
    } while (u > 1.0f - 0.0331 * n * n * n * n && (logf(u) > (0.5 * n * n + a1 * (1.0 - v + logf(v)))));
    if (alpha == malpha) return a1 * v * 1.0f;
    do u = canonical_f(g);
    while (u == 0.0);
    return powf(u, 1.0f / alpha) * a1 * v * 1.0f;
}
static void dirichlet_noise_g(MinStd* g, int n, float alpha, float* out) { /* get_dirichlet_noise blazeutil.h:113-124 */
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) {
        out[i] = gamma_f(g, alpha);
        sum += out[i];
    }
    for (int i = 0; i < n; ++i) out[i] /= sum;
}
static uint32_t minstd_seed(unsigned long long seed) { /* std::default_random_engine(seed) */
    const uint32_t x = (uint32_t)(seed % 2147483647ULL);
    return x == 0 ? 1u : x;
}
void odirichlet_noise(unsigned long long seed, int n, float alpha, float* out) { /* a fresh generator */
    MinStd g;
    g.x = minstd_seed(seed);
    dirichlet_noise_g(&g, n, alpha, out);
}

/* ------------------------------------------------------------------ helpers */
static void traj_push(Traj* t, ONode* n, int ci) {
    if (t->len == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 64;
        t->steps = (Step*)realloc(t->steps, sizeof(Step) * (size_t)t->cap);
    }
    t->steps[t->len].node = n;
    t->steps[t->len].child_idx = ci;
    t->len++;
}
static void traj_copy(Traj* dst, const Traj* src) {
    dst->len = 0;
    for (int i = 0; i < src->len; ++i) traj_push(dst, src->steps[i].node, src->steps[i].child_idx);
}

static int get_virtual_style(const OSettings* s, uint32_t visits) { /* node.h:87-95 */
    if (s->virtual_style == OVS_VIRTUAL_MIX) return visits > s->virtual_mix_threshold ? OVS_VIRTUAL_LOSS : OVS_VIRTUAL_VISIT;
    return s->virtual_style;
}

static void node_set_value(ONode* n, float value) { /* node.cpp:716-720 */
    ++n->real_visits;
    n->value_sum = (double)(value * (float)n->real_visits); /* float * uint32 -> float, stored in a double */
}
static float node_get_value(const ONode* n) { return (float)(n->value_sum / n->real_visits); }

static void alloc_node_data(ONode* n, int m) {
    n->has_d = 1;
    n->n = (uint32_t*)calloc((size_t)(m > 0 ? m : 1), sizeof(uint32_t));
    n->q = (float*)calloc((size_t)(m > 0 ? m : 1), sizeof(float));
    n->child = (ONode**)calloc((size_t)(m > 0 ? m : 1), sizeof(ONode*));
    n->vl = (uint8_t*)calloc((size_t)(m > 0 ? m : 1), 1);
    n->types = (uint8_t*)calloc((size_t)(m > 0 ? m : 1), 1);
    n->free_visits = 0;
    n->visit_sum = 0;
    n->checkmate_idx = NO_CHECKMATE;
    n->end_in_ply = 0;
    n->no_visit_idx = 1;
    n->node_type = ONT_UNSOLVED;
    n->n_unsolved = m;
}
static void add_empty_node(ONode* n, int idx) { /* nodedata.cpp:30-37 */
    n->n[idx] = 0;
    n->q[idx] = Q_INIT;
    n->vl[idx] = 0;
    n->types[idx] = ONT_UNSOLVED;
    n->child[idx] = NULL;
}

static ONode* node_new(OSearch* s, const OPos* state) { /* Node::Node node.cpp:82-106 */
    ONode* n = (ONode*)calloc(1, sizeof(ONode));
    uint32_t mv[OPOS_MAX_MOVES];
    n->n_actions = opos_legal_moves(state, mv);
    n->actions = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n->n_actions > 0 ? n->n_actions : 1));
    memcpy(n->actions, mv, sizeof(uint32_t) * (size_t)n->n_actions);
    n->pidx = (int*)malloc(sizeof(int) * (size_t)(n->n_actions > 0 ? n->n_actions : 1));
    for (int i = 0; i < n->n_actions; ++i) n->pidx[i] = opolicy_move_index(state, mv[i], s->st.mode, 1);
    n->node_type = ONT_UNSOLVED;
    n->key = state->key;
    n->plies_from_null = state->game_ply; /* steps_from_null() = game_ply (boardstate.cpp:81-84) */
    n->number_parents = 1;
    /* check_for_terminal node.cpp:880-904 */
    const int tt = opos_is_terminal(state, n->n_actions);
    if (tt != OT_NONE) {
        n->is_terminal = 1; /* mark_as_terminal: NodeData(), sorted, noVisitIdx = 0 */
        alloc_node_data(n, 0);
        n->sorted = 1;
        n->no_visit_idx = 0;
        if (tt == OT_WIN) {
            node_set_value(n, WIN_VALUE);
            n->node_type = ONT_WIN;
        } else if (tt == OT_DRAW) {
            node_set_value(n, DRAW_VALUE);
            n->node_type = ONT_DRAW;
            n->n_actions = 0; /* legalActions.clear() */
        } else if (tt == OT_LOSS) {
            node_set_value(n, LOSS_VALUE);
            n->node_type = ONT_LOSS;
        }
    }
    n->policy = (float*)calloc((size_t)(n->n_actions > 0 ? n->n_actions : 1), sizeof(float));
    if (s->n_all == s->cap_all) {
        s->cap_all = s->cap_all ? s->cap_all * 2 : 1024;
        s->all_nodes = (ONode**)realloc(s->all_nodes, sizeof(ONode*) * s->cap_all);
    }
    s->all_nodes[s->n_all++] = n;
    s->num_nodes++;
    return n;
}
static void node_free(ONode* n) {
    free(n->policy);
    free(n->actions);
    free(n->pidx);
    if (n->has_d) {
        free(n->n);
        free(n->q);
        free(n->child);
        free(n->vl);
        free(n->types);
    }
    free(n);
}

/* sort_moves_by_probabilities node.cpp:464-470.  std::sort on a permutation is unstable for equal priors, i.e. the
 * reference leaves the order of tied moves unspecified; the oracle (and the product) break ties by ascending
 * policy-vector index, which does not depend on the move generator's emission order. */
typedef struct {
    float p;
    uint32_t a;
    int i;
} SortItem;
static int cmp_desc(const void* x, const void* y) {
    const SortItem* a = (const SortItem*)x;
    const SortItem* b = (const SortItem*)y;
    if (a->p > b->p) return -1;
    if (a->p < b->p) return 1;
    return a->i - b->i; /* i = policy-vector index */
}
static void prepare_node_for_visits(ONode* n) { /* node.cpp:634-644 */
    SortItem* it = (SortItem*)malloc(sizeof(SortItem) * (size_t)(n->n_actions > 0 ? n->n_actions : 1));
    for (int i = 0; i < n->n_actions; ++i) it[i].p = n->policy[i], it[i].a = n->actions[i], it[i].i = n->pidx[i];
    qsort(it, (size_t)n->n_actions, sizeof(SortItem), cmp_desc);
    for (int i = 0; i < n->n_actions; ++i) n->policy[i] = it[i].p, n->actions[i] = it[i].a;
    free(it);
    n->sorted = 1;
    if (!n->has_d) {
        alloc_node_data(n, n->n_actions);
        add_empty_node(n, 0); /* reserve_initial_space */
    }
}
static void increment_no_visit_idx(ONode* n) { /* node.cpp:571-580 */
    if (n->no_visit_idx < n->n_actions) {
        add_empty_node(n, n->no_visit_idx);
        ++n->no_visit_idx;
    }
}
static void fully_expand_node(ONode* n) { /* node.cpp:582-593 */
    if (n->no_visit_idx != n->n_actions) {
        for (int i = n->no_visit_idx; i < n->n_actions; ++i) add_empty_node(n, i);
        n->no_visit_idx = n->n_actions;
        n->sorted = 1;
    }
}

static float get_current_cput(float visits, const OSettings* s) { /* node.cpp:1243-1246 (all float) */
    return logf((visits + s->cpuct_base + 1) / s->cpuct_base) + s->cpuct_init;
}

static int select_child_node(OSearch* s, ONode* n) { /* node.cpp:1150-1167 */
    if (!n->sorted) prepare_node_for_visits(n);
    if (n->no_visit_idx == 1) return 0;
    if (n->checkmate_idx != NO_CHECKMATE) return n->checkmate_idx;
    /* get_current_u_values node.cpp:1056-1063:  cput * subvector(P, 0, k) * (sqrt(N) / (n + 1.0))  with cput float,
       P float, sqrt(N) double (std::sqrt of an integer), n uint32.  Read left to right this is (cput*P) in float times a
       double vector; blaze, however, restructures (vector*scalar)*vector into (vector*vector)*scalar (the restructuring
       operators of blaze/math/expressions/DVecScalarMultExpr.h), so element i is evaluated as
           float( (double(P_i) * (sqrt(N) / (double(n_i) + 1.0))) * double(cput) ).
       blaze's source is absent here (empty submodule): this is its documented behaviour, not a measured one. */
    const float cput = get_current_cput((float)n->visit_sum, &s->st);
    const double sq = sqrt((double)n->visit_sum);
    int best = 0;
    float best_v = 0;
    for (int i = 0; i < n->no_visit_idx; ++i) {
        const float u = (float)(((double)n->policy[i] * (sq / ((double)n->n[i] + 1.0))) * (double)cput);
        const float v = n->q[i] + u;
        if (i == 0 || v > best_v) best = i, best_v = v;
    }
    s->sum_select_k += (unsigned long long)n->no_visit_idx;
    return best;
}

static void apply_virtual_loss_to_child(ONode* n, int ci, const OSettings* st) { /* node.cpp:507-529 */
    if (get_virtual_style(st, n->n[ci]) == OVS_VIRTUAL_LOSS)
        n->q[ci] = (float)(((double)n->q[ci] * n->n[ci] - 1) / (double)(n->n[ci] + 1));
    ++n->n[ci];
    ++n->visit_sum;
    ++n->vl[ci];
}

static void disable_action(ONode* n, int ci) { /* node.cpp:1006-1010 */
    n->policy[ci] = 0;
    n->q[ci] = (float)(-INT_MAX);
}

static int at_least_one_drawn_child(const ONode* n) { /* node.cpp:135-149; iterates d->childNodes (opened slots) */
    int drawn = 0;
    for (int i = 0; i < n->no_visit_idx; ++i) {
        const ONode* c = n->child[i];
        if (c == NULL || !c->has_d || (c->node_type != ONT_DRAW && c->node_type != ONT_WIN)) return 0;
        if (c->node_type == ONT_DRAW) drawn = 1;
    }
    return drawn;
}
static int only_won_child_nodes(const ONode* n) { /* node.h only_child_nodes_of_one_kind<WIN> */
    for (int i = 0; i < n->no_visit_idx; ++i)
        if (n->child[i] == NULL || n->child[i]->node_type != ONT_WIN) return 0;
    return 1;
}
static void define_end_ply(ONode* n, const ONode* child) { /* node.cpp:265-289 */
    if (n->node_type == ONT_LOSS) {
        for (int i = 0; i < n->no_visit_idx; ++i)
            if (n->child[i] && n->child[i]->end_in_ply + 1 > n->end_in_ply) n->end_in_ply = n->child[i]->end_in_ply + 1;
        return;
    }
    if (n->node_type == ONT_DRAW) {
        for (int i = 0; i < n->no_visit_idx; ++i)
            if (n->child[i] && n->child[i]->node_type == ONT_DRAW && n->child[i]->end_in_ply + 1 < n->end_in_ply)
                n->end_in_ply = n->child[i]->end_in_ply + 1;
        return;
    }
    n->end_in_ply = child->end_in_ply + 1;
}
static void update_solved_terminal(ONode* n, const ONode* child, int ci, int target) { /* node.cpp:291-297 */
    define_end_ply(n, child);
    node_set_value(n, (float)target);
    n->q[ci] = (float)target;
}
static int solve_for_terminal(ONode* n, int ci) { /* node.cpp:365-453, MODE_TWO_PLAYER, no tablebases */
    const ONode* c = n->child[ci];
    if (!c->has_d) return 0;
    if (c->node_type == ONT_UNSOLVED) return 0;
    if (n->node_type == ONT_WIN || n->node_type == ONT_LOSS || n->node_type == ONT_DRAW) return 0;
    if (n->types[ci] == ONT_UNSOLVED) {
        --n->n_unsolved;
        n->types[ci] = (uint8_t)c->node_type;
        if (c->node_type == ONT_WIN) disable_action(n, ci);
    }
    if (c->node_type == ONT_LOSS) { /* solved_win */
        n->node_type = ONT_WIN;
        update_solved_terminal(n, c, ci, WIN_VALUE);
        n->checkmate_idx = ci;
        return 1;
    }
    if (n->n_unsolved == 0 && c->node_type == ONT_WIN && only_won_child_nodes(n)) { /* solved_loss */
        n->node_type = ONT_LOSS;
        update_solved_terminal(n, c, ci, LOSS_VALUE);
        return 1;
    }
    if (n->n_unsolved == 0 && c->node_type != ONT_LOSS && at_least_one_drawn_child(n)) { /* solved_draw */
        n->node_type = ONT_DRAW;
        update_solved_terminal(n, c, ci, DRAW_VALUE);
        return 1;
    }
    return 0;
}

static void revert_virtual_loss_and_update(ONode* n, int ci, float value, const OSettings* st, int free_backup,
                                           int solve) { /* node.h:199-246 */
    n->value_sum += value;
    ++n->real_visits;
    if (n->n[ci] == 1) {
        n->q[ci] = value;
    } else {
        switch (get_virtual_style(st, n->n[ci])) {
            case OVS_VIRTUAL_LOSS:
                n->q[ci] = (float)(((double)n->q[ci] * n->n[ci] + 1 + value) / n->n[ci]);
                break;
            case OVS_VIRTUAL_VISIT: {
                const uint32_t real = n->n[ci] - n->vl[ci];
                n->q[ci] = (float)(((double)n->q[ci] * real + value) / (real + 1));
                break;
            }
            default: break;
        }
    }
    --n->vl[ci];
    if (free_backup) ++n->free_visits;
    if (solve) solve_for_terminal(n, ci);
}

static void backup_value(float value, const OSettings* st, const Traj* t, int free_backup, int solve) {
    /* node.h:819-843 with MODE_TWO_PLAYER; transposition branches are dead (SURVEY A-7): only the root has
       numberParentNodes != 1 and it is the last element visited. */
    for (int i = t->len - 1; i >= 0; --i) {
        value = -value;
        revert_virtual_loss_and_update(t->steps[i].node, t->steps[i].child_idx, value, st, free_backup, solve);
    }
}

static void revert_virtual_loss(ONode* n, int ci, const OSettings* st) { /* node.cpp:661-679 */
    if (get_virtual_style(st, n->n[ci]) == OVS_VIRTUAL_LOSS)
        n->q[ci] = (float)(((double)n->q[ci] * n->n[ci] + 1) / (n->n[ci] - 1));
    --n->n[ci];
    --n->visit_sum;
    --n->vl[ci];
}

/* ------------------------------------------------------------------ NN results -> node */
static void apply_temperature(float* p, const int* pidx, int n, float t) { /* util/blazeutil.h:78-88 */
    /* dist = pow(dist, 1/T); dist /= sum(dist).  blaze::sum reduces in the order of the vector = the order of
       Stockfish's move generator (absent); the restatement sums SEQUENTIALLY in ascending policy-index order, the
       generator-independent order the device uses too (ties cannot occur: an index names one move). */
    if (t == 1) return;
    int* order = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        p[i] = powf(p[i], 1.0f / t);
        int j = i; /* insertion by (pidx, i) */
        while (j > 0 && pidx[order[j - 1]] > pidx[i]) {
            order[j] = order[j - 1];
            --j;
        }
        order[j] = i;
    }
    float sum = 0.0f;
    for (int r = 0; r < n; ++r) sum += p[order[r]];
    free(order);
    for (int i = 0; i < n; ++i) p[i] /= sum;
}

static void fill_nn_results(OSearch* s, ONode* node, float value, const float* prob) {
    /* searchthread.cpp:290-299; set_probabilities_for_moves node.cpp:961-979 (mirroring inside the index lookup) */
    for (int i = 0; i < node->n_actions; ++i) node->policy[i] = node->pidx[i] >= 0 ? prob[node->pidx[i]] : 0.0f;
    apply_temperature(node->policy, node->pidx, node->n_actions, s->st.node_policy_temperature); /* node_post_process_policy */
    node_set_value(node, value);                                                      /* node_assign_value */
    node->has_nn = 1;
}

/* ------------------------------------------------------------------ public API */
OSearch* osearch_new(const OSettings* st) {
    OSearch* s = (OSearch*)calloc(1, sizeof(OSearch));
    s->st = *st;
    s->channels = oplanes_channels(st->mode, st->input_version);
    s->n_labels = opolicy_nb_policy_channels(st->mode) * 64;
    const int B = st->batch_size;
    s->new_nodes = (ONode**)calloc((size_t)B, sizeof(ONode*));
    s->new_stm = (int*)calloc((size_t)B, sizeof(int));
    s->new_traj = (Traj*)calloc((size_t)B, sizeof(Traj));
    s->coll_traj = (Traj*)calloc((size_t)B, sizeof(Traj));
    s->planes = (float*)calloc((size_t)B * (size_t)s->channels * 64, sizeof(float));
    s->actions_cap = 256;
    s->actions_buf = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)s->actions_cap);
    s->parked.new_nodes = (ONode**)calloc((size_t)B, sizeof(ONode*));
    s->parked.new_stm = (int*)calloc((size_t)B, sizeof(int));
    s->parked.new_traj = (Traj*)calloc((size_t)B, sizeof(Traj));
    s->parked.coll_traj = (Traj*)calloc((size_t)B, sizeof(Traj));
    s->parked.planes = (float*)calloc((size_t)B * (size_t)s->channels * 64, sizeof(float));
    s->parked.actions_cap = 256;
    s->parked.actions_buf = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)s->parked.actions_cap);
    s->active = 0;
    s->rng.x = minstd_seed(st->seed);
    glibc_srand(&s->crand, (unsigned)st->seed);
    return s;
}
/* makes logical search thread t (0 or 1) the owner of the batch members */
static void use_thread(OSearch* s, int t) {
    if (s->active == t) return;
    struct OParked tmp = s->parked;
#define SWAPF(f) s->parked.f = s->f, s->f = tmp.f
    SWAPF(new_nodes);
    SWAPF(new_stm);
    SWAPF(n_new);
    SWAPF(new_traj);
    SWAPF(coll_traj);
    SWAPF(n_coll);
    SWAPF(cur);
    SWAPF(actions_buf);
    SWAPF(n_actions_buf);
    SWAPF(actions_cap);
    SWAPF(planes);
#undef SWAPF
    s->active = t;
}
static void free_tree(OSearch* s) {
    for (size_t i = 0; i < s->n_all; ++i) node_free(s->all_nodes[i]);
    s->n_all = 0;
    s->root = NULL;
    s->next_root = NULL;
    s->next_root_valid = 0;
}
void osearch_free(OSearch* s) {
    use_thread(s, 0);
    for (int i = 0; i < s->st.batch_size; ++i) {
        free(s->parked.new_traj[i].steps);
        free(s->parked.coll_traj[i].steps);
    }
    free(s->parked.cur.steps);
    free(s->parked.new_nodes);
    free(s->parked.new_stm);
    free(s->parked.new_traj);
    free(s->parked.coll_traj);
    free(s->parked.planes);
    free(s->parked.actions_buf);
    free_tree(s);
    free(s->all_nodes);
    for (int i = 0; i < s->st.batch_size; ++i) {
        free(s->new_traj[i].steps);
        free(s->coll_traj[i].steps);
    }
    free(s->cur.steps);
    free(s->new_nodes);
    free(s->new_stm);
    free(s->new_traj);
    free(s->coll_traj);
    free(s->planes);
    free(s->actions_buf);
    free(s);
}
int osearch_channels(const OSearch* s) { return s->channels; }
int osearch_nb_labels(const OSearch* s) { return s->n_labels; }
const float* osearch_planes(const OSearch* s) { return s->planes; }
/* Threads = 2: the same entry points for logical search thread t */
int osearch_create_mini_batch_t(OSearch* s, int t) {
    use_thread(s, t);
    return osearch_create_mini_batch(s);
}
void osearch_apply_results_t(OSearch* s, int t, const float* values, const float* probs) {
    use_thread(s, t);
    osearch_apply_results(s, values, probs);
}
const float* osearch_planes_t(OSearch* s, int t) {
    use_thread(s, t);
    return s->planes;
}
void osearch_batch_keys_t(OSearch* s, int t, unsigned long long* out) {
    use_thread(s, t);
    osearch_batch_keys(s, out);
}
void osearch_batch_keys(const OSearch* s, unsigned long long* out) {
    if (s->n_new == 0) {
        out[0] = s->root->key;
        return;
    }
    for (int i = 0; i < s->n_new; ++i) out[i] = s->new_nodes[i]->key;
}

/* MCTSAgent::apply_move_to_tree + pick_next_node (mctsagent.cpp:230-247): the child behind `move` becomes the
 * candidate root of the next search; applying a second move (the opponent's reply) descends once more.  Returns 1 if a
 * candidate exists afterwards. */
int osearch_apply_move(OSearch* s, uint32_t move) {
    ONode* base = s->next_root_valid ? s->next_root : s->root;
    s->next_root = NULL;
    s->next_root_valid = 1; /* from now on only next_root counts (NULL = the tree is dropped at the next search) */
    if (base == NULL || !base->has_d) return 0; /* is_playout_node */
    for (int i = 0; i < base->n_actions; ++i)
        if (base->actions[i] == move) {
            s->next_root = base->child[i];
            break;
        }
    return s->next_root != NULL;
}

/* MCTSAgent::evaluate_board_state :292-296 with init_root_node / get_root_node_from_tree (:113-160): 0 = nothing to
 * search, 1 = new tree (the root needs its network evaluation), 2 = the candidate root of the former tree is reused. */
int osearch_set_root(OSearch* s, const OPos* pos) {
    ONode* cand = s->next_root_valid ? s->next_root : NULL;
    s->next_root = NULL;
    s->next_root_valid = 0;
    s->sum_select_k = s->sum_depth = 0;
    use_thread(s, 0);
    s->n_new = s->n_coll = 0;
    s->parked.n_new = s->parked.n_coll = 0;
    if (cand != NULL && cand->key == pos->key && cand->has_d && cand->visit_sum - cand->free_visits > 0) {
        /* the rest of the old tree is only garbage-collected by the reference; here it stays allocated */
        s->root = cand;
        opos_copy(&s->root_state, pos);
        s->root->number_parents = 0; /* make_to_root */
        if (s->root->is_terminal || s->root->n_actions == 0) return 0;
        return 2;
    }
    free_tree(s);
    s->num_nodes = 0;
    opos_copy(&s->root_state, pos);
    s->root = node_new(s, pos);
    s->root->number_parents = 0; /* make_to_root */
    if (s->root->is_terminal || s->root->n_actions == 0) return 0;
    oplanes_encode(pos, s->st.mode, s->st.input_version, 1, s->planes);
    return 1;
}

static void root_noise_and_open(OSearch* s) {
    if (s->st.dirichlet_epsilon > 0.009f) { /* mctsagent.cpp:311-316 */
        float* noise = (float*)malloc(sizeof(float) * (size_t)s->root->n_actions);
        /* the reference's generator is process-wide (util/randomgen.h:35) and advances from search to search; here it
           belongs to the OSearch, seeded once from the settings */
        dirichlet_noise_g(&s->rng, s->root->n_actions, s->st.dirichlet_alpha, noise);
        for (int i = 0; i < s->root->n_actions; ++i)
            s->root->policy[i] = (1 - s->st.dirichlet_epsilon) * s->root->policy[i] + s->st.dirichlet_epsilon * noise[i];
        free(noise);
        fully_expand_node(s->root);
    }
}
void osearch_root_results(OSearch* s, const float* value, const float* prob) {
    fill_nn_results(s, s->root, value[0], prob);
    prepare_node_for_visits(s->root);
    root_noise_and_open(s);
}
/* reused root: it already has its network results and NodeData; the noise is applied to it like to a new root */
void osearch_root_reused(OSearch* s) { root_noise_and_open(s); }

static int get_best_action_index(const ONode* n, const OSettings* st, int fast);

/* ---- epsilon-greedy / epsilon-check exploration (searchthread.cpp:124-185, :451-473, :497-501) */
static int get_random_depth(OSearch* s) { /* :497-501: ceil(-log2(1 - r/100) - 1), r = rand() % 100 + 1 */
    const int r = glibc_rand(&s->crand) % 100 + 1;
    const double d = ceil(-log2(1 - r / 100.0) - 1);
    /* r = 100: size_t(+inf) -- formally undefined; GCC on x86-64 converts through cvttsd2si(x - 2^63) ^ 2^63 = 0
       (checked against the compiled reference, tests/test_ref_mcts.py): depth 0 */
    return d > 1e18 ? 0 : (int)d;
}
static void push_action(OSearch* s, uint32_t a) {
    if (s->n_actions_buf == s->actions_cap) {
        s->actions_cap *= 2;
        s->actions_buf = (uint32_t*)realloc(s->actions_buf, sizeof(uint32_t) * (size_t)s->actions_cap);
    }
    s->actions_buf[s->n_actions_buf++] = a;
}
/* :144-162: walks down the most visited line for a random number of plies; the trajectory starts at the node reached */
static ONode* get_starting_node(OSearch* s, ONode* cur, int* depth, int* ci) {
    const int d = get_random_depth(s);
    for (int k = 0; k < d; ++k) {
        *ci = get_best_action_index(cur, &s->st, 1);
        ONode* next = cur->child[*ci];
        if (next == NULL || !next->has_d || next->visit_sum < (uint32_t)s->st.epsilon_greedy_counter || next->node_type != ONT_UNSOLVED) break;
        push_action(s, cur->actions[*ci]);
        cur = next;
        ++*depth;
    }
    return cur;
}
static void random_playout(OSearch* s, ONode* cur, int* ci) { /* :124-142 */
    if (cur->n_actions == cur->no_visit_idx) { /* is_fully_expanded */
        const int idx = (int)((size_t)glibc_rand(&s->crand) % (size_t)cur->n_actions);
        ONode* c = cur->child[idx];
        if (c == NULL || !c->has_d || c->node_type == ONT_UNSOLVED) {
            *ci = idx;
            return;
        }
        *ci = -1;
    } else {
        *ci = cur->no_visit_idx < cur->n_actions - 1 ? cur->no_visit_idx : cur->n_actions - 1;
        increment_no_visit_idx(cur);
    }
}
static int select_enhanced_move(OSearch* s, ONode* cur) { /* :451-473: an unopened move that gives check */
    if (cur->has_d && !cur->inspected && !cur->is_terminal) {
        OPos* pos = (OPos*)malloc(sizeof(OPos));
        opos_copy(pos, &s->root_state);
        for (int i = 0; i < s->n_actions_buf; ++i) opos_do_move(pos, s->actions_buf[i]);
        for (int c = cur->no_visit_idx; c < cur->n_actions; ++c)
            if (opos_gives_check(pos, cur->actions[c])) {
                for (int i = cur->no_visit_idx; i < c + 1; ++i) increment_no_visit_idx(cur);
                free(pos);
                return c;
            }
        free(pos);
        cur->inspected = 1;
    }
    return -1;
}

static ONode* get_new_child_to_evaluate(OSearch* s, int* type) { /* searchthread.cpp:164-271 */
    ONode* cur = s->root;
    int depth = 0;
    int forced = -1; /* childIdx chosen by the exploration branches, (uint16_t)-1 in the reference */
    const int egc = s->st.epsilon_greedy_counter, ecc = s->st.epsilon_checks_counter;
    if (egc && s->root->has_d && glibc_rand(&s->crand) % egc == 0) {
        cur = get_starting_node(s, cur, &depth, &forced);
        random_playout(s, cur, &forced);
    } else if (ecc && s->root->has_d && glibc_rand(&s->crand) % ecc == 0) {
        cur = get_starting_node(s, cur, &depth, &forced);
        forced = select_enhanced_move(s, cur);
        if (forced == -1) random_playout(s, cur, &forced);
    }
    for (;;) {
        const int ci = forced != -1 ? forced : select_child_node(s, cur);
        forced = -1;
        apply_virtual_loss_to_child(cur, ci, &s->st);
        traj_push(&s->cur, cur, ci);
        ONode* next = cur->child[ci];
        depth++;
        if (next == NULL) {
            /* newState = rootState->clone(); replay actionsBuffer; do_action(move) (:201-207) */
            OPos* ns = (OPos*)malloc(sizeof(OPos));
            opos_copy(ns, &s->root_state);
            for (int i = 0; i < s->n_actions_buf; ++i) opos_do_move(ns, s->actions_buf[i]);
            opos_do_move(ns, cur->actions[ci]);
            increment_no_visit_idx(cur);
            next = node_new(s, ns); /* add_new_node_to_tree; MCGS hash linking is dead code (SURVEY A-7) */
            cur->child[ci] = next;
            if (next->is_terminal) {
                *type = NB_TERMINAL;
            } else {
                *type = NB_NEW;
                oplanes_encode(ns, s->st.mode, s->st.input_version, 1,
                               s->planes + (size_t)s->n_new * (size_t)s->channels * 64);
                s->new_stm[s->n_new] = ns->stm;
            }
            free(ns);
            s->sum_depth += (unsigned long long)depth;
            return next;
        }
        if (next->is_terminal) {
            *type = NB_TERMINAL;
            s->sum_depth += (unsigned long long)depth;
            return next;
        }
        if (!next->has_nn) {
            *type = NB_COLLISION;
            s->sum_depth += (unsigned long long)depth;
            return next;
        }
        push_action(s, cur->actions[ci]);
        cur = next;
    }
}

int osearch_create_mini_batch(OSearch* s) { /* searchthread.cpp:347-380 */
    const int B = s->st.batch_size;
    int num_terminal = 0;
    const int terminal_cache = 2 * B;
    while (s->n_new < B && s->n_coll != B && num_terminal < terminal_cache) {
        s->cur.len = 0;
        s->n_actions_buf = 0;
        int type;
        ONode* node = get_new_child_to_evaluate(s, &type);
        if (type == NB_TERMINAL) {
            ++num_terminal;
            backup_value(node_get_value(node), &s->st, &s->cur, 1, s->st.mcts_solver);
        } else if (type == NB_COLLISION) {
            traj_copy(&s->coll_traj[s->n_coll++], &s->cur);
        } else {
            s->new_nodes[s->n_new] = node;
            traj_copy(&s->new_traj[s->n_new], &s->cur);
            s->n_new++;
        }
    }
    return s->n_new;
}

void osearch_apply_results(OSearch* s, const float* values, const float* probs) {
    /* set_nn_results_to_child_nodes (:301-310) */
    for (int b = 0; b < s->n_new; ++b)
        fill_nn_results(s, s->new_nodes[b], values[b], probs + (size_t)b * (size_t)s->n_labels);
    /* backup_value_outputs (:312-317, :428-439) */
    for (int b = 0; b < s->n_new; ++b) backup_value(node_get_value(s->new_nodes[b]), &s->st, &s->new_traj[b], 0, 0);
    s->n_new = 0;
    /* backup_collisions (:319-324) */
    for (int c = 0; c < s->n_coll; ++c)
        for (int i = s->coll_traj[c].len - 1; i >= 0; --i)
            revert_virtual_loss(s->coll_traj[c].steps[i].node, s->coll_traj[c].steps[i].child_idx, &s->st);
    s->n_coll = 0;
}

int osearch_continue(const OSearch* s) { /* searchthread.cpp:326-340, :418-426 */
    const ONode* r = s->root;
    if (r == NULL || !r->has_d) return 0;
    const unsigned node_count = r->visit_sum - r->free_visits;
    const int limits_ok = (s->st.nodes == 0 || node_count < s->st.nodes) &&
                          (s->st.simulations == 0 || r->visit_sum < s->st.simulations);
    return limits_ok && r->node_type == ONT_UNSOLVED;
}

/* ------------------------------------------------------------------ results */
int osearch_root_num_children(const OSearch* s) { return s->root->n_actions; }
int osearch_root_no_visit_idx(const OSearch* s) { return s->root->has_d ? s->root->no_visit_idx : 0; }
float osearch_root_value(const OSearch* s) { return node_get_value(s->root); }
unsigned osearch_root_visits(const OSearch* s) { return s->root->has_d ? s->root->visit_sum : 0; }
unsigned osearch_root_free_visits(const OSearch* s) { return s->root->has_d ? s->root->free_visits : 0; }
int osearch_root_node_type(const OSearch* s) { return s->root->has_d ? s->root->node_type : ONT_UNSOLVED; }
unsigned long long osearch_num_nodes(const OSearch* s) { return s->num_nodes; }
unsigned long long osearch_sum_select_k(const OSearch* s) { return s->sum_select_k; }
unsigned long long osearch_sum_depth(const OSearch* s) { return s->sum_depth; }

static int argmax_d(const double* v, int n) {
    int b = 0;
    for (int i = 1; i < n; ++i)
        if (v[i] > v[b]) b = i;
    return b;
}

/* Node::get_mcts_policy node.cpp:1070-1109; out has n_actions entries (padded with 0 beyond noVisitIdx as
 * update_eval_info does, evalinfo.cpp:208-214).  Returns bestMoveIdx. */
static int get_mcts_policy(const ONode* n, const OSettings* st, double* out) {
    const int k = n->no_visit_idx;
    for (int i = 0; i < n->n_actions; ++i) out[i] = 0.0;
    if (n->node_type == ONT_WIN) { /* mcts_policy_based_on_wins */
        for (int i = 0; i < k; ++i)
            if (n->child[i] && n->child[i]->has_d && n->child[i]->node_type == ONT_LOSS) out[i] = 1.0;
    } else if (n->node_type == ONT_LOSS) { /* mcts_policy_based_on_losses */
        int longest = 0, end = 0;
        for (int i = 0; i < k; ++i)
            if (n->child[i] && n->child[i]->has_d && n->child[i]->end_in_ply > end) end = n->child[i]->end_in_ply, longest = i;
        out[longest] = 1.0;
    } else {
        for (int i = 0; i < k; ++i) out[i] = (double)n->n[i];
        /* prune_losses_in_mcts_policy node.cpp:340-363 */
        if (n->n_unsolved != n->n_actions && n->node_type != ONT_LOSS)
            for (int i = 0; i < k; ++i)
                if (n->child[i] && n->child[i]->has_d && n->child[i]->node_type == ONT_WIN) out[i] = 0;
        if (st->q_value_weight > 0) {
            int best_q = 0;
            for (int i = 1; i < k; ++i)
                if (n->q[i] > n->q[best_q]) best_q = i;
            /* first_and_second_max blazeutil.h:155-180 */
            double first = out[0], second = 2.2250738585072014e-308; /* numeric_limits<double>::min() */
            int first_arg = 0, second_arg = 0;
            for (int i = 1; i < k; ++i) {
                if (out[i] > first) {
                    second = first;
                    second_arg = first_arg;
                    first = out[i];
                    first_arg = i;
                } else if (out[i] > second) {
                    second = out[i];
                    second_arg = i;
                }
            }
            const int best = first_arg;
            if (st->q_veto_delta != 0 && best_q != best && n->q[best_q] > n->q[best] + st->q_veto_delta && n->n[best_q] > 1) {
                if (out[best] > out[best_q]) {
                    const double save = out[best_q];
                    out[best_q] = out[best];
                    out[best] = save;
                }
            } else if (best != second_arg && n->q[second_arg] > n->q[best]) {
                const float q_diff = n->q[second_arg] - n->q[best];
                out[second_arg] += q_diff * st->q_value_weight * out[best];
            }
        }
    }
    double sum = 0;
    for (int i = 0; i < k; ++i) sum += out[i];
    for (int i = 0; i < k; ++i) out[i] /= sum;
    return argmax_d(out, k);
}

static int get_best_action_index(const ONode* n, const OSettings* st, int fast) { /* node.cpp:1123-1148 */
    if (n->checkmate_idx != NO_CHECKMATE) return n->checkmate_idx;
    if (n->node_type == ONT_LOSS) {
        int longest = 0, idx = 0;
        for (int i = 0; i < n->n_actions; ++i)
            if (n->child[i]->end_in_ply > longest) longest = n->child[i]->end_in_ply, idx = i;
        return idx;
    }
    if (fast) {
        int b = 0;
        for (int i = 1; i < n->no_visit_idx; ++i)
            if (n->n[i] > n->n[b]) b = i;
        return b;
    }
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(n->n_actions > 0 ? n->n_actions : 1));
    const int b = get_mcts_policy(n, st, tmp);
    free(tmp);
    return b;
}

void osearch_root_stats(const OSearch* s, uint32_t* moves, uint32_t* visits, float* q, float* prior, double* mcts_policy) {
    const ONode* r = s->root;
    for (int i = 0; i < r->n_actions; ++i) {
        if (moves) moves[i] = r->actions[i];
        if (visits) visits[i] = i < r->no_visit_idx ? r->n[i] : 0;
        if (q) q[i] = i < r->no_visit_idx ? r->q[i] : (float)LOSS_VALUE; /* evalinfo.cpp:213 */
        if (prior) prior[i] = r->policy[i];
    }
    if (mcts_policy) {
        if (r->n_actions == 1) mcts_policy[0] = 1.0;
        else get_mcts_policy(r, &s->st, mcts_policy);
    }
}
int osearch_best_move_idx(const OSearch* s) { return get_best_action_index(s->root, &s->st, 0); }

static float value_display(const ONode* n) { /* node.cpp:604-616 */
    if (n->node_type == ONT_WIN) return WIN_VALUE;
    if (n->node_type == ONT_LOSS) return LOSS_VALUE;
    if (n->node_type == ONT_DRAW) return DRAW_VALUE;
    return node_get_value(n);
}
float osearch_best_move_q(const OSearch* s) { /* set_eval_for_single_pv evalinfo.cpp:123-178, idx 0 */
    const ONode* r = s->root;
    const int ci = get_best_action_index(r, &s->st, 0);
    const ONode* next = r->child[ci];
    if (next == NULL) return Q_INIT;
    if (next->has_d) return -value_display(next);
    return -node_get_value(next);
}
int osearch_pv(const OSearch* s, uint32_t* out, int max_len) { /* get_principal_variation node.cpp:1111-1121 */
    const ONode* r = s->root;
    int n = 0;
    int ci = get_best_action_index(r, &s->st, 0);
    if (n < max_len) out[n++] = r->actions[ci];
    const ONode* cur = r->child[ci];
    while (cur != NULL && cur->has_d && !cur->is_terminal && n < max_len) {
        ci = get_best_action_index(cur, &s->st, 1);
        out[n++] = cur->actions[ci];
        cur = cur->child[ci];
    }
    return n;
}

/* oracle/mcts.h -- CPU ORACLE (test infrastructure only): single-threaded restatement of the reference search. */
#ifndef ORACLE_MCTS_H
#define ORACLE_MCTS_H
#include "chess.h"
#include "planes.h"
#include "policy.h"
#ifdef __cplusplus
extern "C" {
#endif

enum { OVS_VIRTUAL_LOSS = 0, OVS_VIRTUAL_VISIT = 1, OVS_VIRTUAL_OFFSET = 2, OVS_VIRTUAL_MIX = 3 };
enum { ONT_WIN = 0, ONT_DRAW = 1, ONT_LOSS = 2, ONT_UNSOLVED = 3 }; /* nodedata.h NodeType without TB support */

/* SearchSettings + SearchLimits of the reference (agents/config/searchsettings.h:51-98, searchlimits.h:37-61);
 * the same plain struct is the argument of the product's ara_search_create (include/ara_b200.h). */
typedef struct OSettings {
    int batch_size;
    float dirichlet_epsilon;
    float dirichlet_alpha;
    float node_policy_temperature;
    float q_value_weight;
    float q_veto_delta;
    float cpuct_init;
    float cpuct_base;
    int mcts_solver;
    int virtual_style;
    unsigned virtual_mix_threshold;
    unsigned simulations; /* SearchLimits::simulations (0 = unlimited) */
    unsigned nodes;       /* SearchLimits::nodes */
    unsigned long long seed; /* explicit seed for the Dirichlet noise generator */
    int mode;             /* OMODE_* build mode: decides plane layout and label set */
    int input_version;    /* 1, 2, 3 */
    int threads;                /* SearchSettings::threads: 1, or 2 in the fixed schedule described below */
    int epsilon_greedy_counter; /* SearchSettings::epsilonGreedyCounter, 0 = off */
    int epsilon_checks_counter; /* SearchSettings::epsilonChecksCounter, 0 = off */
    int reserved;
} OSettings;

typedef struct OSearch OSearch;

void osettings_default(OSettings* s, int mode);
OSearch* osearch_new(const OSettings* s);
void osearch_free(OSearch* s);
int osearch_channels(const OSearch* s);
int osearch_nb_labels(const OSearch* s); /* policy-map length P*64 */
/* MCTSAgent::evaluate_board_state, first half: new root for `pos`; root planes are placed in slot 0 of planes() */
int osearch_set_root(OSearch* s, const OPos* pos); /* 0 nothing to search, 1 new tree, 2 reused subtree */
int osearch_apply_move(OSearch* s, uint32_t move);  /* MCTSAgent::apply_move_to_tree */
void osearch_root_reused(OSearch* s);
/* set_root_node_predictions second half + prepare_node_for_visits + optional Dirichlet noise */
void osearch_root_results(OSearch* s, const float* value, const float* prob);
/* SearchThread::create_mini_batch: returns the number of new leaves whose planes sit in planes() */
int osearch_create_mini_batch(OSearch* s);
/* set_nn_results_to_child_nodes + backup_value_outputs + backup_collisions */
void osearch_apply_results(OSearch* s, const float* values, const float* probs);
/* Threads = 2 (the reference's default, uci/optionsuci.cpp:182): the same three entry points for logical search thread
 * t in {0, 1}.  The two threads take turns on the tree in a fixed schedule (oracle/search.py, Search.run(threads=2)):
 *   sel(0) sel(1) | bk(0) sel(0) bk(1) sel(1) | bk(0) sel(0) ...   with every thread testing the loop condition before
 * its own sel -- one of the interleavings two real threads of the reference can produce (each phase atomic). */
int osearch_create_mini_batch_t(OSearch* s, int t);
void osearch_apply_results_t(OSearch* s, int t, const float* values, const float* probs);
const float* osearch_planes_t(OSearch* s, int t);
void osearch_batch_keys_t(OSearch* s, int t, unsigned long long* out);
/* run_search_thread loop condition: is_running && nodes_limits_ok && is_root_node_unsolved */
int osearch_continue(const OSearch* s);
const float* osearch_planes(const OSearch* s);
/* Zobrist keys of the positions whose planes sit in planes(): slot 0 = root after set_root, else the new leaves */
void osearch_batch_keys(const OSearch* s, unsigned long long* out);
void ofake_eval(unsigned long long key, int n_labels, float* value, float* prob);

/* results (update_eval_info, evalinfo.cpp:195-249) */
int osearch_root_num_children(const OSearch* s);
int osearch_root_no_visit_idx(const OSearch* s);
void osearch_root_stats(const OSearch* s, uint32_t* moves, uint32_t* visits, float* q, float* prior, double* mcts_policy);
float osearch_root_value(const OSearch* s);
unsigned osearch_root_visits(const OSearch* s);
unsigned osearch_root_free_visits(const OSearch* s);
int osearch_best_move_idx(const OSearch* s);
float osearch_best_move_q(const OSearch* s);
int osearch_root_node_type(const OSearch* s);
int osearch_pv(const OSearch* s, uint32_t* out, int max_len);
unsigned long long osearch_num_nodes(const OSearch* s);
/* traffic accounting for the HBM roofline of the select/backup kernels */
unsigned long long osearch_sum_select_k(const OSearch* s);
unsigned long long osearch_sum_depth(const OSearch* s);
/* Dirichlet noise as std::gamma_distribution<float> over std::default_random_engine (util/blazeutil.h:113-124) */
void odirichlet_noise(unsigned long long seed, int n, float alpha, float* out);

#ifdef __cplusplus
}
#endif
#endif

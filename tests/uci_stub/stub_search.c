/* Test scaffolding: stand-in for the SEARCH entry points of the C-ABI, pre-loaded in front of libara_b200.so so that the
 * UCI front-end's command loop (threads, `go infinite` / `stop`, ordering of its answers) can be exercised on a box
 * without a GPU.  The "search" just waits until its move time is over or it is told to stop.  Never part of the product. */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ara_b200.h"

static volatile int g_stop;
static double g_movetime_ms, g_elapsed_ms;
static unsigned g_polls;

/* ARA_STUB_LOG=<file>: one line per call of interest, for the tests to read */
static void log_call(const char* fmt, double a, double b, double c, double d) {
    const char* path = getenv("ARA_STUB_LOG");
    if (!path) return;
    FILE* f = fopen(path, "a");
    if (!f) return;
    fprintf(f, fmt, a, b, c, d);
    fputc('\n', f);
    fclose(f);
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

ara_search_t ara_search_create(ara_net_t net, const ara_search_settings_t* s, int device, int n_trees, int max_nodes) {
    (void)net, (void)s, (void)device, (void)n_trees, (void)max_nodes;
    return (ara_search_t)0x1;
}
void ara_search_destroy(ara_search_t s) { (void)s; }
int ara_search_set_position(ara_search_t s, int tree, const ara_board_t* root, const unsigned long long* keys, const short* reps, int n) {
    (void)s, (void)tree, (void)root, (void)keys, (void)reps, (void)n;
    return 0;
}
int ara_search_go(ara_search_t s) {
    (void)s;
    const double t0 = now_ms();
    const double limit = g_movetime_ms > 0 ? g_movetime_ms : 5000.0; /* "infinite": until stop (bounded for the test) */
    g_stop = 0; /* like the library: a stop that came before the search started is forgotten */
    g_polls = 0;
    while (!g_stop && now_ms() - t0 < limit) {
        struct timespec ts = {0, 2000000};
        nanosleep(&ts, 0);
        ++g_polls;
    }
    g_elapsed_ms = now_ms() - t0;
    log_call("go movetime=%.0f%.0s%.0s%.0s", g_movetime_ms, 0, 0, 0);
    return 0;
}
int ara_search_stop(ara_search_t s) {
    (void)s;
    g_stop = 1;
    return 0;
}
int ara_search_result(ara_search_t s, int tree, ara_search_result_t* out) {
    (void)s, (void)tree;
    memset(out, 0, sizeof(*out));
    out->n_moves = 1;
    out->no_visit_idx = 1;
    out->node_type = 3;
    out->pv_len = 1;
    out->moves[0] = out->pv[0] = (unsigned short)(12 | (28 << 6)); /* e2e4 */
    out->visits[0] = g_polls + 1;
    out->visit_sum = g_polls + 1;
    out->policy[0] = 1.0;
    return 0;
}
double ara_search_last_go_ms(ara_search_t s) { (void)s; return g_elapsed_ms; }
int ara_search_apply_move(ara_search_t s, int tree, unsigned short move) {
    (void)s;
    log_call("apply_move tree=%.0f from=%.0f to=%.0f flag=%.0f", tree, move & 63, (move >> 6) & 63, move >> 12);
    return 0;
}
int ara_search_set_movetime(ara_search_t s, double ms) { (void)s; g_movetime_ms = ms; return 0; }
int ara_search_set_time_control(ara_search_t s, const ara_time_control_t* tc) {
    (void)s;
    if (tc) log_call("time_control movetime=%.0f in_game=%.0f can_prolong=%.0f nps_known=%.0f", tc->movetime_ms, tc->in_game, tc->can_prolong,
                     tc->overall_nps > 0 ? 1 : 0);
    else log_call("time_control off%.0s%.0s%.0s%.0s", 0, 0, 0, 0);
    return 0;
}
int ara_search_time_report(ara_search_t s, ara_time_report_t* out) { (void)s; memset(out, 0, sizeof(*out)); return 0; }

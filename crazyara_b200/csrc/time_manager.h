// Host-side stop / prolong decisions of the reference's ThreadManager (manager/threadmanager.cpp:114-178) as pure
// functions of the root statistics, so that they can be unit-tested without a clock.  The search loop (search.cu)
// evaluates them every update interval on statistics read back from the device tree.
#pragma once
#include "../../include/ara_b200.h"

namespace ara {

struct RootStatsHost {  // mirror of RootTimeStats (search_dev.cuh)
    unsigned node_count, first_visits, second_visits;
    float q_first, q_second;
    int max_q_is_max_visits;
    float value_eval;
    int valid;
};

// ThreadManager::early_stopping: 1 "max nodes" rule, 2 "second move cannot catch up" rule, 0 keep searching
inline int tm_early_stopping(const ara_time_control_t& tc, double remaining_ms, const RootStatsHost& r) {
    if (!tc.in_game || tc.overall_nps == 0.0 || !r.valid) return 0;
    if (r.node_count > tc.overall_nps * (tc.movetime_ms / 1000.0f) * 2 && r.max_q_is_max_visits) return 1;
    // (the reference divides the NPS by 1000 in float before multiplying by the remaining milliseconds)
    if (r.second_visits + static_cast<float>(remaining_ms) * (static_cast<float>(tc.overall_nps) / 1000) < r.first_visits * 2.0f &&
        r.q_first > r.q_second)
        return 2;
    return 0;
}

// ThreadManager::continue_search: prolong by another move time if the evaluation dropped since the last move.
// `checked` is checkedContinueSearch, `last_eval` tData->lastValueEval; both are updated like the reference does.
inline bool tm_continue_search(const ara_time_control_t& tc, double remaining_ms, const RootStatsHost& r, int* checked,
                               float* last_eval) {
    if (!tc.in_game || !tc.can_prolong || tc.overall_nps == 0.0 || *checked > 1) return false;
    if (tc.movetime_ms * 2 > tc.safe_remaining_ms) return false;  // make sure not to flag when continuing
    if (r.value_eval < *last_eval) {
        if (remaining_ms < tc.update_interval_ms + tc.move_overhead_ms) return false;
        *last_eval = r.value_eval;
        ++*checked;
        return true;
    }
    return false;
}

}  // namespace ara

/* oracle/policy.h -- CPU ORACLE (test infrastructure only): move <-> policy-vector index. */
#ifndef ORACLE_POLICY_H
#define ORACLE_POLICY_H
#include "chess.h"
#include "planes.h"
#ifdef __cplusplus
extern "C" {
#endif
int opolicy_nb_labels(int mode);          /* 2272 / 1968 / 2316 (boardstate.h:51-60) */
int opolicy_nb_policy_channels(int mode); /* 81 / 76 / 84 (boardstate.h:246-254) */
const char* opolicy_label(int mode, int idx);
int opolicy_flat_plane_idx(int mode, int label_idx);
int opolicy_label_index(int mode, const char* uci); /* -1 if absent */
void opolicy_mirror(const char* uci, char* out);    /* sfutil.cpp:183-197 */
/* StateConstants::action_to_index<normal|classic, mirrored?> for a legal move of pos (node.cpp:961-979) */
int opolicy_move_index(const OPos* pos, uint32_t move, int mode, int is_policy_map);
#ifdef __cplusplus
}
#endif
#endif

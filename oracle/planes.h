/* oracle/planes.h -- CPU ORACLE (test infrastructure only): board -> NN input planes. */
#ifndef ORACLE_PLANES_H
#define ORACLE_PLANES_H
#include "chess.h"
#ifdef __cplusplus
extern "C" {
#endif
/* the reference's compile-time product modes (engine/CMakeLists.txt:5-24) */
enum { OMODE_CRAZYHOUSE = 0, OMODE_CHESS = 1, OMODE_LICHESS = 2 };
int oplanes_channels(int mode, int version);
/* writes [C,8,8] floats; returns C or -1 */
int oplanes_encode(const OPos* pos, int mode, int version, int normalize, float* out);
#ifdef __cplusplus
}
#endif
#endif

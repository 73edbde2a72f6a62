// Test infrastructure: a C entry point over the UNMODIFIED reference TimeManager, compiled from the reference's own
// sources where they lie (/root/reference/engine/src/manager/timemanager.cpp, agents/config/searchlimits.cpp -- the two
// files of the engine that build without its absent submodules).  Built by `make -C oracle ref` into oracle/_ref/; used
// only by tests/ to pin ara_time_for_move (and, through tests/golden/timeman.json, on boxes without /root/reference).
#include "manager/timemanager.h"

extern "C" int ref_time_for_move(long movetime, int wtime, int btime, int winc, int binc, int movestogo, int move_overhead,
                                 int me, int move_number) {
    SearchLimits limits;
    limits.movetime = movetime;
    limits.time[0] = wtime;
    limits.time[1] = btime;
    limits.inc[0] = winc;
    limits.inc[1] = binc;
    limits.movestogo = movestogo;
    limits.moveOverhead = move_overhead;
    // MCTSAgent's TimeManager (agents/mctsagent.cpp: make_unique<TimeManager>(randomMoveFactor, ...)) with the random
    // factor off and the engine's constants (constants.h:94-98)
    TimeManager tm(0.0f);  // expected game length, threshold move, moves to go, increment factor: the header's defaults
    return tm.get_time_for_move(&limits, static_cast<SideToMove>(me), move_number);
}

// oracle/ref/pommermanstate.h -- TEST INFRASTRUCTURE.  The environment the reference's MCTS sources are compiled against
// when oracle/Makefile builds oracle/_ref/libref_mcts.so.
//
// The reference selects its environment in stateobj.h by build mode; its chess environment (boardstate.h) sits on the
// absent Stockfish fork, so the build uses the one mode whose header stateobj.h includes by a plain name --
// `-DMODE_POMMERMAN` -> #include "pommermanstate.h" -- and this file provides that header: `PommermanState` here is a
// State (engine/src/state.h:287-509) over the repository's own rules / planes / policy oracle (oracle/chess.c, planes.c,
// policy.c -- themselves pinned to the reference's perft / plane / label vectors).  Nothing of the reference's
// Pommerman environment is involved; the name is the include hook.  MODE_POMMERMAN changes nothing else in the compiled
// sources except TERMINAL_NODE_CACHE (constants.h:101-105), a cache size.
#pragma once
#include <algorithm>
#include <cstring>
#include <iostream>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "state.h"

extern "C" {
#include "../chess.h"
#include "../planes.h"
#include "../policy.h"
}

namespace refshim {
struct Config {
    int mode = OMODE_CRAZYHOUSE;  // the reference's compile-time product mode, a run-time value here
    int input_version = 1;
    int n_aux = 0;
};
inline Config& config() {
    static Config c;
    return c;
}
// Action -> policy index is position-free in the reference (MV_LOOKUP tables, outputrepresentation.cpp:39-56); the
// oracle computes it from (position, move), so legal_actions() records it here.  Key: move | mirrored << 20 | is960 << 21.
inline std::unordered_map<uint32_t, int>& index_cache() {
    static std::unordered_map<uint32_t, int> m;
    return m;
}
// key of the position whose planes were last written to a given address (the hash-derived fake network needs keys)
inline std::unordered_map<const float*, unsigned long long>& plane_keys() {
    static std::unordered_map<const float*, unsigned long long> m;
    return m;
}
// the two maps are shared by the search threads when the reference runs its own OS threads (bench.py's CPU arm)
inline std::mutex& maps_mutex() {
    static std::mutex m;
    return m;
}
}  // namespace refshim

class StateConstantsPommerman : public StateConstantsInterface<StateConstantsPommerman> {
   public:
    static uint BOARD_WIDTH() { return 8; }
    static uint BOARD_HEIGHT() { return 8; }
    static uint NB_CHANNELS_TOTAL() { return oplanes_channels(refshim::config().mode, refshim::config().input_version); }
    static uint NB_LABELS() { return opolicy_nb_labels(refshim::config().mode); }
    static uint NB_LABELS_POLICY_MAP() { return opolicy_nb_policy_channels(refshim::config().mode) * 64; }
    static uint NB_AUXILIARY_OUTPUTS() { return refshim::config().n_aux; }
    static uint NB_PLAYERS() { return 2; }
    static std::string action_to_uci(Action action, bool is960) {
        OPos p;
        memset(&p, 0, sizeof(p));
        p.chess960 = is960;
        char buf[16];
        opos_move_to_uci(&p, static_cast<uint32_t>(action), buf);
        return buf;
    }
    template <PolicyType p, MirrorType m>
    static MoveIdx action_to_index(Action action) {
        std::lock_guard<std::mutex> lock(refshim::maps_mutex());
        auto& c = refshim::index_cache();
        const uint32_t base = static_cast<uint32_t>(action) | (m == mirrored ? 1u << 20 : 0u);
        auto it = c.find(base);
        if (it == c.end()) it = c.find(base | (1u << 21));
        return it == c.end() ? MoveIdx(0) : MoveIdx(it->second);
    }
    static void init(bool, bool) {}
    static std::vector<std::string> available_variants() {
        return {"chess", "crazyhouse", "kingofthehill", "3check", "giveaway", "atomic", "horde", "racingkings"};
    }
    static std::string start_fen(int variant) { return opos_start_fen(variant); }
};

class PommermanState : public State {
   public:
    OPos pos;
    PommermanState() { memset(&pos, 0, sizeof(pos)); }
    bool mirror_policy(SideToMove side) const { return side != 0 && pos.variant != OV_RACE; }  // flip_board, sfutil.h:135
    // Move ORDER: the reference's is that of Stockfish's generator (absent).  The order matters in one place only -- the
    // float sum that renormalises the priors in apply_temperature (util/blazeutil.h:78-88) runs over the node's moves in
    // this order -- so this environment fixes a generator-independent one: ascending policy index.  oracle/mcts.c and the
    // device sum in the same order.
    std::vector<Action> legal_actions() const override {
        uint32_t mv[OPOS_MAX_MOVES];
        const int n = opos_legal_moves(&pos, mv);
        std::vector<std::pair<int, uint32_t>> keyed(n);
        std::lock_guard<std::mutex> lock(refshim::maps_mutex());
        auto& c = refshim::index_cache();
        for (int i = 0; i < n; ++i) {
            const int idx = opolicy_move_index(&pos, mv[i], refshim::config().mode, 1);
            keyed[i] = {idx, mv[i]};
            c[mv[i] | (mirror_policy(pos.stm) ? 1u << 20 : 0u) | (pos.chess960 ? 1u << 21 : 0u)] = idx;
        }
        std::sort(keyed.begin(), keyed.end());
        std::vector<Action> out(n);
        for (int i = 0; i < n; ++i) out[i] = static_cast<Action>(keyed[i].second);
        return out;
    }
    void set(const std::string& fen, bool is960, int variant) override { opos_set(&pos, fen.c_str(), variant, is960); }
    void get_state_planes(bool normalize, float* planes, Version version) const override {
        int v = version::get_major(version);
        if (v == 0) v = 1;
        oplanes_encode(&pos, refshim::config().mode, v, normalize, planes);
        std::lock_guard<std::mutex> lock(refshim::maps_mutex());
        refshim::plane_keys()[planes] = pos.key;
    }
    unsigned int steps_from_null() const override { return pos.game_ply; }
    bool is_chess960() const override { return pos.chess960; }
    std::string fen() const override {
        char buf[256];
        opos_fen(&pos, buf);
        return buf;
    }
    void do_action(Action a) override { opos_do_move(&pos, static_cast<uint32_t>(a)); }
    void undo_action(Action) override {}
    void prepare_action() override {}
    unsigned int number_repetitions() const override { return opos_number_repetitions(&pos); }
    int side_to_move() const override { return pos.stm; }
    Key hash_key() const override { return pos.key; }
    void flip() override {}
    Action uci_to_action(std::string& uci) const override { return static_cast<Action>(opos_uci_to_move(&pos, uci.c_str())); }
    std::string action_to_san(Action a, const std::vector<Action>&, bool, bool) const override {
        return StateConstantsPommerman::action_to_uci(a, pos.chess960);
    }
    TerminalType is_terminal(size_t n_legal, float& custom) const override {
        custom = 0;
        return static_cast<TerminalType>(opos_is_terminal(&pos, static_cast<int>(n_legal)));
    }
    bool gives_check(Action a) const override { return opos_gives_check(&pos, static_cast<uint32_t>(a)); }
    void print(std::ostream& os) const override { os << fen(); }
    Tablebase::WDLScore check_for_tablebase_wdl(Tablebase::ProbeState& result) override {
        result = Tablebase::FAIL;
        return Tablebase::WDLScoreNone;
    }
    void set_auxiliary_outputs(const float*) override {}
    PommermanState* clone() const override {
        PommermanState* s = new PommermanState();
        opos_copy(&s->pos, &pos);
        return s;
    }
    void init(int variant, bool is960) override { opos_set(&pos, opos_start_fen(variant), variant, is960); }
    GamePhase get_phase(unsigned int, GamePhaseDefinition) const override { return 0; }
};

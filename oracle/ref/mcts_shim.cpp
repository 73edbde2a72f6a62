// oracle/ref/mcts_shim.cpp -- TEST INFRASTRUCTURE.  Drives the REFERENCE's own search code -- node.cpp, nodedata.cpp,
// searchthread.cpp, agents/mctsagent.cpp, agents/agent.cpp, evalinfo.cpp, manager/*.cpp, util/blazeutil.h, compiled
// unchanged from /root/reference by oracle/Makefile -- through MCTSAgent::evaluate_board_state
// (agents/mctsagent.cpp:292-337) with Threads = 1, over the State shim (pommermanstate.h), the blaze stand-in
// (blaze/Math.h) and the network stand-in below.  oracle/mcts.c, the restatement every device test is compared with, is
// itself compared with THIS in tests/test_ref_mcts.py: that is what pins the search oracle to the reference's code.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>

#include <blaze/Math.h>  // (the stand-in: pulls the standard headers in before the access hack below)
// Threads = 2 is driven in a FIXED schedule (ref_mcts_run below), which needs the two halves of
// SearchThread::thread_iteration (searchthread.cpp:403-416) separately; they are private members.  Access specifiers
// do not change the classes' layout, so the reference's own translation units (compiled unchanged) link as they are.
#define private public
#define protected public
#include "agents/mctsagent.h"
#undef private
#undef protected
#include "evalinfo.h"
#include "nn/neuralnetapi.h"
#include "stateobj.h"

extern "C" {
#include "../mcts.h"
void ref_seed_node_generator(unsigned long long seed);
}

// ---- the abstract NeuralNetAPI's non-virtual members (nn/neuralnetapi.cpp is not compiled: its constructor scans a
// model directory for ONNX / params files, which has no meaning here)
NeuralNetAPI::NeuralNetAPI(const string& ctx, int deviceID, unsigned int batchSize, const string& modelDirectory, bool enableTensorrt)
    : deviceID(deviceID), batchSize(batchSize), enableTensorrt(enableTensorrt), modelName("shim"), deviceName(ctx),
      modelDir(modelDirectory), nbNNInputValues(0), nbNNAuxiliaryOutputs(0), nbPolicyValues(0), version(0), gamePhase(0) {}
bool NeuralNetAPI::is_policy_map() const { return nnDesign.isPolicyMap; }
string NeuralNetAPI::get_model_name() const { return modelName; }
string NeuralNetAPI::get_device_name() const { return deviceName; }
GamePhase NeuralNetAPI::get_game_phase() const { return gamePhase; }
unsigned int NeuralNetAPI::get_batch_size() const { return batchSize; }
void NeuralNetAPI::validate_neural_network() {}
void NeuralNetAPI::initialize() {}
void NeuralNetAPI::initialize_nn_design() {}
bool NeuralNetAPI::file_exists(const std::string&) { return false; }

typedef void (*ref_net_fn)(void* ctx, const float* planes, const unsigned long long* keys, int n, float* value, float* prob);

namespace {
class ShimNet : public NeuralNetAPI {
   public:
    ShimNet(unsigned batch, ref_net_fn fn, void* ctx) : NeuralNetAPI("cpu", 0, batch, "", false), fn_(fn), ctx_(ctx) {
        nbNNInputValues = StateConstants::NB_VALUES_TOTAL();
        nbPolicyValues = StateConstants::NB_LABELS_POLICY_MAP();
        nbNNAuxiliaryOutputs = 0;
        nnDesign.isPolicyMap = true;
        nnDesign.hasAuxiliaryOutputs = false;
        version = make_version(refshim::config().input_version, 0, 0);
    }
    // TensorrtAPI::predict semantics (nn/tensorrtapi.cpp:195-237): the whole fixed-size batch, prob soft-maxed
    void predict(float* planes, float* value, float* prob, float*) override {
        std::vector<unsigned long long> keys(batchSize, 0);
        int n = 0;  // slots that have been written at least once (the reference evaluates stale slots too and ignores them)
        {
            std::lock_guard<std::mutex> lock(refshim::maps_mutex());
            for (unsigned i = 0; i < batchSize; ++i) {
                auto it = refshim::plane_keys().find(planes + static_cast<size_t>(i) * nbNNInputValues);
                if (it == refshim::plane_keys().end()) break;
                keys[i] = it->second;
                n = static_cast<int>(i) + 1;
            }
        }
        if (fn_ != nullptr) {
            fn_(ctx_, planes, keys.data(), n, value, prob);
        } else {
            for (int i = 0; i < n; ++i) ofake_eval(keys[i], static_cast<int>(nbPolicyValues), value + i, prob + static_cast<size_t>(i) * nbPolicyValues);
        }
    }

   private:
    void init_nn_design() override {}
    void load_model() override {}
    void load_parameters() override {}
    void bind_executor() override {}
    ref_net_fn fn_;
    void* ctx_;
};
}  // namespace

extern "C" {

struct RefResult {  // layout mirrored by tests/test_ref_mcts.py
    int n_moves;
    int no_visit_idx;  // children opened so far: qValues / childNumberVisits hold this many entries
    int best_idx;
    int pv_len;
    float root_value;
    float best_move_q;
    unsigned visit_sum;
    unsigned free_visits;
    unsigned nodes;
    unsigned moves[512];
    unsigned visits[512];
    float q[512];
    float prior[512];
    double policy[512];
    unsigned pv[256];
};

// One MCTSAgent::evaluate_board_state on `fen` (+ uci moves played first), with the settings of the oracle's struct.
int ref_mcts_run(const char* fen, int variant, int is960, const char* const* uci_moves, int n_moves, const OSettings* st,
                 ref_net_fn fn, void* ctx, RefResult* out) {
    opos_init_tables();
    refshim::config().mode = st->mode;
    refshim::config().input_version = st->input_version;
    refshim::index_cache().clear();
    refshim::plane_keys().clear();
    std::ostringstream sink;  // info strings of the engine
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());

    SearchSettings ss;
    ss.multiPV = 1;
    ss.threads = 1;
    ss.batchSize = st->batch_size;
    ss.dirichletEpsilon = st->dirichlet_epsilon;
    ss.dirichletAlpha = st->dirichlet_alpha;
    ss.nodePolicyTemperature = st->node_policy_temperature;
    ss.qValueWeight = st->q_value_weight;
    ss.qVetoDelta = st->q_veto_delta;
    ss.verbose = false;
    ss.epsilonChecksCounter = static_cast<uint_fast8_t>(st->epsilon_checks_counter);  // round(100 / Centi_Epsilon_Checks), 0 = off
    ss.epsilonGreedyCounter = static_cast<uint_fast8_t>(st->epsilon_greedy_counter);  // (crazyara.cpp:748-749)
    ss.useMCGS = false;
    ss.cpuctInit = st->cpuct_init;
    ss.cpuctBase = st->cpuct_base;
    ss.randomMoveFactor = 0;
    ss.allowEarlyStopping = false;
    ss.useNPSTimemanager = false;
    ss.useTablebase = false;
    ss.reuseTree = false;
    ss.mctsSolver = st->mcts_solver != 0;
    ss.searchPlayerMode = MODE_TWO_PLAYER;
    ss.virtualStyle = static_cast<VirtualStyle>(st->virtual_style);
    ss.virtualMixThreshold = st->virtual_mix_threshold;
    PlaySettings ps;
    ps.initTemperature = 0;
    ps.temperatureMoves = 0;
    ps.temperatureDecayFactor = 1;
    ps.quantileClipping = 0;
    SearchLimits limits;
    limits.reset();
    limits.simulations = st->simulations;
    limits.nodes = st->nodes;

    int rc = 0;
    {
        std::vector<std::unique_ptr<NeuralNetAPI>> single;
        single.emplace_back(new ShimNet(1, fn, ctx));
        ss.threads = st->threads == 2 ? 2 : 1;
        std::vector<std::vector<std::unique_ptr<NeuralNetAPI>>> batches(ss.threads);
        for (auto& b : batches) b.emplace_back(new ShimNet(st->batch_size, fn, ctx));  // one net per thread (mctsagent.cpp:54-56)
        MCTSAgent agent(single, batches, &ss, &ps);
        StateObj state;
        state.set(fen != nullptr && fen[0] ? fen : StateConstants::start_fen(variant), is960 != 0, variant);
        for (int i = 0; i < n_moves; ++i) {
            std::string u = uci_moves[i];
            state.do_action(state.uci_to_action(u));
        }
        EvalInfo eval;
        agent.set_search_settings(&state, &limits, &eval);
        ref_seed_node_generator(st->seed);
        srand(static_cast<unsigned>(st->seed));  // the exploration branches draw from the C library's rand()
        eval.start = chrono::steady_clock::now();
        if (st->threads != 2 || st->reserved == 1) {
            // Threads 1, or (reserved == 1: bench.py's CPU arm) the reference's own OS threads, unscheduled
            agent.evaluate_board_state();
        } else {
            // MCTSAgent::evaluate_board_state (agents/mctsagent.cpp:292-337) with run_mcts_search's two OS threads
            // replaced by one of the interleavings they can produce: the fixed schedule of oracle/mcts.h
            //     sel(0) sel(1) | bk(0) sel(0) bk(1) sel(1) | ...
            // sel = the loop test of run_search_thread (searchthread.cpp:418-426) + create_mini_batch,
            // bk  = the rest of thread_iteration (:403-416): predict, set_nn_results_to_child_nodes, the backups.
            ss.threads = 2;
            agent.rootState = unique_ptr<StateObj>(state.clone());
            eval.nodesPreSearch = agent.init_root_node(&state);
            eval.isChess960 = state.is_chess960();
            Node* root = agent.rootNode.get();
            if (root->get_number_child_nodes() == 1) {
                agent.handle_single_move();
            } else if (root->get_number_child_nodes() > 1) {
                if (ss.dirichletEpsilon > 0.009f) {
                    root->apply_dirichlet_noise_to_prior_policy(&ss);
                    root->fully_expand_node();
                }
                if (!root->is_root_node()) root->make_to_root();
                SearchThread* th[2] = {agent.searchThreads[0], agent.searchThreads[1]};
                bool alive[2] = {true, true}, pending[2] = {false, false};
                for (SearchThread* t : th) {
                    t->set_root_node(root);
                    t->set_root_state(agent.rootState.get());
                    t->set_search_limits(&limits);
                    t->set_reached_tablebases(false);
                    t->set_is_running(true);
                    t->reset_stats();
                }
                // the loop test of run_search_thread; nodes_limits_ok / is_root_node_unsolved (searchthread.cpp:326-340) are
                // declared inline but defined in the .cpp, so they cannot be called from here: their two expressions
                auto loop_ok = [&](SearchThread* t) {
                    const Node* r = t->rootNode;
                    const SearchLimits* l = t->searchLimits;
                    const bool limits = (l->nodes == 0 || r->get_node_count() < l->nodes) &&
                                        (l->simulations == 0 || r->get_visits() < l->simulations) &&
                                        (l->nodesLimit == 0 || r->get_node_count() < l->nodesLimit);
                    return t->is_running() && limits && r->get_node_type() == UNSOLVED;
                };
                auto sel = [&](int i) {
                    SearchThread* t = th[i];
                    if (!(alive[i] && loop_ok(t))) {
                        alive[i] = false;
                        return;
                    }
                    t->create_mini_batch();
                    pending[i] = true;
                };
                auto bk = [&](int i) {
                    SearchThread* t = th[i];
                    if (!pending[i]) return;
                    if (t->newNodes->size() != 0) {
                        t->nets[t->select_nn_index()]->predict(t->inputPlanes, t->valueOutputs, t->probOutputs, t->auxiliaryOutputs);
                        t->set_nn_results_to_child_nodes();
                    }
                    t->backup_value_outputs();
                    t->backup_collisions();
                    pending[i] = false;
                };
                sel(0);
                sel(1);
                while (pending[0] || pending[1])
                    for (int i = 0; i < 2; ++i) {
                        bk(i);
                        sel(i);
                    }
                agent.update_stats();
            }
            update_eval_info(eval, root, agent.tbHits, agent.maxDepth, &ss);
        }

        const Node* root = agent.get_root_node();
        memset(out, 0, sizeof(*out));
        out->n_moves = static_cast<int>(root->get_number_child_nodes());
        if (out->n_moves > 512) rc = -1;
        for (int i = 0; i < out->n_moves && rc == 0; ++i) out->moves[i] = static_cast<unsigned>(root->get_action(i));
        if (rc == 0 && root->is_playout_node()) {
            const DynamicVector<uint32_t> nv = root->get_child_number_visits();
            const DynamicVector<float> q = root->get_q_values();
            for (size_t i = 0; i < nv.size(); ++i) out->visits[i] = nv[i];
            for (size_t i = 0; i < q.size(); ++i) out->q[i] = q[i];
            for (int i = 0; i < out->n_moves; ++i) out->prior[i] = const_cast<Node*>(root)->get_policy_prob_small()[i];
            for (size_t i = 0; i < eval.policyProbSmall.size(); ++i) out->policy[i] = eval.policyProbSmall[i];
            out->no_visit_idx = root->get_no_visit_idx();
            out->no_visit_idx = root->get_no_visit_idx();
            out->root_value = root->get_value();
            out->visit_sum = root->get_visits();
            out->free_visits = root->get_free_visits();
            out->nodes = eval.nodes;
            out->best_move_q = eval.bestMoveQ.empty() ? 0.0f : eval.bestMoveQ[0];
            if (!eval.pv.empty()) {
                out->pv_len = static_cast<int>(eval.pv[0].size());
                for (int i = 0; i < out->pv_len && i < 256; ++i) out->pv[i] = static_cast<unsigned>(eval.pv[0][i]);
                for (int i = 0; i < out->n_moves; ++i)
                    if (out->pv_len > 0 && out->moves[i] == out->pv[0]) out->best_idx = i;
            }
        }
    }
    std::cout.rdbuf(old);
    return rc;
}
}

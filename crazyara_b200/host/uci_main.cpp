// Minimal UCI front-end over the C++ host classes (engine/src/uci/crazyara.cpp:76-143 command loop, option names of
// uci/optionsuci.cpp:66-220).  Supported: uci, isready, setoption, ucinewgame, position [startpos|fen] [moves ...],
// go [nodes N] (or the Simulations / Nodes options), root, quit.  Time management (TimeManager / ThreadManager) is not
// part of the hot path: the search always runs to its Simulations / Nodes limit.
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>

#include "ara_host.h"

using namespace crazyara;

namespace {

struct Options {
    std::map<std::string, std::string> kv = {{"UCI_Variant", "crazyhouse"}, {"Model_Path", ""},        {"Batch_Size", "16"},
                                             {"Simulations", "800"},        {"Nodes", "0"},            {"Centi_CPuct_Init", "250"},
                                             {"CPuct_Base", "19652"},       {"Centi_Node_Temperature", "170"},
                                             {"Centi_Dirichlet_Epsilon", "0"}, {"Centi_Dirichlet_Alpha", "20"},
                                             {"Centi_Q_Value_Weight", "100"},  {"Centi_Q_Veto_Delta", "40"},
                                             {"MCTS_Solver", "true"},       {"Virtual_Style", "virtual_mix"},
                                             {"Virtual_Mix_Threshold", "1000"}, {"First_Device_ID", "0"},
                                             {"UCI_Chess960", "false"},     {"Input_Version", "0"},    {"Dirichlet_Seed", "42"}};
    int i(const std::string& k) const { return std::stoi(kv.at(k)); }
    bool b(const std::string& k) const { return kv.at(k) == "true"; }
};

int variant_id(const std::string& v) {
    if (v == "crazyhouse") return 1;
    if (v == "kingofthehill") return 2;
    if (v == "3check" || v == "threecheck") return 3;
    return 0;
}
int mode_of_variant(int variant) { return variant == 0 ? 1 : (variant == 1 ? 0 : 2); }

}  // namespace

int main() {
    Options opt;
    std::unique_ptr<NeuralNetAPI> net;
    std::unique_ptr<MCTSAgent> agent;
    BoardState state;
    EvalInfo info;
    bool ready = false;
    auto variant = [&]() { return variant_id(opt.kv["UCI_Variant"]); };
    auto new_game = [&]() { state.init(variant(), opt.b("UCI_Chess960")); };
    auto prepare = [&]() {  // CrazyAra::is_ready (crazyara.cpp:597): build net + agent from the options
        const int mode = mode_of_variant(variant());
        SearchSettings s(mode);
        s.batch_size = opt.i("Batch_Size");
        s.simulations = static_cast<unsigned>(opt.i("Simulations"));
        s.nodes = static_cast<unsigned>(opt.i("Nodes"));
        s.cpuct_init = opt.i("Centi_CPuct_Init") / 100.0f;
        s.cpuct_base = static_cast<float>(opt.i("CPuct_Base"));
        s.node_policy_temperature = opt.i("Centi_Node_Temperature") / 100.0f;
        s.dirichlet_epsilon = opt.i("Centi_Dirichlet_Epsilon") / 100.0f;
        s.dirichlet_alpha = opt.i("Centi_Dirichlet_Alpha") / 100.0f;
        s.q_value_weight = opt.i("Centi_Q_Value_Weight") / 100.0f;
        s.q_veto_delta = opt.i("Centi_Q_Veto_Delta") / 100.0f;
        s.mcts_solver = opt.b("MCTS_Solver") ? 1 : 0;
        const std::string vs = opt.kv["Virtual_Style"];
        s.virtual_style = vs == "virtual_loss" ? 0 : (vs == "virtual_visit" ? 1 : 3);
        s.virtual_mix_threshold = static_cast<unsigned>(opt.i("Virtual_Mix_Threshold"));
        s.seed = static_cast<unsigned long long>(opt.i("Dirichlet_Seed"));
        if (opt.i("Input_Version") > 0) s.input_version = opt.i("Input_Version");
        agent.reset();
        net.reset();
        if (!opt.kv["Model_Path"].empty())
            net.reset(new NeuralNetAPI("gpu", opt.i("First_Device_ID"), static_cast<unsigned>(s.batch_size), opt.kv["Model_Path"]));
        agent.reset(new MCTSAgent(net.get(), s, opt.i("First_Device_ID"), 0));
        ready = true;
    };
    new_game();
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream ss(line);
        std::string cmd;
        ss >> cmd;
        try {
            if (cmd == "uci") {
                std::cout << "id name CrazyAra-B200\nid author crazyara_b200 (hot path of QueensGambit/CrazyAra on sm_100a)\n";
                for (const auto& kv : opt.kv) std::cout << "option name " << kv.first << " type string default " << kv.second << "\n";
                std::cout << "uciok" << std::endl;
            } else if (cmd == "isready") {
                if (!ready) prepare();
                std::cout << "readyok" << std::endl;
            } else if (cmd == "setoption") {
                std::string tok, name, value;
                ss >> tok;  // "name"
                while (ss >> tok && tok != "value") name += (name.empty() ? "" : " ") + tok;
                while (ss >> tok) value += (value.empty() ? "" : " ") + tok;
                if (opt.kv.count(name) == 0) {
                    std::cout << "info string unknown option " << name << std::endl;
                } else {
                    opt.kv[name] = value;
                    ready = false;
                    if (name == "UCI_Variant" || name == "UCI_Chess960") new_game();
                }
            } else if (cmd == "ucinewgame") {
                new_game();
            } else if (cmd == "position") {
                std::string tok, fen;
                ss >> tok;
                if (tok == "startpos") {
                    new_game();
                    ss >> tok;  // optional "moves"
                } else if (tok == "fen") {
                    while (ss >> tok && tok != "moves") fen += (fen.empty() ? "" : " ") + tok;
                    state.set(fen, opt.b("UCI_Chess960"), variant());
                }
                while (ss >> tok) {
                    const Action a = state.uci_to_action(tok);
                    if (a == 0) {
                        std::cout << "info string illegal move " << tok << std::endl;
                        break;
                    }
                    state.do_action(a);
                }
            } else if (cmd == "go") {
                std::string tok;
                while (ss >> tok)
                    if (tok == "nodes") {
                        ss >> tok;
                        opt.kv["Nodes"] = tok;
                        ready = false;
                    }
                if (!ready) prepare();
                agent->evaluate_board_state(state, info);
                std::cout << "info depth " << info.depth << " nodes " << info.nodes << " nps " << info.calculate_nps() << " score cp "
                          << info.centipawns << " time " << static_cast<long>(info.elapsedMs) << " pv";
                for (Action a : info.pv) std::cout << " " << state.action_to_uci(a);
                std::cout << "\nbestmove " << (info.bestMove ? state.action_to_uci(info.bestMove) : std::string("(none)")) << std::endl;
            } else if (cmd == "root") {  // Node::print_node_statistics (node.cpp:1248-1301): the parity dump format
                std::cout << "  #  | Move  |    Visits    |  Policy   |  Q-values  |  CP   \n";
                std::cout << std::fixed << std::setprecision(7);
                for (size_t i = 0; i < info.legalMoves.size(); ++i)
                    std::cout << " " << std::setw(3) << std::setfill('0') << i << std::setfill(' ') << " | " << std::setw(5)
                              << state.action_to_uci(info.legalMoves[i]) << " | " << std::setw(12) << info.childNumberVisits[i] << " | "
                              << std::setw(9) << info.priors[i] << " | " << std::setw(10) << info.qValues[i] << " | " << std::setw(5)
                              << value_to_centipawn(info.qValues[i], 1.2f) << "\n";
                std::cout << "value:\t" << info.rootValue << "\nVisits:\t" << info.nodes << std::endl;
            } else if (cmd == "quit") {
                break;
            }
        } catch (const std::exception& e) {
            std::cout << "info string error: " << e.what() << std::endl;
        }
    }
    return 0;
}

"""The parts of the drop-in boundary beyond predict / go: stepping the search like
SearchThread::thread_iteration, reading nodes like Node's getters, the legal-move prior seam (fill_nn_results), pinned
caller buffers -- and the C++ host classes SearchThread / Node over them."""
import os
import subprocess

import numpy as np
import pytest

from oracle import search as osr
from oracle.chess import Position

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _agent(sims=400, batch=8, **kw):
    from crazyara_b200.engine import BoardState, MCTSAgent, default_settings
    st = default_settings("crazyhouse", batch_size=batch, simulations=sims, **kw)
    return MCTSAgent(None, st, 0, 1), BoardState().set("", False, 1), st


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 2])
def test_stepping_equals_go_and_the_oracle(threads):
    """ara_search_begin + ara_search_step until it returns 0 == ara_search_go == the oracle; results readable between steps."""
    from tests.test_search_hostemu import assert_same_search
    agent, state, st = _agent(threads=threads)
    whole = agent.evaluate_board_state(state)
    agent.close()
    agent, state, _ = _agent(threads=threads)
    agent.begin(state)
    seen, steps = [], 0
    while True:
        running = agent.thread_iteration(1)
        steps += 1
        seen.append(agent.result(0)["visit_sum"])
        if not running:
            break
    stepped = agent.result(0)
    assert steps >= 400 // (8 * threads) and seen[0] < seen[-1]
    if threads == 1:   # (with two threads a step can end with fewer visits: the other thread's collisions were reverted)
        assert seen == sorted(seen)
    assert agent.thread_iteration(3) == 0 and agent.result(0)["visit_sum"] == stepped["visit_sum"]   # further steps: no-ops
    assert_same_search(whole, stepped)
    ost = osr.default_settings("crazyhouse", batch_size=8, simulations=400, threads=threads)
    S = osr.Search(ost)
    assert_same_search(S.run(Position(variant="crazyhouse"), osr.fake_net(S.n_labels), with_keys=True, threads=threads), stepped)
    # Node view of the root and of its most visited child
    root = agent.node(-1)
    assert root["moves"] == stepped["moves"] and np.array_equal(root["child_visits"], stepped["visits"])
    assert np.array_equal(root["q"][:root["no_visit_idx"]], stepped["q"][:root["no_visit_idx"]])
    assert np.array_equal(root["prior"], stepped["prior"]) and root["visits"] == stepped["visit_sum"]
    assert root["parent"] == -1 and root["is_playout_node"] and abs(root["value"] - stepped["root_value"]) < 1e-7
    best = int(np.argmax(root["child_visits"]))
    child = agent.node(root["child"][best])
    assert child["parent"] == root["node_id"] and child["has_nn_results"]
    # a child's own visit count is the edge count of its parent (minus nothing: no virtual visits are pending now)
    assert child["visits"] + 1 == root["child_visits"][best] or child["visits"] == root["child_visits"][best] - 1 or child["is_terminal"]
    assert sum(c >= 0 for c in root["child"]) == int((root["child_visits"] > 0).sum())
    with pytest.raises(Exception):
        agent.node(10 ** 8)
    agent.close()


@pytest.mark.gpu
def test_predict_priors_equals_predict_gathered_and_pinned_buffers(tmp_path):
    from crazyara_b200.engine import BoardState, legal_moves_gpu
    from crazyara_b200.nn import NeuralNetAPI
    from crazyara_b200.weights import export_blob
    from oracle import net as onet
    arch = onet.arch_risev2(34, 81)
    blob = export_blob(onet.make_state_dict(arch, 0), arch, str(tmp_path / "n.arab"), input_version=10)
    net = NeuralNetAPI("gpu", 0, 8, blob)
    states = [BoardState().set("", False, 1).do_uci(*m) for m in ([], ["e2e4"], ["e2e4", "e7e5"], ["g1f3", "d7d5", "d2d4"])]
    n = len(states)
    x = NeuralNetAPI.pinned_array((8, 34, 8, 8))          # pinned caller buffers (neuralnetapiuser.cpp:52-59)
    v = NeuralNetAPI.pinned_array((8,))
    p = NeuralNetAPI.pinned_array((8, 81 * 64))
    for i, s in enumerate(states):
        x[i] = s.get_state_planes(True, 0, 1)
    net.predict(x, v, p, None, n=n)
    moves, _, pidx = legal_moves_gpu([s.board() for s in states])
    idx = np.full((8, 128), -1, np.int32)
    cnt = np.zeros(8, np.int32)
    for i, row in enumerate(pidx):
        idx[i, :len(row)] = row
        cnt[i] = len(row)
    v2 = np.zeros(8, np.float32)
    pri = np.full((8, 128), np.nan, np.float32)
    net.predict_priors(x, idx, cnt, v2, pri, None, n=n)
    assert np.array_equal(v2[:n], v[:n])
    for i in range(n):
        assert cnt[i] == len(moves[i]) > 0
        assert np.array_equal(pri[i, :cnt[i]], p[i][idx[i, :cnt[i]]])       # exactly the entries predict returns
    net.close()


@pytest.mark.gpu
def test_cpp_host_searchthread_and_node(tmp_path):
    """crazyara_b200/host/ara_host.h: SearchThread::thread_iteration + Node getters, compiled and run against the library."""
    src = tmp_path / "t.cpp"
    src.write_text(r"""
#include <cstdio>
#include "ara_host.h"
using namespace crazyara;
int main() {
    SearchSettings st(0);
    st.batch_size = 8; st.simulations = 300;
    BoardState pos; pos.set("", false, 1);
    SearchThread th(nullptr, st);
    th.set_root_state(pos);
    int iters = 0;
    while (th.thread_iteration()) ++iters;
    Node root = th.get_root_node();
    unsigned sum = 0, best = 0; size_t bi = 0;
    auto n = root.get_child_number_visits();
    for (size_t i = 0; i < n.size(); ++i) { sum += n[i]; if (n[i] > best) best = n[i], bi = i; }
    auto child = root.get_child_node(bi);
    MCTSAgent agent(nullptr, st);
    EvalInfo ev; agent.evaluate_board_state(pos, ev);
    std::printf("%d %u %u %zu %d %u %zu %d\n", iters, root.get_visits(), sum, root.get_number_child_nodes(),
                child ? (int)child->has_nn_results() : -1, (unsigned)ev.nodes, ev.legalMoves.size(),
                (int)(ev.childNumberVisits[bi] == best && ev.legalMoves[bi] == root.get_action(bi)));
    return 0;
}
""")
    exe = tmp_path / "t"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "crazyara_b200", "host"),
                    str(src), "-o", str(exe), "-L" + os.path.join(ROOT, "crazyara_b200"), "-lara_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "crazyara_b200"), "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    iters, visits, child_sum, n_moves, child_nn, nodes, n_legal, same = (int(x) for x in out)
    assert iters >= 300 // 8 - 1 and visits >= 300 and child_sum == visits and n_moves == 20 == n_legal
    assert child_nn == 1 and nodes == visits and same == 1

"""The GPU search code (crazyara_b200/csrc/search_dev.cuh) run as a 1-lane host emulation against the search
oracle (oracle/mcts.c) with the hash-derived fake backend: visit counts, Q values, priors, posterior, root value,
best move, node counters must agree BIT-EXACTLY (the GPU run of the same comparison is tests/test_search_gpu.py)."""
import numpy as np
import pytest

from oracle import search as osr
from oracle.chess import Position
from tests.hostemu import HeSearch, HeState

CASES = [
    # variant, vid, mode, fen, is960, moves-to-play-first, batch, sims, extra settings
    ("crazyhouse", 1, "crazyhouse", None, False, [], 8, 400, {}),
    ("crazyhouse", 1, "crazyhouse", None, False, [], 64, 1600, {}),
    ("chess", 0, "chess", None, False, ["e2e4", "c7c5", "g1f3"], 64, 1600, {}),
    ("chess", 0, "chess", None, False, [], 1, 100, {}),
    ("chess", 0, "chess", "bqnb1rkr/pp3ppp/3ppn2/2p5/5P2/P2P4/NPP1P1PP/BQ1BNRKR w HFhf - 2 9", True, [], 8, 300, {}),
    ("kingofthehill", 2, "lichess", None, False, ["e2e4", "e7e5", "e1e2", "e8e7", "e2e3", "e7e6"], 16, 600, {}),
    ("3check", 3, "lichess", None, False, ["e2e4", "f7f6", "d1h5", "g7g6"], 16, 600, {}),
    # mate-in-few positions: exercise the MCTS solver and terminal (free) backups
    ("chess", 0, "chess", "6k1/5ppp/8/8/8/8/8/R3K2R w KQ - 0 1", False, [], 8, 400, {}),
    ("chess", 0, "chess", "7k/5Q2/6K1/8/8/8/8/8 w - - 0 1", False, [], 4, 200, {}),
    ("crazyhouse", 1, "crazyhouse", "4R2b/1N3rkb/1p2P1pp/p2P4/2P1P3/8/PP4Q1/3R3K[QRBBNNNPPPPpp] w - - 2 53", False, [], 8, 300, {}),
    # virtual loss style, solver off, nodes limit
    ("crazyhouse", 1, "crazyhouse", None, False, ["e2e4"], 16, 500, dict(virtual_style=0)),
    ("crazyhouse", 1, "crazyhouse", None, False, ["e2e4"], 16, 0, dict(nodes=300, mcts_solver=0, virtual_mix_threshold=20)),
    ("chess", 0, "chess", None, False, ["d2d4"], 32, 800, dict(q_value_weight=0.0, q_veto_delta=0.0)),
    # the UCI default node temperature 1.7 (glibc powf + sequential normalising sum on both sides), other temperatures,
    # and Dirichlet noise at the root (libstdc++ gamma sampler over glibc logf / powf): still bit-exact
    ("crazyhouse", 1, "crazyhouse", None, False, [], 64, 3200, dict(node_policy_temperature=1.7)),
    ("crazyhouse", 1, "crazyhouse", None, False, [], 8, 800, dict(node_policy_temperature=1.7)),
    ("chess", 0, "chess", None, False, [], 64, 1600, dict(node_policy_temperature=1.7)),
    ("kingofthehill", 2, "lichess", None, False, ["e2e4", "e7e5"], 128, 800, dict(node_policy_temperature=1.7)),
    ("3check", 3, "lichess", None, False, ["e2e4", "e7e5"], 128, 800, dict(node_policy_temperature=1.7)),
    ("crazyhouse", 1, "crazyhouse", None, False, ["e2e4", "e7e5"], 16, 500, dict(node_policy_temperature=0.7)),
    ("crazyhouse", 1, "crazyhouse", None, False, [], 8, 400, dict(dirichlet_epsilon=0.25, dirichlet_alpha=0.3, seed=7)),
    ("chess", 0, "chess", None, False, ["e2e4"], 8, 0, dict(nodes=800, dirichlet_epsilon=0.25, dirichlet_alpha=0.3, seed=12345,
                                                          node_policy_temperature=1.7)),
    ("crazyhouse", 1, "crazyhouse", None, False, [], 16, 400, dict(dirichlet_epsilon=0.25, dirichlet_alpha=1.5, seed=99)),
]


def case_settings(mode, batch, sims, extra):
    """Settings of a CASES row: node temperature 1 unless the row sets it."""
    kw = dict(node_policy_temperature=1.0)
    kw.update(extra)
    return osr.default_settings(mode, batch_size=batch, simulations=sims, **kw)


def _run_both(variant, vid, mode, fen, is960, premoves, batch, sims, extra):
    st = case_settings(mode, batch, sims, extra)
    pos = Position(fen, variant, is960)
    he = HeState(pos.fen(), vid, is960)
    for u in premoves:
        pos.push_uci(u)
        he.do_move(he.move_from_uci(u))
    S = osr.Search(st)
    ro = S.run(pos, osr.fake_net(S.n_labels), with_keys=True)
    H = HeSearch(st)
    rh = H.run(he, osr.fake_net(H.n_labels), with_keys=True)
    return ro, rh


def assert_same_search(ro, rh):
    assert rh["error"] == 0
    assert ro["moves"] == rh["moves"]
    assert np.array_equal(ro["visits"], rh["visits"])
    assert np.array_equal(ro["q"], rh["q"])
    assert np.array_equal(ro["prior"], rh["prior"])
    assert np.array_equal(ro["policy"], rh["policy"])
    for k in ("visit_sum", "free_visits", "nodes", "best_idx", "node_type", "iterations", "evals", "tree_nodes",
              "sum_select_k", "sum_depth", "pv_len"):
        assert ro[k] == rh[k], k
    assert ro["root_value"] == rh["root_value"] and ro["best_move_q"] == rh["best_move_q"]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-b{c[6]}-s{c[7]}-{i}" for i, c in enumerate(CASES)])
def test_hostemu_search_equals_oracle(case):
    ro, rh = _run_both(*case)
    assert ro["visit_sum"] > 0
    assert_same_search(ro, rh)


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-b{c[6]}-s{c[7]}-{i}" for i, c in enumerate(CASES)])
def test_hostemu_two_thread_search_equals_oracle(case):
    """Threads = 2: two logical search threads per tree in the fixed schedule of oracle/mcts.h (one thread's batch is at
    the network while the other selects); the device code with two batch slots against the oracle's two threads."""
    variant, vid, mode, fen, is960, premoves, batch, sims, extra = case
    st = case_settings(mode, batch, sims, dict(extra, threads=2))
    pos = Position(fen, variant, is960)
    he = HeState(pos.fen(), vid, is960)
    for u in premoves:
        pos.push_uci(u)
        he.do_move(he.move_from_uci(u))
    S = osr.Search(st)
    ro = S.run(pos, osr.fake_net(S.n_labels), with_keys=True, threads=2)
    rh = HeSearch(st).run(he, osr.fake_net(S.n_labels), with_keys=True, threads=2)
    assert ro["visit_sum"] > 0
    assert_same_search(ro, rh)


EPS = dict(epsilon_greedy_counter=20, epsilon_checks_counter=100)  # UCI defaults Centi_Epsilon_Greedy 5, Centi_Epsilon_Checks 1


@pytest.mark.parametrize("threads", [1, 2])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-b{c[6]}-s{c[7]}-{i}" for i, c in enumerate(CASES)])
def test_hostemu_epsilon_exploration_equals_oracle(case, threads):
    """The reference's default exploration (searchthread.cpp:124-185, :451-473): random playouts / unexplored checks below
    a randomly deep node of the main line, drawn from glibc's rand() -- restated in the device code (select_kernel<true>)
    and in the oracle, which tests/test_ref_mcts.py pins to the compiled reference."""
    variant, vid, mode, fen, is960, premoves, batch, sims, extra = case
    st = case_settings(mode, batch, sims, dict(extra, threads=threads, **EPS))
    pos = Position(fen, variant, is960)
    he = HeState(pos.fen(), vid, is960)
    for u in premoves:
        pos.push_uci(u)
        he.do_move(he.move_from_uci(u))
    S = osr.Search(st)
    ro = S.run(pos, osr.fake_net(S.n_labels), with_keys=True, threads=threads)
    rh = HeSearch(st).run(he, osr.fake_net(S.n_labels), with_keys=True, threads=threads)
    assert ro["visit_sum"] > 0
    assert_same_search(ro, rh)


def test_fake_backends_identical():
    import ctypes
    from oracle.search import _lib
    from tests.hostemu import lib as helib
    for key in (0, 1, 0xDEADBEEFCAFEF00D, 2**64 - 1, 123456789):
        vo, po = np.zeros(1, np.float32), np.zeros(5184, np.float32)
        vh, ph = np.zeros(1, np.float32), np.zeros(5184, np.float32)
        _lib().ofake_eval(key, 5184, vo.ctypes.data, po.ctypes.data)
        HeSearch(osr.default_settings("crazyhouse"))  # sets argtypes
        helib().he_fake_eval(key, 5184, vh.ctypes.data, ph.ctypes.data)
        assert vo[0] == vh[0] and np.array_equal(po, ph)


def test_oracle_search_reproduces_its_recorded_results():
    """tests/golden/search_oracle.json: the oracle's own results for the cases above, bit patterns included.  A change of
    the oracle that moves any of them has to be made visibly (by regenerating the file)."""
    import json
    import os
    from tests.golden.gen_search_golden import run_case
    recorded = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "search_oracle.json")))
    assert len(recorded) == len(CASES)
    for rec, case in zip(recorded, CASES):
        assert rec["case"][:8] == [list(x) if isinstance(x, (list, tuple)) else x for x in case[:8]] and rec["case"][8] == case[8]
        assert run_case(case) == rec["result"], case[:8]

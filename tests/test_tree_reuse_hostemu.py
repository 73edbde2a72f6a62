"""Tree reuse between moves (MCTSAgent::apply_move_to_tree / init_root_node, agents/mctsagent.cpp:113-160, 230-247):
the device search code (1-lane host emulation) against the oracle over a sequence of searched and played moves.
Every search of the sequence must agree bit-exactly, and the kept subtree must really be used."""
import numpy as np
import pytest

from oracle import search as osr
from oracle.chess import Position
from tests.hostemu import HeSearch, HeState
from tests.test_search_hostemu import assert_same_search

CASES = [("crazyhouse", 1, "crazyhouse", 8, 300, {}), ("chess", 0, "chess", 16, 400, {}),
         ("crazyhouse", 1, "crazyhouse", 8, 0, dict(nodes=250, dirichlet_epsilon=0.25, dirichlet_alpha=0.3)),
         ("3check", 3, "lichess", 4, 200, {})]


@pytest.mark.parametrize("variant,vid,mode,batch,sims,extra", CASES)
def test_reused_tree_equals_oracle(variant, vid, mode, batch, sims, extra):
    st = osr.default_settings(mode, batch_size=batch, simulations=sims, node_policy_temperature=1.0, **extra)
    pos = Position(None, variant, False)
    he = HeState(pos.fen(), vid, False)
    S = osr.Search(st)
    H = HeSearch(st, max_nodes=1 << 16)
    reused = 0
    for ply in range(6):
        ro = S.run(pos, osr.fake_net(S.n_labels), with_keys=True)
        rh = H.run(he, osr.fake_net(H.n_labels), with_keys=True)
        assert S.reused == H.reused == (ply > 0)
        reused += int(S.reused)
        assert_same_search(ro, rh)
        assert rh["nodes_pre_search"] == S.nodes_pre_search
        if ply > 0:
            assert ro["evals"] < ro["nodes"] and S.nodes_pre_search > 0  # the kept visits count towards the budget
        # play the second most visited move now and then, so that the kept subtree is not always the biggest one
        order = np.argsort(-ro["visits"].astype(np.int64), kind="stable")
        pick = int(order[1 if (ply % 3 == 2 and len(order) > 1 and ro["visits"][order[1]] > 0) else 0])
        uci = ro["moves"][pick]
        assert S.apply_move(pos.move_from_uci(uci))
        H.apply_move(he.move_from_uci(uci))
        pos.push_uci(uci)
        he.do_move(he.move_from_uci(uci))
    assert reused == 5


def test_a_different_position_starts_a_new_tree():
    st = osr.default_settings("chess", batch_size=8, simulations=200, node_policy_temperature=1.0)
    pos = Position(None, "chess", False)
    he = HeState(pos.fen(), 0, False)
    S, H = osr.Search(st), HeSearch(st)
    r0 = S.run(pos, osr.fake_net(S.n_labels), with_keys=True)
    H.run(he, osr.fake_net(H.n_labels), with_keys=True)
    uci = r0["best_move"]
    S.apply_move(pos.move_from_uci(uci))
    H.apply_move(he.move_from_uci(uci))
    # ... but the next search is on another position: nothing may be reused
    other = [m for m in r0["moves"] if m != uci][0]
    pos.push_uci(other)
    he.do_move(he.move_from_uci(other))
    ro = S.run(pos, osr.fake_net(S.n_labels), with_keys=True)
    rh = H.run(he, osr.fake_net(H.n_labels), with_keys=True)
    assert not S.reused and not H.reused
    assert_same_search(ro, rh)

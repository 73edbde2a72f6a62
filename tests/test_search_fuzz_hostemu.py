"""Randomised differential test of the device search code (1-lane host emulation) against the search oracle: positions
reached by random legal play in every supported variant, random search settings.  Everything must agree bit-exactly
(tests/test_search_hostemu.py holds the hand-picked cases; this one hunts for the rare paths: repetitions on the path,
terminal leaves, fast-path / exact-path select decisions, prepared and in-line expansions, collisions)."""
import numpy as np
import pytest

from oracle import search as osr
from oracle.chess import Position
from tests.hostemu import HeSearch, HeState
from tests.test_search_hostemu import assert_same_search

VARIANTS = [("crazyhouse", 1, "crazyhouse"), ("chess", 0, "chess"), ("kingofthehill", 2, "lichess"), ("3check", 3, "lichess")]


def _random_case(seed):
    rng = np.random.default_rng(seed)
    variant, vid, mode = VARIANTS[seed % len(VARIANTS)]
    pos = Position(None, variant, False)
    he = HeState(pos.fen(), vid, False)
    plies = int(rng.integers(0, 40))
    played = []
    for _ in range(plies):
        moves = pos.legal_uci()
        if not moves or pos.terminal(len(moves)) != 4:  # 4 = TERMINAL_NONE
            break
        u = moves[int(rng.integers(0, len(moves)))]
        nxt = pos.clone().push_uci(u)
        nm = nxt.legal_uci()
        if not nm or nxt.terminal(len(nm)) != 4:  # keep a searchable root
            continue
        pos.push_uci(u)
        he.do_move(he.move_from_uci(u))
        played.append(u)
    batch = int(rng.choice([1, 4, 8, 16, 32]))
    sims = int(rng.choice([64, 150, 300, 500]))
    extra = {}
    if rng.random() < 0.3:
        extra["virtual_style"] = int(rng.choice([0, 1, 3]))
    if rng.random() < 0.3:
        extra["virtual_mix_threshold"] = int(rng.choice([5, 30, 1000]))
    if rng.random() < 0.2:
        extra["mcts_solver"] = 0
    if rng.random() < 0.2:
        extra["nodes"] = int(sims // 2)
    # node temperature: 1 (no renormalisation), the UCI default 1.7 and others -- glibc powf restated on the device and a
    # sequential normalising sum in policy-index order on both sides keep T != 1 inside the bit-exact contract; the same
    # for Dirichlet noise at the root (libstdc++ gamma sampler over glibc logf / powf)
    temp = float(rng.choice([1.0, 1.7, 1.7, 1.3, 0.8, 2.5]))
    if rng.random() < 0.3:
        extra["dirichlet_epsilon"] = float(rng.choice([0.25, 0.1]))
        extra["dirichlet_alpha"] = float(rng.choice([0.2, 0.3, 0.6, 1.0, 2.0]))
        extra["seed"] = int(rng.integers(1, 2**31 - 2))
    st = osr.default_settings(mode, batch_size=batch, simulations=sims, node_policy_temperature=temp, **extra)
    return pos, he, st, (vid, played)


@pytest.mark.parametrize("seed", range(96))
def test_random_position_and_settings(seed):
    pos, he, st, _ = _random_case(seed)
    S = osr.Search(st)
    ro = S.run(pos, osr.fake_net(S.n_labels), with_keys=True)
    H = HeSearch(st)
    rh = H.run(he, osr.fake_net(H.n_labels), with_keys=True)
    assert ro["visit_sum"] > 0
    assert_same_search(ro, rh)

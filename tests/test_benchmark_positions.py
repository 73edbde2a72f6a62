"""The reference's tactical benchmark suite (engine/tests/benchmarkpositions.cpp:28-57) as a pin of the rules code: all
fifteen crazyhouse FENs (both pocket spellings: a ninth rank `/NQp` and brackets `[QNbpp]`) must parse, and the blunder
and the alternative the reference lists for each must be legal moves there -- in the host rules (the C-ABI state), in the
device rules (1-lane host emulation) and in the oracle."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _positions():
    src = open(os.path.join(ROOT, "crazyara_b200", "host", "benchmark_positions.h")).read()
    return re.findall(r'\{"([^"]+)", "([^"]+)", "([^"]+)"\}', src)


POSITIONS = _positions()


def test_suite_is_complete():
    assert len(POSITIONS) == 15 and all(" w " in f or " b " in f for f, _, _ in POSITIONS)


@pytest.mark.parametrize("fen,blunder,alternative", POSITIONS)
def test_listed_moves_are_legal_everywhere(fen, blunder, alternative):
    from crazyara_b200.engine import BoardState
    from oracle.chess import Position
    from tests.hostemu import HeState
    host = BoardState().set(fen, False, 1)
    host_moves = sorted(host.action_to_uci(a) for a in host.legal_actions())
    dev = HeState(fen, 1, False)
    dev_moves = sorted(dev.uci(m) for m in dev.legal_moves())
    orc = Position(fen, "crazyhouse", False)
    orc_moves = sorted(orc.uci(m) for m in orc.legal_moves())
    assert host_moves == dev_moves == orc_moves
    assert blunder in host_moves and alternative in host_moves
    # both spellings of the pocket describe the same position
    assert host.fen() == dev.fen() and "[" in host.fen()
    # and the moves can be played and spelled
    for mv in (blunder, alternative):
        st = BoardState().set(fen, False, 1)
        assert st.action_to_san(mv)
        st.do_uci(mv)
        assert st.side_to_move() != host.side_to_move()

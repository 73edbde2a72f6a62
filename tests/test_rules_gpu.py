"""Device move generator, terminal rules, policy-map indices and the plane-encode kernel on the GPU, through the
C-ABI with host buffers, against the CPU oracle: bit-exact on seeded random playouts and on the reference's
PlaneStatistics literals (engine/tests/tests.cpp)."""
import ctypes
import random

import numpy as np
import pytest

from oracle.chess import Position, lib as olib, planes as oplanes

VARIANTS = [("chess", 0), ("crazyhouse", 1), ("kingofthehill", 2), ("3check", 3)]


def _collect_positions(name, vid, rnd, games, max_plies):
    from crazyara_b200.engine import BoardState
    out = []
    for _ in range(games):
        pos = Position(variant=name)
        st = BoardState().set("", False, vid)
        for _ in range(max_plies):
            moves = pos.legal_uci()
            out.append((pos.clone(), st.clone()))
            if pos.terminal(len(moves)) != 4 or not moves:
                break
            u = rnd.choice(sorted(moves))
            pos.push_uci(u)
            st.do_uci(u)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name,vid", VARIANTS)
def test_gpu_movegen_terminal_policy_planes_match_oracle(name, vid):
    from crazyara_b200.engine import encode_planes, legal_moves_gpu, move_to_uci
    L = olib()
    L.opolicy_move_index.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
    rnd = random.Random(99 + vid)
    positions = _collect_positions(name, vid, rnd, games=10, max_plies=160)
    boards = [st.board() for _, st in positions]
    moves, term, pidx = legal_moves_gpu(boards)
    policy_mode = 0 if vid == 1 else (1 if vid == 0 else 2)
    for (pos, st), mv, tt, pi in zip(positions, moves, term, pidx):
        assert st.fen() == pos.fen()
        o = {pos.uci(m): L.opolicy_move_index(pos._buf, m, policy_mode, 1) for m in pos.legal_moves()}
        g = {move_to_uci(m, False): p for m, p in zip(mv, pi)}
        assert o == g, pos.fen()
        assert tt == pos.terminal(len(o)), pos.fen()
    mode_versions = {0: [(1, 1), (1, 3)], 1: [(0, 1), (0, 2), (0, 3)], 2: [(2, 1), (2, 3)], 3: [(2, 1), (2, 3)]}[vid]
    for mode, version in mode_versions:
        for norm in (False, True):
            got = encode_planes(boards, mode, version, norm)
            for i, (pos, _) in enumerate(positions):
                assert np.array_equal(got[i], oplanes(pos, mode, version, norm)), (pos.fen(), mode, version, norm)


@pytest.mark.gpu
def test_gpu_planes_reference_goldens():
    """engine/tests/tests.cpp:1493-1527 (crazyhouse V1) and :341-365 (chess v3) through the CUDA encoder."""
    from crazyara_b200.engine import BoardState
    st = BoardState().set("5r2/ppp2pkp/3p4/2bP4/2Pnp1N1/3P2pP/PP2n1P1/R2Q1R1K[PBRQnbb] w - - 0 28", False, 1)
    st.do_uci("Q@f6", "g7g8", "R@h8")
    assert st.fen() == "5rkR/ppp2p1p/3p1Q2/2bP4/2Pnp1N1/3P2pP/PP2n1P1/R2Q1R1K[BPbbn] b - - 3 29"
    p = st.get_state_planes(False, 0, 1).reshape(-1).astype(np.float64)
    assert p.sum() == 2395 and p.max() == 29 and (np.arange(p.size) * p).sum() == 4170903 and int(p.argmax()) == 1792
    st = BoardState().set("", False, 0)
    p = st.get_state_planes(False, 1, 3).reshape(-1).astype(np.float64)
    assert p.sum() == 1312 and p.max() == 8 and (np.arange(p.size) * p).sum() == 3430384 and int(p.argmax()) == 3008
    p = st.get_state_planes(True, 1, 3).reshape(-1).astype(np.float64)
    assert p.sum() == 472 and (np.arange(p.size) * p).sum() == 819860
    st = BoardState().set("b1qnrnkr/p2ppppp/1p6/2p1b3/2P5/4N1P1/PP1PPP1P/BBQNRK1R b he - 1 4", True, 0)
    p = st.get_state_planes(False, 1, 3).reshape(-1).astype(np.float64)
    assert p.sum() == 1312 and (np.arange(p.size) * p).sum() == 3512322

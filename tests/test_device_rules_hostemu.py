"""The DEVICE rules / planes / policy-index code (crazyara_b200/csrc/chess_dev.cuh, planes_dev.cuh), compiled for
the host as a 1-lane warp, against the CPU oracle on seeded random playouts: bit-exact move sets, FENs, Zobrist
keys, repetition / terminal verdicts, policy-map indices and input planes.  (The same comparisons run on the GPU
through the C-ABI in tests/test_rules_gpu.py.)"""
import ctypes
import random

import numpy as np
import pytest

from oracle.chess import Position, lib as olib, planes as oplanes
from tests.hostemu import HeState

VARIANTS = [("chess", 0, False), ("crazyhouse", 1, False), ("kingofthehill", 2, False), ("3check", 3, False)]
FENS_960 = ["bqnb1rkr/pp3ppp/3ppn2/2p5/5P2/P2P4/NPP1P1PP/BQ1BNRKR w HFhf - 2 9",
            "bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BN1RK1RQ w GDgd - 0 1",
            "nrbbqnkr/pppppppp/8/8/8/8/PPPPPPPP/NR4KR w HBhb - 0 1"]


def _playout(pos, he, rnd, mode_versions, max_plies, check_every=1):
    L = olib()
    L.opolicy_move_index.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
    plies = 0
    while plies < max_plies:
        omoves = pos.legal_moves()
        hmoves = he.legal_moves()
        ouci = sorted(pos.uci(m) for m in omoves)
        huci = sorted(he.uci(m) for m in hmoves)
        assert ouci == huci, pos.fen()
        assert he.fen() == pos.fen()
        assert he.key() == pos.key() == he.key_scratch(), pos.fen()
        assert he.terminal() == pos.terminal(len(omoves)), pos.fen()
        assert (he.repetition() != 0) == (pos.number_repetitions() != 0)
        if plies % check_every == 0:
            policy_mode = 0 if pos.variant == 1 else (1 if pos.variant == 0 else 2)
            o_idx = {pos.uci(m): L.opolicy_move_index(pos._buf, m, policy_mode, 1) for m in omoves}
            h_idx = {he.uci(m): he.policy_index(m) for m in hmoves}
            assert o_idx == h_idx, pos.fen()
            for mode, version in mode_versions:
                for norm in (False, True):
                    a = oplanes(pos, mode, version, norm)
                    b = he.planes(mode, version, norm)
                    assert a.shape == b.shape and np.array_equal(a, b), (pos.fen(), mode, version, norm)
        if pos.terminal(len(omoves)) != 4 or not omoves:
            break
        u = rnd.choice(ouci)
        pos.push_uci(u)
        he.do_move(he.move_from_uci(u))
        plies += 1
    return plies


@pytest.mark.parametrize("name,vid,is960", VARIANTS)
def test_random_playouts_match_oracle(name, vid, is960):
    rnd = random.Random(1234 + vid)
    mv = {0: [(1, 1), (1, 3)], 1: [(0, 1), (0, 2), (0, 3), (2, 1), (2, 3)], 2: [(2, 1), (2, 3)], 3: [(2, 1), (2, 3)]}[vid]
    total = 0
    for game in range(12):
        pos = Position(variant=name)
        he = HeState(pos.fen(), vid, is960)
        total += _playout(pos, he, rnd, mv, 200, check_every=3)
    assert total > 400


def test_chess960_playouts_match_oracle():
    rnd = random.Random(77)
    for fen in FENS_960:
        for game in range(4):
            pos = Position(fen, "chess", True)
            he = HeState(fen, 0, True)
            _playout(pos, he, rnd, [(1, 3)], 120, check_every=4)


def test_threefold_and_goldens_through_device_code():
    he = HeState("1rr3k1/1pp2ppp/p1n5/P2p1b2/3Pn3/R3PNP1/1P3PBP/2R1B1K1 b - - 4 17", 0)
    seq = ["e4d6", "f3h4", "f5e6", "h4f3", "e6f5", "f3h4", "f5e6", "h4f3"]
    for u in seq:
        he.do_move(he.move_from_uci(u))
        assert he.terminal() == 4
    he.do_move(he.move_from_uci("e6f5"))
    assert he.terminal() == 1  # draw by 3-fold
    # a golden plane literal straight from engine/tests/tests.cpp:1500-1512 through the device encoder
    he = HeState("5r2/ppp2pkp/3p4/2bP4/2Pnp1N1/3P2pP/PP2n1P1/R2Q1R1K[PBRQnbb] w - - 0 28", 1)
    for u in ("Q@f6", "g7g8", "R@h8"):
        he.do_move(he.move_from_uci(u))
    p = he.planes(0, 1, False).reshape(-1).astype(np.float64)
    assert p.sum() == 2395 and p.max() == 29 and (np.arange(p.size) * p).sum() == 4170903
    assert he.fen() == "5rkR/ppp2p1p/3p1Q2/2bP4/2Pnp1N1/3P2pP/PP2n1P1/R2Q1R1K[BPbbn] b - - 3 29"

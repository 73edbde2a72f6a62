"""Pins the CPU rules oracle (oracle/chess.c) with
 - perft known answers (public chess-programming-wiki values), and
 - the rule / FEN assertions of the reference's own suite, engine/tests/tests.cpp, restated case by case
   (3-fold :599-620, insufficient material :203-251, King of the hill :975-998, 3check :1139-1184,
    Chess960 :1297-1406, crazyhouse drops :1450-1481)."""
import pytest

from oracle.chess import Position, T_DRAW, T_NONE

PERFT = [
    ("chess", False, "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", [20, 400, 8902, 197281, 4865609]),
    ("chess", False, "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", [48, 2039, 97862, 4085603]),
    ("chess", False, "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", [14, 191, 2812, 43238, 674624]),
    ("chess", False, "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", [6, 264, 9467, 422333]),
    ("chess", False, "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", [44, 1486, 62379, 2103487]),
    ("chess", False, "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", [46, 2079, 89890]),
    ("chess", True, "bqnb1rkr/pp3ppp/3ppn2/2p5/5P2/P2P4/NPP1P1PP/BQ1BNRKR w HFhf - 2 9", [21, 528, 12189, 326672]),
    ("chess", True, "2nnrbkr/p1qppppp/8/1ppb4/6PP/3PP3/PPP2P2/BQNNRBKR w HEhe - 1 9", [21, 807, 18002, 667366]),
    ("crazyhouse", False, "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR[] w KQkq - 0 1", [20, 400, 8902, 197281, 4888832]),
]


@pytest.mark.parametrize("variant,is960,fen,counts", PERFT)
def test_perft(variant, is960, fen, counts):
    p = Position(fen, variant, is960)
    for d, expect in enumerate(counts, start=1):
        assert p.perft(d) == expect, (fen, d)


def test_threefold_repetition():  # tests.cpp:599-620
    s = Position("1rr3k1/1pp2ppp/p1n5/P2p1b2/3Pn3/R3PNP1/1P3PBP/2R1B1K1 b - - 4 17")
    b0, w1, b1, w2, b2 = "e4d6", "f3h4", "f5e6", "h4f3", "e6f5"
    for mv in [b0, w1, b1, w2, b2, w1, b1, w2]:
        s.push_uci(mv)
        assert s.terminal() == T_NONE
    s.push_uci(b2)
    assert s.terminal() == T_DRAW


def _insufficient(fen, variant="chess"):
    p = Position(fen, variant)
    return p.terminal() == T_DRAW


def test_insufficient_material():  # tests.cpp:203-251
    assert _insufficient("8/8/2k5/8/8/4K3/8/8 w - - 0 1")
    assert _insufficient("8/8/2k5/8/5B2/4K3/8/8 w - - 0 1")
    assert _insufficient("8/8/2k5/8/5N2/4K3/8/8 w - - 0 1")
    assert _insufficient("8/8/2k5/8/8/3NKN2/8/8 w - - 0 1")
    assert not _insufficient("kn6/8/NK6/8/8/8/8/8 w - - 0 2")
    assert not _insufficient("rnbqkb1r/pp2pppp/3p1n2/8/3NP3/8/PPP2PPP/RNBQKB1R w KQkq - 1 5")
    assert not _insufficient("8/8/2k5/8/8/4K3/8/8 w - - 0 1", "kingofthehill")


def test_king_of_the_hill():  # tests.cpp:975-998
    black = ["8/7p/8/1b1pk3/4p3/5n2/1K3P2/8 w - - 4 43", "8/8/2p2p2/3k3p/4p2P/r6P/8/5K2 w - - 0 47",
             "rnbq1bnr/pppp1ppp/8/8/P2kp3/1RN4N/1PPPPPPP/2BQKB1R w K - 0 1",
             "rnbq1bnr/ppp1pppp/3p4/8/P3k3/1RN4N/1PPPPPPP/2BQKB1R w K - 0 1"]
    white = ["5k2/1p5p/8/p7/2P1Kp2/5N1P/Pr3PP1/3RR3 b - - 1 29", "6k1/5p2/bP3B2/3N4/3K4/5P1p/P7/8 b - - 2 37",
             "rnbqkb1r/1pppppp1/p7/3K1n1p/8/5N2/PPPP1PPP/RNBQ1B1R w kq - 0 1",
             "rnb1kb1r/1ppqppp1/p6n/3PK2p/8/8/PPPP1PPP/RNBQ1BNR w kq - 0 1"]
    for fen in black:
        assert Position(fen, "kingofthehill").check_result() == -1, fen
    for fen in white:
        assert Position(fen, "kingofthehill").check_result() == 1, fen


def test_three_check():  # tests.cpp:1139-1184
    p = Position(variant="3check")
    assert p.fen() == "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 3+3 0 1"
    p = Position("1r4k1/1p2bp1p/3p2p1/PprPp2n/1R2PPq1/3Q4/1P1B1NPP/5RK1 b - - 3+3 2 22", "3check")
    p.push_uci("g4g2")
    assert p.fen() == "1r4k1/1p2bp1p/3p2p1/PprPp2n/1R2PP2/3Q4/1P1B1NqP/5RK1 w - - 3+2 0 23"
    p.push_uci("g1g2", "h5f4")
    assert p.fen() == "1r4k1/1p2bp1p/3p2p1/PprPp3/1R2Pn2/3Q4/1P1B1NKP/5R2 w - - 3+1 0 24"
    p = Position("2r3k1/1p2bp1p/3p2p1/Pp1P4/1R2Pp2/7Q/1P3N1P/2r2R1K w - - 3+1 4 27", "3check")
    p.push_uci("h3c8")
    assert p.fen() == "2Q3k1/1p2bp1p/3p2p1/Pp1P4/1R2Pp2/8/1P3N1P/2r2R1K b - - 2+1 0 27"
    p = Position("4Rr1k/3P3p/5pp1/8/8/4p3/1P3p1P/5R1K w - - 2+1 0 39", "3check")
    p.push_uci("e8f8")
    assert p.fen() == "5R1k/3P3p/5pp1/8/8/4p3/1P3p1P/5R1K b - - 1+1 0 39"
    p = Position("5R2/3P2kp/5pp1/8/8/4p3/1P3p1P/5R1K w - - 1+1 1 40", "3check")
    p.push_uci("f8f7")
    assert p.fen() == "8/3P1Rkp/5pp1/8/8/4p3/1P3p1P/5R1K b - - 0+1 2 40"
    assert p.check_result() == 1
    assert Position("8/pk6/8/p1p4p/2P1p3/8/6p1/B2Kr3 w - - 3+0 2 44", "3check").check_result() == -1
    assert Position("8/ppp2k2/5b2/1P1n4/P3bP2/3PP1P1/5K1r/5R2 w - - 1+0 3 27", "3check").check_result() == -1
    assert Position("8/k1R1P3/p2p4/2pP4/8/1p2P2p/P6P/K7 b - - 0+3 1 34", "3check").check_result() == 1
    assert Position("6R1/6k1/8/5P1p/5P1P/8/6K1/8 b - - 0+2 4 56", "3check").check_result() == 1


def _legal(p, mv):
    return mv in p.legal_uci()


def test_chess960_castling():  # tests.cpp:1331-1401
    p = Position("bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BN1RK1RQ w GDgd - 0 1", "chess", True)
    assert _legal(p, "e1d1") and _legal(p, "e1g1")
    assert not _legal(Position("bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BNNRKBRQ w GDgd - 0 4", "chess", True), "e1g1")
    assert not _legal(Position("1nnrkbrq/p1pppppp/1p6/1b6/4P3/8/PPPP1PPP/BN1RK1RQ w GDgd - 0 4", "chess", True), "e1g1")
    assert not _legal(Position("bnnrk1rq/p1pp1ppp/1p2p3/2b5/4P3/5P2/PPPP2PP/BN1RK1RQ w GDgd - 0 4", "chess", True), "e1g1")
    p = Position("bnnrkbr1/ppppppp1/8/4q2p/8/5P2/PPPP2PP/BN1RK1RQ w GDgd - 0 4", "chess", True)
    assert not _legal(p, "e1d1") and not _legal(p, "e1g1")
    assert not _legal(Position("nrbbqnkr/pppppppp/8/8/8/8/PPPPPPPP/NRBBQNKR w HBhb - 0 4", "chess", True), "g1h1")
    assert not _legal(Position("bnnrkbrq/pppppppp/8/8/8/4P3/PPPP1PPP/BNNRK1RQ w GDgd - 0 4", "chess", True), "e1d1")
    p = Position("bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BN1RK1RQ w gd - 0 1", "chess", True)
    assert not _legal(p, "e1d1") and not _legal(p, "e1g1")
    start = "bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BN1RK1RQ w GDgd - 0 1"
    cases = {"d1c1": "bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BNR1K1RQ b Ggd - 1 1",
             "g1f1": "bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BN1RKR1Q b Dgd - 1 1",
             "e1f1": "bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BN1R1KRQ b gd - 1 1",
             "e1d1": "bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BNKR2RQ b gd - 1 1",
             "e1g1": "bnnrkbrq/pppppppp/8/8/8/8/PPPPPPPP/BN1R1RKQ b gd - 1 1"}
    for mv, fen in cases.items():
        assert Position(start, "chess", True).push_uci(mv).fen() == fen, mv
    p = Position("nrbbqnkr/pppppppp/8/8/8/8/PPPPPPPP/NR4KR w HBhb - 0 1", "chess", True).push_uci("g1h1")
    assert p.fen() == "nrbbqnkr/pppppppp/8/8/8/8/PPPPPPPP/NR3RK1 b hb - 1 1"


def test_crazyhouse_rules():  # tests.cpp:1450-1481 and the FEN checks of :1493-1527
    assert Position(variant="crazyhouse").fen() == "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR[] w KQkq - 0 1"
    p = Position("1k1r3r/pppb1p2/2nbqn1p/3p2p1/3PP1P1/3Q1PP1/PPN2NBP/R1B2RK1[p] b - - 0 12", "crazyhouse")
    assert _legal(p, "P@c4")
    for mv in ["P@d5", "P@d4", "P@e4", "P@b1", "P@d1", "P@e1", "P@e8", "P@f8", "P@g8"]:
        assert not _legal(p, mv), mv
    p = Position("4R2b/1N3rkb/1p2P1pp/p2P4/2P1P3/8/PP4Q1/3R3K[QRBBNNNPPPPpp] w - - 2 53", "crazyhouse")
    assert _legal(p, "N@h5")
    p.push_uci("N@h5")
    assert p.check_result() == 1
    p = Position("r2qk3/1pP2r1n/p1nP4/8/3P1Bb1/2Pp1PP1/PPp2PP1/3q1K1R[Bbnnppr] w - - 2 29", "crazyhouse")
    assert p.check_result() is None
    assert _legal(p, "B@e1")
    p = Position("5r2/ppp2pkp/3p4/2bP4/2Pnp1N1/3P2pP/PP2n1P1/R2Q1R1K[PBRQnbb] w - - 0 28", "crazyhouse")
    p.push_uci("Q@f6", "g7g8", "R@h8")
    assert p.fen() == "5rkR/ppp2p1p/3p1Q2/2bP4/2Pnp1N1/3P2pP/PP2n1P1/R2Q1R1K[BPbbn] b - - 3 29"


def test_clone_keeps_history_and_fen():  # State fuzz (tests.cpp:652-740): clone() FEN equality, playouts end
    import random
    rnd = random.Random(42)
    for variant in ("chess", "crazyhouse", "kingofthehill", "3check"):
        p = Position(variant=variant)
        plies = 0
        while p.terminal() == T_NONE and plies < 300:
            mv = rnd.choice(p.legal_moves())
            q = p.clone()
            assert q.fen() == p.fen() and q.key() == p.key()
            p.do_move(mv)
            q.do_move(mv)
            assert q.fen() == p.fen()
            plies += 1
        assert plies > 5

"""Pins oracle/planes.c with the PlaneStatistics literals of the reference's suite (engine/tests/tests.cpp):
chess v1 :287-331, chess v3.0 :333-548, crazyhouse V1/V2/V3 :1493-1603, chess960 v3 :1643-1669,
anti start FEN (lichess v1 layout) :164-175, racing kings :254-283, atomic v3 start :1606-1620."""
import pytest

from oracle.chess import Position, plane_stats, planes


def _stats(pos, mode, version, normalize):
    return plane_stats(planes(pos, mode, version, normalize))


def _check(st, sum=None, max=None, key=None, argmax=None, rel=None):
    if rel is None:
        if sum is not None:
            assert st["sum"] == sum
        if key is not None:
            assert st["key"] == key
    else:
        if sum is not None:
            assert st["sum"] == pytest.approx(sum, rel=rel)
        if key is not None:
            assert st["key"] == pytest.approx(key, rel=rel)
    if max is not None:
        assert st["max"] == pytest.approx(max, rel=1e-6)
    if argmax is not None:
        assert st["argmax"] == argmax


def test_chess_v1():
    p = Position().push_uci("e2e4", "e7e5", "d1h5", "b8c6", "f1c4", "g8f6", "h5f7")
    _check(_stats(p, "chess", 1, False), sum=557, max=4, key=617997, argmax=1024)
    st = _stats(p, "chess", 1, True)
    assert 0.99 < st["max"] < 1.01 and 301.512 < st["sum"] < 301.513 and 348329.41 < st["key"] < 348329.42
    p2 = Position().push_uci("e2e4", "c7c5", "c2c4", "b8c6", "g1e2", "g8f6", "b1c3", "c6b4", "g2g3", "b4d3")
    _check(_stats(p2, "chess", 1, False), sum=816, max=6, key=909458, argmax=1024)


def test_chess_v3():
    _check(_stats(Position(), "chess", 3, False), sum=1312, argmax=3008, max=8, key=3430384)
    _check(_stats(Position(), "chess", 3, True), sum=472, argmax=8, max=1, key=819860)
    p = Position("rnbqk1nr/pppp1ppp/8/4p3/1b1PP3/8/PPP2PPP/RNBQKBNR w KQkq - 1 3")
    _check(_stats(p, "chess", 3, False), sum=1377, argmax=3008, max=8, key=3513153)
    p = Position().push_uci("e2e4", "c7c5")
    _check(_stats(p, "chess", 3, False), sum=1316, argmax=3008, max=8, key=3436012)
    p = Position("r1br2k1/p4ppp/2p2n2/Q1b1p3/8/NP3N1P/P1P1BPP1/R1B1K2R b KQ - 0 12")
    _check(_stats(p, "chess", 3, True), sum=284, argmax=8, max=1, key=529254)
    p = Position().push_uci("e2e4", "c7c5", "d2d3", "a7a6", "e4e5", "d7d5")
    _check(_stats(p, "chess", 3, False), sum=1325, argmax=3008, max=8, key=3451283)
    p = Position().push_uci("e2e4", "c7c5", "e4e5", "d7d5")
    _check(_stats(p, "chess", 3, False), sum=1321, argmax=3008, max=8, key=3443613)
    p = Position("r3k1nr/pbp4p/p2p2pb/4P3/3P4/N2q1n2/PPP2PPP/5K1R w kq - 0 14")
    _check(_stats(p, "chess", 3, False), sum=659, argmax=3008, max=8, key=1715209)
    p = Position("2kr3r/pbqp1ppp/2n2n2/4b3/4P3/2NPB3/PPP1QPPP/R4RK1 b - - 4 11")
    _check(_stats(p, "chess", 3, True), sum=179.12, key=442487.2, argmax=8, max=1, rel=0.001)


CZ_FEN = "5r2/ppp2pkp/3p4/2bP4/2Pnp1N1/3P2pP/PP2n1P1/R2Q1R1K[PBRQnbb] w - - 0 28"


def test_crazyhouse_v1_v2_v3():
    start = Position(variant="crazyhouse")
    mid = Position(CZ_FEN, "crazyhouse").push_uci("Q@f6", "g7g8", "R@h8")
    _check(_stats(start, "crazyhouse", 1, False), sum=416, max=1, key=746928, argmax=8)
    _check(_stats(mid, "crazyhouse", 1, False), sum=2395, max=29, key=4170903, argmax=1792)
    _check(_stats(mid, "crazyhouse", 1, True), sum=45.512, key=37011.632, max=1, argmax=8, rel=0.001)
    _check(_stats(start, "crazyhouse", 2, False), sum=416, max=1, key=746928, argmax=8)
    _check(_stats(mid, "crazyhouse", 2, False), sum=2399, max=29, key=4180615, argmax=1792)
    _check(_stats(mid, "crazyhouse", 2, True), sum=49.512, key=46723.632, max=1, argmax=8, rel=0.001)
    _check(_stats(start, "crazyhouse", 3, False), sum=1312, max=8, key=3430384, argmax=3008)
    _check(_stats(mid, "crazyhouse", 3, False), sum=1307, max=8, key=3700213, argmax=3008)
    _check(_stats(mid, "crazyhouse", 3, True), sum=193.8, key=474696, max=1, argmax=8, rel=0.001)


def test_chess960_v3():
    p = Position("b1qnrnkr/p2ppppp/1p6/2p1b3/2P5/4N1P1/PP1PPP1P/BBQNRK1R b he - 1 4", "chess", True)
    assert p.fen() == "b1qnrnkr/p2ppppp/1p6/2p1b3/2P5/4N1P1/PP1PPP1P/BBQNRK1R b he - 1 4"
    _check(_stats(p, "chess", 3, False), sum=1312, max=8, key=3512322, argmax=3008)
    _check(_stats(p, "chess", 3, True), sum=409.28, key=823554.8, max=1, argmax=8, rel=0.001)


def test_lichess_layouts():
    anti = Position(variant="anti")
    st = _stats(anti, "lichess", 1, False)
    assert planes(anti, "lichess", 1, False).size == 4032
    _check(st, sum=224, max=1, key=417296)
    race = Position(variant="racingkings")
    _check(_stats(race, "lichess", 1, False), sum=208, argmax=68, max=1, key=425624)
    race_b = Position("8/8/8/8/8/6K1/krbnNBR1/qrbnNBRQ b - - 1 1", "racingkings")
    _check(_stats(race_b, "lichess", 1, False), sum=208, argmax=67, max=1, key=450207)
    atomic = Position(variant="atomic")
    assert atomic.fen() == "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
    _check(_stats(atomic, "lichess", 3, False), sum=1440, max=8, key=5932976, argmax=4736)

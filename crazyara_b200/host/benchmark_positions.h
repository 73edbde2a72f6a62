// The tactical suite of the reference's UCI `benchmark` command (engine/tests/benchmarkpositions.cpp:28-57): fifteen
// crazyhouse positions, each with the blunder a weak search plays and a sound alternative.  Position data only; the
// command itself is in uci_main.cpp.
#pragma once

namespace crazyara {

struct TestPosition {
    const char* fen;
    const char* blunderMove;
    const char* alternativeMove;
};

static const TestPosition kBenchmarkPositions[] = {
    {"r1b2bk1/pp3ppp/2pn1bn1/4r3/3Q3P/2N1PB1p/PPP1PPP1/3RK2R/NQp w K - 0 24", "h4h5", "Q@h2"},
    {"r1bq1rk1/pppp1ppp/2n2n2/1Bb1p3/4P3/2NP1N2/PPP2PPP/R1BQ1RK1/ b - - 11 6", "c6d4", "f6g4"},
    {"r2qr1k1/ppp2ppp/2n1bp2/8/1b1P4/2N5/1PP1NPPP/R1BQKB1R/PPNp w KQ - 0 11", "N@e3", "P@h6"},
    {"r1bq1bk1/ppp2ppp/5p2/3pNn2/3PpB2/P1N5/1PP1QPPP/R4RK1/RNb b - - 0 13", "f6e5", "c8e6"},
    {"r2q1rk1/pp3ppp/2np2b1/6BB/3p4/3P2N1/PPrQBPKP/R7/PNPPPn w - - 0 29", "P@c7", "d2c2"},
    {"r2qr3/p1p3pk/2p3pp/3b1p1n/3P4/4PPB1/PPPBQ1PP/R4RK1/NBNpn b - - 0 30", "P@h4", "h5g3"},
    {"1r1qr3/p1p3pk/2p3pp/3b1p1n/3P3p/1P2PP2/P1PB2PP/R3BQK1/NNBNr b - - 0 34", "b8b3", "R@g5"},
    {"r2q4/1pp2kPp/5prP/2pP1N2/5PB1/2N2P2/PP3PPN/2r1rQ1K/Nbpbbp w - - 0 54", "N@h8", "f1e1"},
    {"r2q2kN/1pp3Pp/5prP/2pP1N2/5PB1/2N2P2/PP3PPN/4r2K/Rqbpbbp w - - 0 62", "R@g1", "R@f1"},
    {"r1bqk1r1/2p1bppp/p1p2n1P/3P4/2B5/2N2p2/PPP2PRP/R1BQK3/PNPn w Qq - 24 13", "d1f3", "g2g7"},
    {"r4rk1/2pPbppp/p3p3/8/4P2n/2N4Q/PPP2PPP/R1B1K2R/BNPPqbn b KQ - 2 16", "B@g5", "Q@g6"},
    {"r3k2r/1pp2bpp/p3b1p1/3np1N1/6R1/bP5p/PnPP1P1P/R3Q1K1[QNbpp] w - - 0 26", "e1e5", "g5f7"},
    {"r1b1kb1r/pp3npp/3p1p2/Q1n5/3PP3/2P5/pBpP1PPP/R3KB1R[Qnnp] b Kkq - 0 21", "c5e4", "N@d3"},
    {"3q1rk1/p1p2p1B/2p4b/8/1PnP2Pb/4P2p/1PPR2PP/2R2R1K/PBQPnnpn b - - 0 41", "g8h8", "g8h7"},
    {"r1b4r/ppp1kp1p/2bp4/6Pn/4n2N/8/P1P1BPPP/R4RK1/QBpnqppp w - - 0 21", "e2h5", "B@f6"},  // lost, but avoid mate in 4
};

}  // namespace crazyara

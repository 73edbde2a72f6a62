"""Classic ("flat") move labels of the reference: the UCI label list per build mode and the move -> label index
mapping used by the training-data exporter (StateConstants::action_to_index<classic, mirrored?>).

The list is generated the way the reference generates it (engine/src/environments/chess_related/
outputrepresentation.cpp:108-163 generate_uci_labels / generate_dropping_moves): queen-line and knight destinations of
every square in file-major order, the promotion moves, then the drops; black's moves are looked up after mirroring the
ranks (sfutil.cpp:183-197).  2272 labels for crazyhouse, 1968 for chess, 2316 for the lichess variants
(boardstate.h:51-60)."""

MODES = {"crazyhouse": 0, "chess": 1, "lichess": 2}
_CACHE = {}


def uci_labels(mode):
    mode = MODES[mode] if isinstance(mode, str) else mode
    if mode in _CACHE:
        return _CACHE[mode][0]
    labels = []
    files, ranks = "abcdefgh", "12345678"
    knight = [(-2, -1), (-1, -2), (-2, 1), (1, -2), (2, -1), (-1, 2), (2, 1), (1, 2)]
    for f in range(8):
        for r in range(8):
            dest = [(i, r) for i in range(8)] + [(f, i) for i in range(8)]
            dest += [(f + i, r + i) for i in range(-7, 8)] + [(f + i, r - i) for i in range(-7, 8)]
            dest += [(f + a, r + b) for a, b in knight]
            for f2, r2 in dest:
                if (f, r) != (f2, r2) and 0 <= f2 < 8 and 0 <= r2 < 8:
                    labels.append(files[f] + ranks[r] + files[f2] + ranks[r2])
    promo = "qrbnk" if mode == 2 else "qrbn"
    for f in range(8):
        for p in promo:
            labels.append(f"{files[f]}2{files[f]}1{p}")
            labels.append(f"{files[f]}7{files[f]}8{p}")
            if f > 0:
                labels.append(f"{files[f]}2{files[f - 1]}1{p}")
                labels.append(f"{files[f]}7{files[f - 1]}8{p}")
            if f < 7:
                labels.append(f"{files[f]}2{files[f + 1]}1{p}")
                labels.append(f"{files[f]}7{files[f + 1]}8{p}")
    if mode != 1:
        for f in range(8):
            for r in range(8):
                for p in "PNBRQ":
                    if p == "P" and r in (0, 7):
                        continue
                    labels.append(f"{p}@{files[f]}{ranks[r]}")
    _CACHE[mode] = (labels, {u: i for i, u in enumerate(labels)})
    return labels


def mirror_uci(uci):
    """Rank mirror of a UCI move string (black's moves are stored from white's point of view)."""
    return "".join(str(9 - int(c)) if c in "12345678" else c for c in uci)


def classic_index(mode, uci, mirror):
    """Index of a legal move's UCI string in the label list (mirrored for the second player); KeyError if absent."""
    mode = MODES[mode] if isinstance(mode, str) else mode
    uci_labels(mode)
    return _CACHE[mode][1][mirror_uci(uci) if mirror else uci]
